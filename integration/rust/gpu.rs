//! src/vectordb/gpu.rs — Rust binding of libcsgpu.so (include/csgpu.h), the B200-native replacement of the arroy block in
//! `VectorStore::search` (src/vectordb/store.rs:446-459).
//!
//! STATUS: source only. There is no cargo/rustc in the image this was written in, so this file has never been compiled;
//! what IS checked mechanically (tests/test_rust_binding.py): the `extern "C"` block below is generated from the header by
//! tools/gen_rust_ffi.py and must match it symbol for symbol (names, arity, pointer constness, integer widths), the
//! `#[repr(C)]` structs must list the header's fields in order with the same widths, and every extern function must be
//! used by a wrapper. The same ABI is exercised end to end from Python ctypes (codesearch_b200/_lib.py) and from C++
//! (include/csgpu_store.hpp, tests/cpp/store_test.cpp).
//!
//! Threading contract (matches `&self` / `&mut self` of VectorStore behind `Arc<RwLock<..>>`, src/index/manager.rs:129,
//! src/server/mod.rs:27): every `search*` method takes `&self` and is re-entrant from any number of threads (rayon
//! par_iter at src/search/mod.rs:508-511); `append / remove / build / clear / load` take `&mut self`.
#![allow(dead_code)]

use anyhow::{anyhow, Result};
use std::ffi::{CStr, CString};
use std::os::raw::{c_char, c_int, c_void};
use std::path::Path;

pub const CSGPU_ABI_VERSION: u32 = 8;
pub const CSGPU_OK: c_int = 0;
pub const CSGPU_ERR_DIM: c_int = 1; // "Query embedding dimension mismatch: expected {}, got {}"   store.rs:432-438
pub const CSGPU_ERR_NOT_BUILT: c_int = 2; // "Index not built. Call build_index() after inserting chunks."   store.rs:440-444
pub const CSGPU_ERR_CUDA: c_int = 3;
pub const CSGPU_ERR_NCCL: c_int = 4;
pub const CSGPU_ERR_OOM: c_int = 5;
pub const CSGPU_ERR_ARG: c_int = 6;
pub const CSGPU_DTYPE_F32: u32 = 0;
pub const CSGPU_DTYPE_BF16: u32 = 1;
pub const CSGPU_MAX_K: u32 = 1024;
pub const CSGPU_TAG_LANG_SHIFT: u32 = 27;
pub const CSGPU_TAG_FILE_MASK: u32 = 0x07FF_FFFF;
pub const CSGPU_TAG_NONE: u32 = 0xFFFF_FFFF;
pub const CSGPU_EXCHANGE_HANDLE_BYTES: usize = 64;

/// Opaque `csgpu_index`.
#[repr(C)]
pub struct CsgpuIndex {
    _private: [u8; 0],
}

/// `csgpu_stats_t`
#[repr(C)]
#[derive(Debug, Default, Clone, Copy)]
pub struct CsgpuStats {
    pub live_rows: u64,
    pub pending_rows: u64,
    pub tombstones: u64,
    pub zero_norm_rows: u64,
    pub nonfinite_rows: u64,
    pub bytes_on_device: u64,
    pub dim: u32,
    pub dtype: u32,
    pub n_devices: u32,
    pub built: u32,
    pub last_search_us: f32,
    pub abi_version: u32,
    pub rows_per_device: [u64; 8],
    pub coalesced_passes: u64,
    pub coalesced_queries: u64,
    pub prefilter_rescored: u64,
    pub shadow_bytes: u64,
    pub byte_shadow_bytes: u64,
    pub byte_searches: u64,
    pub byte_fallbacks: u64,
    pub byte_candidates: u64,
    pub byte_rescored: u64,
    /// contraction of the last GEMM-shaped batch: 0 none, 1 fp32 SIMT, 2 tcgen05 bf16, 3 tcgen05 tf32 off the fp32 rows
    pub batch_route: u32,
    /// largest |d_filter - d_f32| the exact rescoring of that batch saw
    pub filter_max_err: f32,
}

/// `csgpu_predicate_t`: a row passes iff its language bit is set AND file_lo <= file_id <= file_hi AND (file_bitmap is
/// null OR bit file_id of it is set).
#[repr(C)]
#[derive(Debug, Clone, Copy)]
pub struct CsgpuPredicate {
    pub lang_mask: u32,
    pub file_lo: u32,
    pub file_hi: u32,
    pub reserved: u32,
    pub file_bitmap: *const u64,
    pub n_file_bits: u64,
}

extern "C" {
    // ---- GENERATED from include/csgpu.h by tools/gen_rust_ffi.py: do not edit by hand ----
    pub fn csgpu_create(out: *mut *mut CsgpuIndex, dim: u32, dtype: u32, devices: *const i32, n_devices: u32) -> c_int;
    pub fn csgpu_destroy(ix: *mut CsgpuIndex);
    pub fn csgpu_append(ix: *mut CsgpuIndex, rows: *const f32, ids: *const u32, n: u64) -> c_int;
    pub fn csgpu_remove(ix: *mut CsgpuIndex, ids: *const u32, n: u64, n_removed: *mut u64) -> c_int;
    pub fn csgpu_reserve(ix: *mut CsgpuIndex, total_rows: u64) -> c_int;
    pub fn csgpu_build(ix: *mut CsgpuIndex) -> c_int;
    pub fn csgpu_clear(ix: *mut CsgpuIndex) -> c_int;
    pub fn csgpu_save(ix: *const CsgpuIndex, dir: *const c_char) -> c_int;
    pub fn csgpu_load(ix: *mut CsgpuIndex, dir: *const c_char) -> c_int;
    pub fn csgpu_search(ix: *const CsgpuIndex, q: *const f32, q_len: u32, k: u32, out_ids: *mut u32, out_dist: *mut f32, out_n: *mut u32) -> c_int;
    pub fn csgpu_set_coalescing(ix: *mut CsgpuIndex, enabled: u32, window_us: u32) -> c_int;
    pub fn csgpu_search_batch(ix: *const CsgpuIndex, q: *const f32, q_len: u32, b: u32, k: u32, out_ids: *mut u32, out_dist: *mut f32, out_n: *mut u32) -> c_int;
    pub fn csgpu_set_tensor_prefilter(ix: *mut CsgpuIndex, enabled: u32) -> c_int;
    pub fn csgpu_set_byte_prefilter(ix: *mut CsgpuIndex, enabled: u32) -> c_int;
    pub fn csgpu_search_variants(ix: *const CsgpuIndex, q: *const f32, q_len: u32, b: u32, k: u32, out_ids: *mut u32, out_dist: *mut f32, out_n: *mut u32) -> c_int;
    pub fn csgpu_search_filtered(ix: *const CsgpuIndex, q: *const f32, q_len: u32, k: u32, id_bitmap: *const u64, n_bits: u64, out_ids: *mut u32, out_dist: *mut f32, out_n: *mut u32) -> c_int;
    pub fn csgpu_append_tagged(ix: *mut CsgpuIndex, rows: *const f32, ids: *const u32, tags: *const u32, n: u64) -> c_int;
    pub fn csgpu_search_tagged(ix: *const CsgpuIndex, q: *const f32, q_len: u32, k: u32, pred: *const CsgpuPredicate, out_ids: *mut u32, out_dist: *mut f32, out_n: *mut u32) -> c_int;
    pub fn csgpu_search_variants_tagged(ix: *const CsgpuIndex, q: *const f32, q_len: u32, b: u32, k: u32, pred: *const CsgpuPredicate, out_ids: *mut u32, out_dist: *mut f32, out_n: *mut u32) -> c_int;
    pub fn csgpu_get_tags(ix: *const CsgpuIndex, ids: *const u32, n: u64, out_tags: *mut u32) -> c_int;
    pub fn csgpu_search_keys_device(ix: *const CsgpuIndex, q_dev: *const f32, k: u32, out_keys_dev: *mut u64, stream: *mut c_void) -> c_int;
    pub fn csgpu_merge_keys_device(ix: *const CsgpuIndex, keys_dev: *const u64, n_lists: u32, k: u32, out_keys_dev: *mut u64, stream: *mut c_void) -> c_int;
    pub fn csgpu_merge_keys_batch_device(ix: *const CsgpuIndex, keys_dev: *const u64, n_lists: u32, nq: u32, k: u32, out_keys_dev: *mut u64, stream: *mut c_void) -> c_int;
    pub fn csgpu_exchange_create(ix: *mut CsgpuIndex, world: u32, rank: u32, out_handle: *mut c_void) -> c_int;
    pub fn csgpu_exchange_connect(ix: *mut CsgpuIndex, handles: *const c_void) -> c_int;
    pub fn csgpu_exchange_connect_local(ix: *mut CsgpuIndex, peers: *const *mut CsgpuIndex) -> c_int;
    pub fn csgpu_search_keys_exchange_device(ix: *const CsgpuIndex, q_dev: *const f32, k: u32, out_keys_dev: *mut u64, stream: *mut c_void) -> c_int;
    pub fn csgpu_search_exchange(ix: *const CsgpuIndex, q: *const f32, q_len: u32, k: u32, out_ids: *mut u32, out_dist: *mut f32, out_n: *mut u32) -> c_int;
    pub fn csgpu_exchange_status(ix: *const CsgpuIndex, timed_out: *mut u32) -> c_int;
    pub fn csgpu_exchange_set_timeout_ms(ix: *mut CsgpuIndex, ms: u32) -> c_int;
    pub fn csgpu_exchange_wait_stats(ix: *const CsgpuIndex, out_ns: *mut u64, max_queries: u32, n_queries: *mut u32) -> c_int;
    pub fn csgpu_exchange_destroy(ix: *mut CsgpuIndex);
    pub fn csgpu_search_tagged_keys_device(ix: *const CsgpuIndex, q_dev: *const f32, k: u32, pred: *const CsgpuPredicate, exchange: u32, out_keys_dev: *mut u64, stream: *mut c_void) -> c_int;
    pub fn csgpu_encode_keys(ids: *const u32, dist: *const f32, n: u32, k: u32, out_keys: *mut u64);
    pub fn csgpu_decode_keys(keys: *const u64, k: u32, out_ids: *mut u32, out_dist: *mut f32, out_n: *mut u32);
    pub fn csgpu_append_synthetic(ix: *mut CsgpuIndex, seed: u64, first_row: u64, n: u64, id_base: u32) -> c_int;
    pub fn csgpu_append_synthetic_tagged(ix: *mut CsgpuIndex, seed: u64, first_row: u64, n: u64, id_base: u32) -> c_int;
    pub fn csgpu_synth_rows_host(ix: *const CsgpuIndex, seed: u64, first_row: u64, n: u64, out_rows: *mut f32) -> c_int;
    pub fn csgpu_stats(ix: *const CsgpuIndex, out: *mut CsgpuStats) -> c_int;
    pub fn csgpu_kernel_launches() -> u64;
    pub fn csgpu_last_error() -> *const c_char;
    pub fn csgpu_abi_version() -> u32;
    // ---- END GENERATED ----
}

/// `(lang_id << 27) | file_id`; lang_id = `Language` declaration order (src/file/language.rs:5-29).
pub fn make_tag(lang_id: u32, file_id: u32) -> u32 {
    ((lang_id & 31) << CSGPU_TAG_LANG_SHIFT) | (file_id & CSGPU_TAG_FILE_MASK)
}

fn last_error() -> String {
    unsafe { CStr::from_ptr(csgpu_last_error()) }.to_string_lossy().into_owned()
}

/// Codes 1 and 2 carry the reference's literal messages (store.rs:432-444), so `anyhow!("{msg}")` reproduces today's
/// errors verbatim; the code rides along for callers that want to branch on it.
fn check(rc: c_int) -> Result<()> {
    if rc == CSGPU_OK {
        Ok(())
    } else {
        Err(anyhow!("{}", last_error()).context(format!("csgpu error code {rc}")))
    }
}

fn c_path(p: &Path) -> Result<CString> {
    CString::new(p.to_string_lossy().as_bytes()).map_err(|_| anyhow!("path contains a NUL byte"))
}

/// Allow-set over chunk ids for `search_filtered` (bit i set <=> chunk id i may be returned).
pub struct RowFilter {
    pub bitmap: Vec<u64>,
    pub n_bits: u64,
}

impl RowFilter {
    pub fn from_ids(ids: impl IntoIterator<Item = u32>, n_bits: u64) -> Self {
        let mut bitmap = vec![0u64; ((n_bits + 63) / 64) as usize];
        for id in ids {
            if (id as u64) < n_bits {
                bitmap[(id >> 6) as usize] |= 1u64 << (id & 63);
            }
        }
        Self { bitmap, n_bits }
    }
}

/// Owning handle of a device index. `Send + Sync`: the library's search entry points are re-entrant, and mutation goes
/// through `&mut self`.
pub struct GpuIndex {
    raw: *mut CsgpuIndex,
    dim: usize,
}
unsafe impl Send for GpuIndex {}
unsafe impl Sync for GpuIndex {}

impl Drop for GpuIndex {
    fn drop(&mut self) {
        unsafe { csgpu_destroy(self.raw) }
    }
}

impl GpuIndex {
    /// One device (ordinal 0), fp32 rows. Fails (code 3) when there is no B200: there is no CPU fallback by design.
    pub fn new(dim: usize) -> Result<Self> {
        Self::with_devices(dim, CSGPU_DTYPE_F32, &[])
    }

    /// `devices` empty => device 0. Several devices => rows are sharded row-wise inside this one process and every search
    /// is one fused launch per device (knob: CODESEARCH_GPU_DEVICES).
    pub fn with_devices(dim: usize, dtype: u32, devices: &[i32]) -> Result<Self> {
        if unsafe { csgpu_abi_version() } != CSGPU_ABI_VERSION {
            return Err(anyhow!("libcsgpu.so ABI version mismatch"));
        }
        let mut raw = std::ptr::null_mut();
        let (ptr, n) = if devices.is_empty() { (std::ptr::null(), 1) } else { (devices.as_ptr(), devices.len() as u32) };
        check(unsafe { csgpu_create(&mut raw, dim as u32, dtype, ptr, n) })?;
        Ok(Self { raw, dim })
    }

    pub fn dim(&self) -> usize {
        self.dim
    }

    // ---- write side (store.rs:618-686 insert, :548-610 delete, :386-430 build, :690-706 clear) ----------------------
    pub fn reserve(&mut self, total_rows: u64) -> Result<()> {
        check(unsafe { csgpu_reserve(self.raw, total_rows) })
    }

    /// `rows`: n x dim row-major; `ids`: n chunk ids. Copies; marks the index dirty (indexed = false).
    pub fn append(&mut self, rows: &[f32], ids: &[u32]) -> Result<()> {
        if rows.len() != ids.len() * self.dim {
            return Err(anyhow!("Embedding dimension mismatch: expected {}, got {}", self.dim, rows.len() / ids.len().max(1)));
        }
        check(unsafe { csgpu_append(self.raw, rows.as_ptr(), ids.as_ptr(), ids.len() as u64) })
    }

    pub fn append_tagged(&mut self, rows: &[f32], ids: &[u32], tags: &[u32]) -> Result<()> {
        if rows.len() != ids.len() * self.dim || tags.len() != ids.len() {
            return Err(anyhow!("Embedding dimension mismatch: expected {}, got {}", self.dim, rows.len() / ids.len().max(1)));
        }
        check(unsafe { csgpu_append_tagged(self.raw, rows.as_ptr(), ids.as_ptr(), tags.as_ptr(), ids.len() as u64) })
    }

    /// Returns how many live rows were tombstoned.
    pub fn remove(&mut self, ids: &[u32]) -> Result<usize> {
        let mut removed = 0u64;
        check(unsafe { csgpu_remove(self.raw, ids.as_ptr(), ids.len() as u64, &mut removed) })?;
        Ok(removed as usize)
    }

    pub fn build(&mut self) -> Result<()> {
        check(unsafe { csgpu_build(self.raw) })
    }

    pub fn clear(&mut self) -> Result<()> {
        check(unsafe { csgpu_clear(self.raw) })
    }

    /// Sidecar snapshot `<db>/gpu/` of a BUILT index (written at the end of build_index, loaded by new/open_readonly).
    pub fn save(&self, dir: &Path) -> Result<()> {
        let d = c_path(dir)?;
        check(unsafe { csgpu_save(self.raw, d.as_ptr()) })
    }

    pub fn load(&mut self, dir: &Path) -> Result<()> {
        let d = c_path(dir)?;
        check(unsafe { csgpu_load(self.raw, d.as_ptr()) })
    }

    // ---- search (store.rs:431-486; the arroy block :446-459) ----------------------------------------------------------
    fn collect(ids: Vec<u32>, dist: Vec<f32>, n: u32) -> Vec<(u32, f32)> {
        ids.into_iter().zip(dist).take(n as usize).collect()
    }

    /// `(chunk id, distance)` ascending by (distance, id); distance = (1 - cos) / 2, what arroy's Cosine returns.
    pub fn search(&self, q: &[f32], limit: usize) -> Result<Vec<(u32, f32)>> {
        let cap = limit.max(1);
        let (mut ids, mut dist, mut n) = (vec![0u32; cap], vec![0f32; cap], 0u32);
        check(unsafe {
            csgpu_search(self.raw, q.as_ptr(), q.len() as u32, limit as u32, ids.as_mut_ptr(), dist.as_mut_ptr(), &mut n)
        })?;
        Ok(Self::collect(ids, dist, n))
    }

    pub fn search_filtered(&self, q: &[f32], limit: usize, filter: &RowFilter) -> Result<Vec<(u32, f32)>> {
        let cap = limit.max(1);
        let (mut ids, mut dist, mut n) = (vec![0u32; cap], vec![0f32; cap], 0u32);
        check(unsafe {
            csgpu_search_filtered(
                self.raw, q.as_ptr(), q.len() as u32, limit as u32, filter.bitmap.as_ptr(), filter.n_bits,
                ids.as_mut_ptr(), dist.as_mut_ptr(), &mut n,
            )
        })?;
        Ok(Self::collect(ids, dist, n))
    }

    /// `file_bitmap`: optional per-FILE allow bitmap (bit file_id), see `CsgpuPredicate`.
    pub fn search_tagged(&self, q: &[f32], limit: usize, lang_mask: u32, file_lo: u32, file_hi: u32, file_bitmap: Option<&[u64]>,
                         n_file_bits: u64) -> Result<Vec<(u32, f32)>> {
        let pred = CsgpuPredicate {
            lang_mask, file_lo, file_hi, reserved: 0,
            file_bitmap: file_bitmap.map_or(std::ptr::null(), |b| b.as_ptr()),
            n_file_bits: if file_bitmap.is_some() { n_file_bits } else { 0 },
        };
        let cap = limit.max(1);
        let (mut ids, mut dist, mut n) = (vec![0u32; cap], vec![0f32; cap], 0u32);
        check(unsafe {
            csgpu_search_tagged(self.raw, q.as_ptr(), q.len() as u32, limit as u32, &pred, ids.as_mut_ptr(), dist.as_mut_ptr(), &mut n)
        })?;
        Ok(Self::collect(ids, dist, n))
    }

    /// The <= 9 query variants of src/search/mod.rs:508-511 in one call; one result list per query.
    pub fn search_batch(&self, queries: &[&[f32]], limit: usize) -> Result<Vec<Vec<(u32, f32)>>> {
        let b = queries.len();
        let mut flat = Vec::with_capacity(b * self.dim);
        for q in queries {
            if q.len() != self.dim {
                return Err(anyhow!("Query embedding dimension mismatch: expected {}, got {}", self.dim, q.len()));
            }
            flat.extend_from_slice(q);
        }
        let cap = limit.max(1);
        let (mut ids, mut dist, mut n) = (vec![0u32; b * cap], vec![0f32; b * cap], vec![0u32; b]);
        check(unsafe {
            csgpu_search_batch(self.raw, flat.as_ptr(), self.dim as u32, b as u32, limit as u32, ids.as_mut_ptr(), dist.as_mut_ptr(), n.as_mut_ptr())
        })?;
        Ok((0..b)
            .map(|j| (0..n[j] as usize).map(|i| (ids[j * limit + i], dist[j * limit + i])).collect())
            .collect())
    }

    /// Variants of ONE user query searched and deduplicated on the device (best distance per id, then the best `limit`):
    /// replaces the par_iter + HashMap/BinaryHeap pass at src/search/mod.rs:508-590.
    pub fn search_variants(&self, queries: &[&[f32]], limit: usize) -> Result<Vec<(u32, f32)>> {
        let b = queries.len();
        let mut flat = Vec::with_capacity(b * self.dim);
        for q in queries {
            flat.extend_from_slice(q);
        }
        let cap = limit.max(1);
        let (mut ids, mut dist, mut n) = (vec![0u32; cap], vec![0f32; cap], 0u32);
        check(unsafe {
            csgpu_search_variants(self.raw, flat.as_ptr(), self.dim as u32, b as u32, limit as u32, ids.as_mut_ptr(), dist.as_mut_ptr(), &mut n)
        })?;
        Ok(Self::collect(ids, dist, n))
    }

    /// `search_variants` under a row-tag predicate: hybrid search with a language / path filter as ONE call
    /// (src/search/mod.rs:508-590 + the post-filters at :727-737); `limit` results survive the filter.
    pub fn search_variants_tagged(&self, queries: &[&[f32]], limit: usize, lang_mask: u32, file_lo: u32, file_hi: u32,
                                  file_bitmap: Option<&[u64]>, n_file_bits: u64) -> Result<Vec<(u32, f32)>> {
        let b = queries.len();
        let mut flat = Vec::with_capacity(b * self.dim);
        for q in queries {
            flat.extend_from_slice(q);
        }
        let pred = CsgpuPredicate {
            lang_mask, file_lo, file_hi, reserved: 0,
            file_bitmap: file_bitmap.map_or(std::ptr::null(), |m| m.as_ptr()),
            n_file_bits: if file_bitmap.is_some() { n_file_bits } else { 0 },
        };
        let cap = limit.max(1);
        let (mut ids, mut dist, mut n) = (vec![0u32; cap], vec![0f32; cap], 0u32);
        check(unsafe {
            csgpu_search_variants_tagged(self.raw, flat.as_ptr(), self.dim as u32, b as u32, limit as u32, &pred,
                                         ids.as_mut_ptr(), dist.as_mut_ptr(), &mut n)
        })?;
        Ok(Self::collect(ids, dist, n))
    }

    // ---- opt-in routes (results never change) ---------------------------------------------------------------------------
    pub fn set_coalescing(&mut self, enabled: bool, window_us: u32) -> Result<()> {
        check(unsafe { csgpu_set_coalescing(self.raw, enabled as u32, window_us) })
    }

    pub fn set_tensor_prefilter(&mut self, enabled: bool) -> Result<()> {
        check(unsafe { csgpu_set_tensor_prefilter(self.raw, enabled as u32) })
    }

    pub fn set_byte_prefilter(&mut self, enabled: bool) -> Result<()> {
        check(unsafe { csgpu_set_byte_prefilter(self.raw, enabled as u32) })
    }

    // ---- introspection ---------------------------------------------------------------------------------------------------
    pub fn stats(&self) -> Result<CsgpuStats> {
        let mut s = CsgpuStats::default();
        check(unsafe { csgpu_stats(self.raw, &mut s) })?;
        Ok(s)
    }

    pub fn is_built(&self) -> bool {
        self.stats().map(|s| s.built != 0).unwrap_or(false)
    }

    pub fn get_tags(&self, ids: &[u32]) -> Result<Vec<u32>> {
        let mut out = vec![CSGPU_TAG_NONE; ids.len()];
        check(unsafe { csgpu_get_tags(self.raw, ids.as_ptr(), ids.len() as u64, out.as_mut_ptr()) })?;
        Ok(out)
    }

    pub fn kernel_launches() -> u64 {
        unsafe { csgpu_kernel_launches() }
    }

    // ---- synthetic corpus (benchmarks / tests) -----------------------------------------------------------------------------
    pub fn append_synthetic(&mut self, seed: u64, first_row: u64, n: u64, id_base: u32, tagged: bool) -> Result<()> {
        check(unsafe {
            if tagged { csgpu_append_synthetic_tagged(self.raw, seed, first_row, n, id_base) } else { csgpu_append_synthetic(self.raw, seed, first_row, n, id_base) }
        })
    }

    pub fn synth_rows_host(&self, seed: u64, first_row: u64, n: u64) -> Result<Vec<f32>> {
        let mut out = vec![0f32; n as usize * self.dim];
        check(unsafe { csgpu_synth_rows_host(self.raw, seed, first_row, n, out.as_mut_ptr()) })?;
        Ok(out)
    }

    // ---- rank-per-GPU deployment: device-resident entry points + the fused cross-GPU exchange ---------------------------
    // All pointers below are DEVICE pointers on this index's device and `stream` is a cudaStream_t; calls only enqueue.

    /// 64-byte handle to all-gather across the ranks (any transport), then `exchange_connect`.
    pub fn exchange_create(&mut self, world: u32, rank: u32) -> Result<[u8; CSGPU_EXCHANGE_HANDLE_BYTES]> {
        let mut h = [0u8; CSGPU_EXCHANGE_HANDLE_BYTES];
        check(unsafe { csgpu_exchange_create(self.raw, world, rank, h.as_mut_ptr() as *mut c_void) })?;
        Ok(h)
    }

    /// `handles`: world x 64 bytes, rank order (own entry ignored).
    pub fn exchange_connect(&mut self, handles: &[u8]) -> Result<()> {
        check(unsafe { csgpu_exchange_connect(self.raw, handles.as_ptr() as *const c_void) })
    }

    /// Indexes living in THIS process (several GPUs, or tests): `peers[p]` = rank p's index.
    pub fn exchange_connect_local(&mut self, peers: &[&GpuIndex]) -> Result<()> {
        let raw: Vec<*mut CsgpuIndex> = peers.iter().map(|p| p.raw).collect();
        check(unsafe { csgpu_exchange_connect_local(self.raw, raw.as_ptr()) })
    }

    pub fn exchange_set_timeout_ms(&mut self, ms: u32) -> Result<()> {
        check(unsafe { csgpu_exchange_set_timeout_ms(self.raw, ms) })
    }

    /// `true` after an in-kernel wait for a peer timed out (every later exchange search returns CSGPU_ERR_NCCL).
    pub fn exchange_timed_out(&self) -> Result<bool> {
        let mut t = 0u32;
        check(unsafe { csgpu_exchange_status(self.raw, &mut t) })?;
        Ok(t != 0)
    }

    /// Per-query, per-peer wait of this rank's exchange tail in ns (skew diagnostic), oldest first; returns (n_queries, data).
    pub fn exchange_wait_stats(&self, world: usize, max_queries: u32) -> Result<(u32, Vec<u64>)> {
        let mut out = vec![0u64; max_queries as usize * world];
        let mut n = 0u32;
        check(unsafe { csgpu_exchange_wait_stats(self.raw, out.as_mut_ptr(), max_queries, &mut n) })?;
        Ok((n, out))
    }

    pub fn exchange_destroy(&mut self) {
        unsafe { csgpu_exchange_destroy(self.raw) }
    }

    /// Host-pointer exchange search (rank-per-GPU processes): the GLOBAL top-k over all ranks, one fused launch per rank.
    /// A collective: every rank calls it for every query, in the same order.
    pub fn search_exchange(&self, q: &[f32], limit: usize) -> Result<Vec<(u32, f32)>> {
        let cap = limit.max(1);
        let (mut ids, mut dist, mut n) = (vec![0u32; cap], vec![0f32; cap], 0u32);
        check(unsafe {
            csgpu_search_exchange(self.raw, q.as_ptr(), q.len() as u32, limit as u32, ids.as_mut_ptr(), dist.as_mut_ptr(), &mut n)
        })?;
        Ok(Self::collect(ids, dist, n))
    }

    /// # Safety
    /// `q_dev` ([dim_pad] f32) and `out_keys_dev` ([k] u64) must be valid device pointers until `stream` drains.
    pub unsafe fn search_keys_device(&self, q_dev: *const f32, k: u32, out_keys_dev: *mut u64, stream: *mut c_void) -> Result<()> {
        check(csgpu_search_keys_device(self.raw, q_dev, k, out_keys_dev, stream))
    }

    /// ONE kernel per rank: local scan + peer stores over NVLink + flag wait + global merge.
    /// # Safety
    /// As `search_keys_device`; every rank must issue the same sequence of exchange searches.
    pub unsafe fn search_keys_exchange_device(&self, q_dev: *const f32, k: u32, out_keys_dev: *mut u64, stream: *mut c_void) -> Result<()> {
        check(csgpu_search_keys_exchange_device(self.raw, q_dev, k, out_keys_dev, stream))
    }

    /// # Safety
    /// As above; `pred.file_bitmap`, if set, is a DEVICE pointer valid until `stream` drains.
    pub unsafe fn search_tagged_keys_device(&self, q_dev: *const f32, k: u32, pred: &CsgpuPredicate, exchange: bool,
                                            out_keys_dev: *mut u64, stream: *mut c_void) -> Result<()> {
        check(csgpu_search_tagged_keys_device(self.raw, q_dev, k, pred, exchange as u32, out_keys_dev, stream))
    }

    /// # Safety
    /// `keys_dev`: [n_lists][k] u64 on the device; `out_keys_dev`: [k].
    pub unsafe fn merge_keys_device(&self, keys_dev: *const u64, n_lists: u32, k: u32, out_keys_dev: *mut u64, stream: *mut c_void) -> Result<()> {
        check(csgpu_merge_keys_device(self.raw, keys_dev, n_lists, k, out_keys_dev, stream))
    }

    /// # Safety
    /// `keys_dev`: [n_lists][nq][k]; `out_keys_dev`: [nq][k].
    pub unsafe fn merge_keys_batch_device(&self, keys_dev: *const u64, n_lists: u32, nq: u32, k: u32, out_keys_dev: *mut u64,
                                          stream: *mut c_void) -> Result<()> {
        check(csgpu_merge_keys_batch_device(self.raw, keys_dev, n_lists, nq, k, out_keys_dev, stream))
    }
}

/// keys (host copy of a device result) -> (ids, distances).
pub fn decode_keys(keys: &[u64]) -> Vec<(u32, f32)> {
    let k = keys.len();
    let (mut ids, mut dist, mut n) = (vec![0u32; k], vec![0f32; k], 0u32);
    unsafe { csgpu_decode_keys(keys.as_ptr(), k as u32, ids.as_mut_ptr(), dist.as_mut_ptr(), &mut n) };
    ids.into_iter().zip(dist).take(n as usize).collect()
}

/// (ids, distances) -> k keys padded with empty slots (to all-gather per-rank batch results).
pub fn encode_keys(ids: &[u32], dist: &[f32], k: usize) -> Vec<u64> {
    let mut out = vec![u64::MAX; k];
    unsafe { csgpu_encode_keys(ids.as_ptr(), dist.as_ptr(), ids.len().min(dist.len()) as u32, k as u32, out.as_mut_ptr()) };
    out
}
