/*
 * csgpu.h — C ABI of the B200-native vector-retrieval hot path for codesearch.
 *
 * This is the drop-in boundary. The reference has no FFI today:
 * `VectorStore` (src/vectordb/store.rs:94-102) calls arroy directly and build.rs
 * compiles nothing (build.rs:1-48). A maintainer replaces the arroy block
 *     src/vectordb/store.rs:446-459   (read_txn / Reader::open / nns / search_k / by_vector)
 * with one csgpu_search*() call, feeds csgpu_append/csgpu_remove/csgpu_build from
 *     src/vectordb/store.rs:618-686   (insert_chunks_with_ids: writer.add_item(id, &embedding))
 *     src/vectordb/store.rs:548-610   (delete_chunks: writer.del_item(id))
 *     src/vectordb/store.rs:386-430   (build_index)
 *     src/vectordb/store.rs:690-706   (clear)
 * and keeps the metadata join (store.rs:464-483) and score = 1 - distance
 * (store.rs:478) on the host. INTEGRATION.md shows the Rust binding.
 *
 * Semantics common to every search entry point
 *   - out_dist[i] = (1 - cos(row, q)) / 2 in [0, 1]   (arroy 0.5.0 Cosine distance)
 *   - results ascending (distance, chunk id); ties broken by the smaller id
 *   - *out_n = min(k, live rows passing the filter)
 *   - the query is normalised inside; a zero-norm query or row has distance 0.0
 *   - rows/queries containing NaN/Inf: queries are rejected (CSGPU_ERR_ARG);
 *     such rows are dropped at csgpu_build and counted in csgpu_stats_t.nonfinite_rows
 *   - caller owns every out_* buffer; the library never hands out device memory
 *     through the host-pointer entry points
 *
 * Threading
 *   csgpu_search* on a built index are re-entrant from any number of host threads
 *   (`&self` in the reference; rayon par_iter at src/search/mod.rs:508-511).
 *   csgpu_append/remove/build/clear/destroy need exclusive access (`&mut self`).
 *
 * Errors: 0 = ok, otherwise a CSGPU_ERR_* code; text via csgpu_last_error()
 * (thread-local). Never aborts. There is NO CPU fallback: no usable device =>
 * csgpu_create fails with CSGPU_ERR_CUDA.
 */
#ifndef CSGPU_H
#define CSGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CSGPU_ABI_VERSION 8

enum {
    CSGPU_OK            = 0,
    CSGPU_ERR_DIM       = 1, /* "Query embedding dimension mismatch: expected {}, got {}"  store.rs:432-438 */
    CSGPU_ERR_NOT_BUILT = 2, /* "Index not built. Call build_index() after inserting chunks."  store.rs:440-444 */
    CSGPU_ERR_CUDA      = 3,
    CSGPU_ERR_NCCL      = 4,
    CSGPU_ERR_OOM       = 5,
    CSGPU_ERR_ARG       = 6
};

enum { CSGPU_DTYPE_F32 = 0, CSGPU_DTYPE_BF16 = 1 };

#define CSGPU_MAX_K   1024u
#define CSGPU_MAX_DIM 4096u

typedef struct csgpu_index csgpu_index;

typedef struct csgpu_stats_t {
    uint64_t live_rows;        /* rows searchable after the last build (all devices)           */
    uint64_t pending_rows;     /* appended since the last build                                */
    uint64_t tombstones;       /* removed since the last build                                 */
    uint64_t zero_norm_rows;   /* rows with |v| = 0: kept, distance 0.0 (arroy pn*qn == 0)     */
    uint64_t nonfinite_rows;   /* rows with NaN/Inf: dropped at build                          */
    uint64_t bytes_on_device;  /* total HBM held by the index (all devices)                    */
    uint32_t dim;
    uint32_t dtype;
    uint32_t n_devices;
    uint32_t built;            /* mirrors VectorStore::is_indexed()  store.rs:747              */
    float    last_search_us;   /* device time of the most recent search: CUDA events; a single-device fp32 scan stamps
                                * %globaltimer in the kernel instead (first CTA started -> result written)              */
    uint32_t abi_version;
    uint64_t rows_per_device[8];
    uint64_t coalesced_passes;   /* micro-batcher: corpus passes launched ...                       */
    uint64_t coalesced_queries;  /* ... for this many csgpu_search calls (csgpu_set_coalescing)     */
    uint64_t prefilter_rescored; /* tensor prefilter: fp32 rows rescored by the last batch (<= 1024 queries) */
    uint64_t shadow_bytes;       /* HBM held by the bf16 shadow (csgpu_set_tensor_prefilter), all devices   */
    uint64_t byte_shadow_bytes;  /* HBM held by the int8 shadow + per-row bounds (csgpu_set_byte_prefilter)  */
    uint64_t byte_searches;      /* csgpu_search calls that streamed the int8 shadow ...                     */
    uint64_t byte_fallbacks;     /* ... of which this many were answered again by the fp32 scan kernel       */
    uint64_t byte_candidates;    /* last such search: rows that reached the final candidate list             */
    uint64_t byte_rescored;      /* last such search: fp32 rows read for the exact rescoring                 */
    uint32_t batch_route;        /* contraction of the last GEMM-shaped batch: CSGPU_ROUTE_* (0 until one ran)          */
    float    filter_max_err;     /* largest |d_filter - d_f32| the rescoring of that batch saw (tensor-core filters)     */
} csgpu_stats_t;

/* csgpu_stats_t.batch_route */
enum { CSGPU_ROUTE_NONE = 0, CSGPU_ROUTE_SIMT_F32 = 1, CSGPU_ROUTE_TC_BF16 = 2, CSGPU_ROUTE_TC_TF32 = 3 };

/* ---- lifecycle (VectorStore::new / open_readonly  store.rs:110-176,183-250) ------------ */

/* devices == NULL => device 0. n_devices in [1, 8]: rows are sharded row-wise over the devices of this ONE process
 * (the reference is one process: rayon inside it, src/search/mod.rs:508-511; one RwLock, src/server/mod.rs:27).
 * csgpu_search / csgpu_search_filtered / csgpu_search_tagged on such an index are n concurrent scan launches whose tails
 * push their k keys into device 0's HBM over NVLink; device 0's last CTA merges and writes the global top-k into mapped
 * host memory (no peer memcpy, no merge launch, one stream synchronised). A shard that never delivers makes the call
 * fail with CSGPU_ERR_NCCL after the exchange timeout (csgpu_exchange_set_timeout_ms). */
int  csgpu_create(csgpu_index **out, uint32_t dim, uint32_t dtype,
                  const int32_t *devices, uint32_t n_devices);
void csgpu_destroy(csgpu_index *ix);

/* rows: [n, dim] row-major host floats; ids: [n] chunk ids. Copies; marks the index dirty
 * (store.rs:674,682: add_item + indexed=false). Appending an id that is live replaces it
 * at the next build (LMDB put semantics). */
int  csgpu_append(csgpu_index *ix, const float *rows, const uint32_t *ids, uint64_t n);

/* Tombstones ids; marks the index dirty (store.rs:595,605-607). n_removed may be NULL. */
int  csgpu_remove(csgpu_index *ix, const uint32_t *ids, uint64_t n, uint64_t *n_removed);

/* Pre-sizes device storage for total_rows rows (avoids grow-and-copy; needed when a shard is
 * more than half of HBM). Optional. */
int  csgpu_reserve(csgpu_index *ix, uint64_t total_rows);

/* Normalises new rows to unit length, applies tombstones/replacements and compacts every shard in place (rows never
 * move between shards: large appends are split evenly, small ones go to the least-loaded shard); the index becomes
 * searchable (store.rs:422-430: indexed=true). */
int  csgpu_build(csgpu_index *ix);

/* Drops every row; index is empty and NOT built (store.rs:690-706). */
int  csgpu_clear(csgpu_index *ix);

/* ---- snapshot / hydrate: HBM is volatile, LMDB is the reference's truth (store.rs:110-176: `new` reopens the
 *      environment and probes Reader::open to learn `indexed`). csgpu_save writes a flat sidecar of a BUILT index,
 *      <dir>/{meta.json, ids.u32, rows.f32 | rows.bf16, zero.u32} (rows exactly as they sit in HBM, checksummed,
 *      published by rename with meta.json last); csgpu_load fills an EMPTY index of the same dim/dtype from it and
 *      leaves it built: searches return bit-identical results to the index that was saved. A maintainer calls
 *      csgpu_save at the end of build_index (store.rs:422-430) and csgpu_load in new/open_readonly. */
int  csgpu_save(const csgpu_index *ix, const char *dir);
int  csgpu_load(csgpu_index *ix, const char *dir);

/* ---- search (VectorStore::search  store.rs:431-486, arroy block :446-459) -------------- */

int  csgpu_search(const csgpu_index *ix, const float *q, uint32_t q_len, uint32_t k,
                  uint32_t *out_ids /*[k]*/, float *out_dist /*[k]*/, uint32_t *out_n);

/* Host micro-batcher (off by default). When enabled, concurrent csgpu_search calls with the same k are coalesced
 * into one csgpu_search_batch over the corpus (groups of up to 16 callers; up to 128 where the group runs as a tensor-core
 * batch, see csgpu_search_batch) — the caller pattern of src/search/mod.rs:508-511
 * (rayon par_iter over <= 9 query variants) and of concurrent MCP/HTTP readers. Results are bit-identical to the
 * uncoalesced call. window_us > 0 lets a pass linger that long for company before launching (0 = never wait:
 * only requests that arrive while a pass is in flight get batched). */
int  csgpu_set_coalescing(csgpu_index *ix, uint32_t enabled, uint32_t window_us);

/* b queries [b, dim]; outputs [b, k] (row j holds out_n[j] valid entries). Serves the
 * <= 9 query variants of src/search/mod.rs:508-511 in one pass over the corpus, and batches of any size.
 * fp32 index: every list is what csgpu_search returns for that query — ids AND distances bit for bit — on every route
 * but the SIMT one. Routes, chosen by a measured cost model (csrc/csgpu.cu tf32_route_is_faster):
 *   - a few queries over a small corpus: one multi-query scan launch (<= 16 queries per pass, scan_multi.cuh);
 *   - otherwise (from 2 queries at 10M rows, ~17 at 100k rows): the tensor cores straight off the fp32 rows —
 *     tcgen05.mma kind::tf32 as a FILTER with a proven margin (|d_tf32 - d_f32| < 1.1e-3 for dim <= 1024), survivors
 *     rescored with the single-query kernel's arithmetic (gemm_tf32.cuh, rescore.cuh). No shadow copy, no opt-in;
 *   - dim > 1024 (or CSGPU_BATCH_SIMT=1): the register-tiled fp32 SIMT kernel (gemm_simt.cuh; same ranking, distances may
 *     differ from csgpu_search in the last ulp).
 * csgpu_stats_t.batch_route reports the contraction the last GEMM-shaped batch used. */
int  csgpu_search_batch(const csgpu_index *ix, const float *q, uint32_t q_len, uint32_t b, uint32_t k,
                        uint32_t *out_ids, float *out_dist, uint32_t *out_n /*[b]*/);

/* Opt-in, fp32 index only: exact fp32 results at tensor-core speed for batches. The index keeps a bf16 SHADOW of its
 * unit rows (+50 % HBM; dim % 64 == 0, dim <= 512). csgpu_search_batch then contracts the batch on the tensor cores
 * (tcgen05 / TMEM) against the shadow as a FILTER with a proven error margin, and rescores the survivors from the fp32
 * rows with the single-query kernel's arithmetic: ids AND distances are bit-identical to csgpu_search on every query
 * (codesearch_b200/csrc/rescore.cuh has the bound). Off by default: the default batched path needs no shadow (tf32 off
 * the fp32 rows, see csgpu_search_batch); this switch buys the bf16 rate of the tensor pipe for large batches (1024 x
 * top-100 over 10M x 384: ~8 ms instead of 12). May be called before or after csgpu_build; the shadow follows every later
 * build / load. */
int  csgpu_set_tensor_prefilter(csgpu_index *ix, uint32_t enabled);

/* Opt-in, fp32 index only: exact fp32 results for csgpu_search from ONE pass over a 1-byte-per-element shadow of
 * the corpus (+25 % HBM; dim <= 1024) instead of the 4-byte rows. The shadow (per-row scaled int8 + a per-row error
 * bound computed at build) is streamed as a FILTER with a proven bound (integer dot products, Cauchy-Schwarz on the
 * two quantisation residuals); the few hundred rows it cannot exclude are rescored from the fp32 rows with the
 * single-query kernel's arithmetic, in the same kernel launch. Ids AND distances are bit-identical to the default path
 * (codesearch_b200/csrc/scan_i8.cuh has the bound); the kernel moves dim + 4 bytes per row, not 4 * dim. Applies to
 * csgpu_search, csgpu_search_filtered and csgpu_search_tagged (and the device entry points) with k <= 256 on shards of
 * >= 524288 rows — under a filter only row groups the filter allows are streamed, so masked rows still cost nothing;
 * everything else, and any query the filter cannot bound (zero-norm query, candidate overflow), runs the fp32 scan
 * kernels. Off by default. May be called before or after csgpu_build; the shadow follows every later build / load. */
int  csgpu_set_byte_prefilter(csgpu_index *ix, uint32_t enabled);

/* b (<= 16) query VARIANTS of one user query (query expansion, src/search/mod.rs:479-483), searched with the same
 * limit and merged on the device: per chunk id the best (smallest) distance over the variants, then the best k of
 * the union, ascending (distance, id). Replaces the par_iter of searches plus the HashMap/BinaryHeap dedup at
 * src/search/mod.rs:508-590 with ceil(b/8) corpus passes and one merge kernel; outputs [k]. */
int  csgpu_search_variants(const csgpu_index *ix, const float *q /*[b, dim]*/, uint32_t q_len, uint32_t b, uint32_t k,
                           uint32_t *out_ids, float *out_dist, uint32_t *out_n);

/* Exact top-k over rows whose chunk id has its bit set (bit i of id_bitmap[i/64]); ids >=
 * n_bits are excluded. New capability: the reference only post-filters on the host
 * (src/search/mod.rs:727-737, src/server/mod.rs:553-559). */
int  csgpu_search_filtered(const csgpu_index *ix, const float *q, uint32_t q_len, uint32_t k,
                           const uint64_t *id_bitmap, uint64_t n_bits,
                           uint32_t *out_ids, float *out_dist, uint32_t *out_n);

/* ---- row-tag filter columns: every row carries ONE packed u32 tag = (lang_id << 27) | file_id
 *      in HBM next to its chunk id, so a file/language filter is a PREDICATE evaluated on the device before a row is
 *      read — the host ships at most a per-FILE bitmap (~37x smaller than a per-chunk one; chunker: ~37 chunks per
 *      file) instead of walking ChunkMetadata to build an N-bit map per query. lang_id follows the declaration order
 *      of `Language` (src/file/language.rs:5-29: Rust = 0 ... Unknown = 22; detection src/file/language.rs:31-88);
 *      file_id is a dense per-file number the host assigns (FileMetaStore, src/cache/file_meta.rs). The reference reads
 *      `primary_language` but never writes it (src/search/mod.rs:120-124 vs src/index/mod.rs:882-887) and only
 *      post-filters paths on the host (src/search/mod.rs:727-737, src/server/mod.rs:553-559); this is the new
 *      capability that replaces both. Rows appended with csgpu_append carry CSGPU_TAG_NONE. Tags follow their rows
 *      through csgpu_build / csgpu_remove / csgpu_save / csgpu_load. */
#define CSGPU_TAG_LANG_SHIFT 27u
#define CSGPU_TAG_FILE_MASK  0x07FFFFFFu
#define CSGPU_TAG_NONE       0xFFFFFFFFu   /* lang 31 (never a Language variant), file 0x7FFFFFF */
#define CSGPU_TAG(lang, file) ((((uint32_t)(lang) & 31u) << CSGPU_TAG_LANG_SHIFT) | ((uint32_t)(file) & CSGPU_TAG_FILE_MASK))

/* A row passes iff  (lang_mask >> lang_id) & 1   AND   file_lo <= file_id <= file_hi
 *                   AND (file_bitmap == NULL  OR  (file_id < n_file_bits AND bit file_id of file_bitmap is set)).
 * lang_mask = 0xFFFFFFFF, file_lo = 0, file_hi = 0xFFFFFFFF, file_bitmap = NULL passes every row (including
 * untagged ones). A path-prefix filter is a file range when the host numbers files in path order, a file bitmap
 * otherwise. file_bitmap is a HOST pointer for csgpu_search_tagged and a DEVICE pointer for the *_device call. */
typedef struct csgpu_predicate_t {
    uint32_t lang_mask;
    uint32_t file_lo, file_hi;
    uint32_t reserved;
    const uint64_t *file_bitmap;
    uint64_t n_file_bits;
} csgpu_predicate_t;

/* csgpu_append + one tag per row (CSGPU_TAG(lang_id, file_id)). */
int  csgpu_append_tagged(csgpu_index *ix, const float *rows, const uint32_t *ids, const uint32_t *tags, uint64_t n);

/* Exact top-k over the rows that pass `pred` (same semantics as csgpu_search_filtered with the equivalent id
 * bitmap; rows that fail are never read from HBM). */
int  csgpu_search_tagged(const csgpu_index *ix, const float *q, uint32_t q_len, uint32_t k,
                         const csgpu_predicate_t *pred,
                         uint32_t *out_ids, float *out_dist, uint32_t *out_n);

/* csgpu_search_variants under a row-tag predicate: b (<= 16) query variants of one user query, each searched over the rows
 * that pass `pred`, merged on the device (best distance per chunk id, then the best k, ascending (distance, id)); outputs [k].
 * The reference's hybrid search with a language / path filter (src/search/mod.rs:508-590 followed by the post-filters at
 * :727-737) as ONE call — and, unlike a post-filter, `k` results survive. Equals the dedup of b csgpu_search_tagged lists
 * bit for bit. Small corpora (<= 768 MB of rows per device) take multi-query passes with the predicate applied to the
 * results that beat a threshold (one launch per <= 16 variants); larger ones one filtered scan per variant (masked rows
 * are never read); the lists stay on the device either way. file_bitmap is a HOST pointer. */
int  csgpu_search_variants_tagged(const csgpu_index *ix, const float *q /*[b, dim]*/, uint32_t q_len, uint32_t b, uint32_t k,
                                  const csgpu_predicate_t *pred,
                                  uint32_t *out_ids /*[k]*/, float *out_dist /*[k]*/, uint32_t *out_n);

/* Reads back the tags of `n` chunk ids (CSGPU_TAG_NONE for ids that are not live). Test/introspection helper. */
int  csgpu_get_tags(const csgpu_index *ix, const uint32_t *ids, uint64_t n, uint32_t *out_tags);

/* ---- device-resident entry points (single-device index; for hosts that already hold the
 *      query on the GPU, and for rank-per-GPU sharding under torch.distributed) -----------
 * All pointers are DEVICE pointers on the index's device; `stream` is a cudaStream_t (NULL
 * = default stream). Calls only enqueue work; nothing is synchronised.
 *
 * A "key" is the sortable 64-bit form of one result:  (okey(distance) << 32) | chunk id,
 * okey = order-preserving map of the f32 bits, so ascending u64 == ascending (distance, id).
 * 0xFFFFFFFFFFFFFFFF is the empty slot. csgpu_decode_keys() converts on the host. */

/* One query -> this index's local top-k keys, sorted ascending, padded with empty slots. */
int  csgpu_search_keys_device(const csgpu_index *ix, const float *q_dev, uint32_t k,
                              uint64_t *out_keys_dev /*[k]*/, void *stream);

/* k-way merge of n_lists lists of k keys (e.g. the all-gathered per-rank results) into the
 * global top-k, ties by chunk id. keys_dev: [n_lists, k]. */
int  csgpu_merge_keys_device(const csgpu_index *ix, const uint64_t *keys_dev, uint32_t n_lists,
                             uint32_t k, uint64_t *out_keys_dev /*[k]*/, void *stream);

/* Batched form: keys_dev [n_lists][nq][k] (e.g. the all-gathered per-rank results of csgpu_search_batch, re-encoded
 * with csgpu_encode_keys) -> out_keys_dev [nq][k], one CTA per query. */
int  csgpu_merge_keys_batch_device(const csgpu_index *ix, const uint64_t *keys_dev, uint32_t n_lists, uint32_t nq,
                                   uint32_t k, uint64_t *out_keys_dev /*[nq][k]*/, void *stream);

/* ---- fused cross-GPU exchange: rank-per-GPU sharding without a collective library on the data path.
 * Each rank (one process per GPU) owns a slot block in its HBM that its peers write over NVLink. The scan
 * kernel's last CTA stores the rank's k local keys straight into every peer's block, raises a sequence
 * flag, waits for the peers' flags and merges: ONE kernel per query returns the GLOBAL top-k on every
 * rank (replaces: scan kernel + NCCL all-gather + merge kernel). Every rank must issue the same sequence
 * of csgpu_search_keys_exchange_device calls (same k), like any collective.
 *   1. csgpu_exchange_create on every rank -> 64-byte handle
 *   2. all-gather the handles (torch.distributed / MPI / a file: host side, once)
 *   3. csgpu_exchange_connect with all world handles (cudaIpcOpenMemHandle on the peers' blocks)
 * csgpu_exchange_connect_local wires indexes that live in ONE process (several GPUs, or tests).
 * A peer that never arrives makes the wait time out (4 s by default; csgpu_exchange_set_timeout_ms, or the environment
 * variable CSGPU_EXCHANGE_TIMEOUT_MS): that query's keys are undefined, csgpu_exchange_status reports 1, and EVERY later
 * exchange search on this rank returns CSGPU_ERR_NCCL until csgpu_exchange_create/connect are run again. The status
 * word lives in pinned host memory, so checking it never touches the device. */
#define CSGPU_EXCHANGE_HANDLE_BYTES 64
int  csgpu_exchange_create(csgpu_index *ix, uint32_t world, uint32_t rank, void *out_handle /*[64]*/);
int  csgpu_exchange_connect(csgpu_index *ix, const void *handles /*[world][64]; own entry ignored*/);
int  csgpu_exchange_connect_local(csgpu_index *ix, csgpu_index *const *peers /*[world]*/);
int  csgpu_search_keys_exchange_device(const csgpu_index *ix, const float *q_dev, uint32_t k,
                                       uint64_t *out_keys_dev /*[k] global top-k*/, void *stream);
/* The same search with HOST pointers (what csgpu_search is to a single index): pinned staging of the query, the one
 * fused launch, the global top-k written by its last CTA straight into mapped host memory, one stream synchronised.
 * Collective like the device form: every rank calls it for every query in the same order, one call at a time per rank. */
int  csgpu_search_exchange(const csgpu_index *ix, const float *q, uint32_t q_len, uint32_t k,
                           uint32_t *out_ids /*[k]*/, float *out_dist /*[k]*/, uint32_t *out_n);
int  csgpu_exchange_status(const csgpu_index *ix, uint32_t *timed_out);
/* Bound of the in-kernel wait for the peers' keys, for the rank-per-GPU exchange and for the in-process multi-device
 * index alike. Call it while no search is in flight. */
int  csgpu_exchange_set_timeout_ms(csgpu_index *ix, uint32_t ms);
/* Skew diagnostic: for the most recent n (<= 128, <= max_queries) exchange searches of this rank, oldest first,
 * out_ns[i * world + p] = nanoseconds between this rank publishing its keys and rank p's keys landing here (0 = they
 * were already there). Synchronise the search stream first. */
int  csgpu_exchange_wait_stats(const csgpu_index *ix, uint64_t *out_ns /*[max_queries][world]*/, uint32_t max_queries,
                               uint32_t *n_queries);
void csgpu_exchange_destroy(csgpu_index *ix);

/* Device-resident predicate search: local top-k (exchange = 0) or, with a connected exchange, the GLOBAL top-k over
 * all ranks in the same kernel (exchange = 1; BASELINE configs[4]: filter mask + 8 GPUs). pred->file_bitmap is a
 * device pointer that must stay valid until `stream` drains. */
int  csgpu_search_tagged_keys_device(const csgpu_index *ix, const float *q_dev, uint32_t k,
                                     const csgpu_predicate_t *pred, uint32_t exchange,
                                     uint64_t *out_keys_dev /*[k]*/, void *stream);

/* Host-side inverse of csgpu_decode_keys: n (<= k) results -> k keys, padded with empty slots. */
void csgpu_encode_keys(const uint32_t *ids, const float *dist, uint32_t n, uint32_t k, uint64_t *out_keys);

/* Host-side: keys -> (ids, distances); returns the number of non-empty slots in *out_n. */
void csgpu_decode_keys(const uint64_t *keys, uint32_t k, uint32_t *out_ids, float *out_dist,
                       uint32_t *out_n);

/* ---- synthetic corpus (bench/tests): rows generated ON the device, never on the host --- */
/* Appends rows [first_row, first_row+n) of the counter-based generator (Philox4x32-10,
 * spec in codesearch_b200/csrc/synth.cuh), chunk id = (uint32_t)global row index + id_base. */
int  csgpu_append_synthetic(csgpu_index *ix, uint64_t seed, uint64_t first_row, uint64_t n,
                            uint32_t id_base);
/* Same rows, plus synthetic tags: file_id = global_row / 37, lang_id = fmix32(file_id) % 23
 * (fmix32 = MurmurHash3's 32-bit finaliser; 23 = `Language` variants, src/file/language.rs:5-29). */
int  csgpu_append_synthetic_tagged(csgpu_index *ix, uint64_t seed, uint64_t first_row, uint64_t n,
                                   uint32_t id_base);
/* The same generator into a host buffer via the device (for cross-checking the generators). */
int  csgpu_synth_rows_host(const csgpu_index *ix, uint64_t seed, uint64_t first_row, uint64_t n,
                           float *out_rows /*[n, dim] host*/);

/* ---- introspection --------------------------------------------------------------------- */
int  csgpu_stats(const csgpu_index *ix, csgpu_stats_t *out);
/* Number of kernels this library has launched in this process (all indexes, all threads). */
uint64_t csgpu_kernel_launches(void);
const char *csgpu_last_error(void);   /* thread-local; valid until the next call on this thread */
uint32_t csgpu_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* CSGPU_H */
