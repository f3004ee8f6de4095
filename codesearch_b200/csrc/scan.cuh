// scan.cuh — single-query exact cosine scan fused with top-k (the headline kernel).
//
// Replaces the arroy block at /root/reference/src/vectordb/store.rs:446-459 with the exact
// ranking of /root/reference/examples/benchmark_models.rs:155-165,323-328 generalised to top-k.
//
// HBM layout: rows are row-major fp32 UNIT vectors, dim padded to a multiple of 4 floats, so a
// row is `dim4` float4 (1536 B at D=384) and every warp-level load is a run of full 128 B lines.
// ids[row] is the chunk id of that row (u32).
//
// Roofline: HBM-bound. Algorithmic bytes per query = n_rows * dim * 4. Per row a
// warp issues V LDG.128, 4V FFMA, 5 SHFL+FADD and one compare against a warp-uniform threshold
// (~30 instructions per 1536 B), far below issue limits; what matters is bytes in flight:
// R rows x V x 512 B per warp x 8 warps x 2 CTAs/SM ~= 96 KB/SM.
//
// Determinism: each row's dot product uses one fixed FMA chain per lane and one fixed xor-shuffle
// tree, independent of which warp/CTA/GPU scans the row. Duplicate rows therefore tie bit-exactly
// and the id tie-break is well defined; a row-sharded multi-GPU search returns bit-identical
// results to a single-GPU one.
#pragma once
#include "topk.cuh"

namespace csgpu {

constexpr int SCAN_WARPS = 8;
constexpr int SCAN_THREADS = SCAN_WARPS * 32;
// dynamic shared memory of the k <= 32 scan kernels: [SCAN_WARPS * 32] keys for the merge trees + as many for the
// last CTA's list of survivors (cta_topk_of_lists32)
constexpr size_t SCAN_SMALL_SMEM = (size_t)2 * SCAN_WARPS * 32 * sizeof(uint64_t);

// Cross-GPU exchange fused into the scan kernel (rank-per-GPU sharding). Every rank owns a
// slot array + flags in its HBM that all peers can write over NVLink (cudaIpc-mapped, or plain pointers in
// one process). The last CTA of a rank's scan stores its k local keys straight into every peer's slot,
// publishes a sequence flag, waits for the peers' flags in its own memory and merges — no NCCL call, no
// separate merge launch. Slots/flags are double-buffered by sequence parity: a rank can be at most one
// query ahead of a peer, because finishing query s needs the peer's keys of query s.
constexpr int XCHG_MAX_WORLD = 8;
constexpr int XCHG_STATS_RING = 128;   // queries whose per-peer wait times are kept (diagnostic: cross-rank skew)
struct ExchangeDev {
    uint64_t *slots[XCHG_MAX_WORLD];   // slots[p]: rank p's array [2][world][kmax], as mapped on THIS device
    unsigned *flags[XCHG_MAX_WORLD];   // flags[p]: rank p's flags [2][world]
    unsigned *status;                  // pinned, device-mapped host word: set to 1 if a wait timed out (sticky; the host reads
                                       // it without touching the device and fails the next call with CSGPU_ERR_NCCL)
    unsigned long long *wait_ring;     // local [XCHG_STATS_RING][XCHG_MAX_WORLD]: ns between this rank's publish and the arrival
                                       // of peer p's keys, per query (ring by sequence number) — the per-step skew, measured
    unsigned long long timeout_ns;     // bound of the flag wait
    uint32_t world, rank, kmax;
    int32_t root;                      // < 0: all-to-all, every rank ends with the global top-k (rank-per-GPU processes);
                                       // >= 0: gather — ranks push their keys to `root` only and leave, `root` alone waits,
                                       // merges and writes the result (in-process multi-device index: one host reads one list)
};

struct ScanArgs {
    const float4 *rows;      // [n_rows, dim4]
    const uint32_t *ids;     // [n_rows]
    uint64_t n_rows;
    uint32_t dim4;
    const float *q;          // [dim4*4] device, raw (normalised in the prologue)
    uint32_t k, kpad;
    const uint64_t *bitmap;  // id-indexed allow bitmap or nullptr; with `tags` set: FILE-id-indexed bitmap or nullptr
    uint64_t n_bits;
    const uint32_t *tags = nullptr;   // [n_rows] packed row tags: non-null selects the predicate filter (csgpu_search_tagged)
    uint32_t lang_mask = 0xFFFFFFFFu, file_lo = 0, file_hi = 0xFFFFFFFFu;
    const uint32_t *zero_ids;  // ascending ids of zero-norm rows (distance 0.0), or nullptr; tags follow at [n_zero..2 n_zero)
    uint32_t n_zero;
    uint64_t *cand;          // [gridDim.x, k] per-CTA results
    unsigned *ticket;        // last-CTA-done counter (self-resetting)
    uint64_t *out_keys;      // [k] final keys, ascending, KEY_EMPTY padded
    const ExchangeDev *xchg = nullptr;  // non-null: out_keys receives the GLOBAL top-k over all ranks
    uint32_t seq = 0;                   // exchange sequence number of this query (same on every rank, >= 1)
    const unsigned *run_if = nullptr;   // non-null: the whole launch is a no-op unless *run_if != 0 (device-side fallback of
                                        // the byte prefilter: the host cannot look at the status without synchronising)
    uint32_t static_split = 0;          // filtered scans: 1 = fixed-stride block split instead of the work counter (CSGPU_SCAN_STATIC)
    // device time of the launch without CUDA events (two event records cost a small-corpus query 8 us of its ~50): CTA 0 stamps
    // %globaltimer into *t0_slot (device memory) when it starts, the CTA that writes the result stores now - t0 (ns) into
    // *elapsed_out (mapped host memory). Both null: no stamps.
    unsigned long long *t0_slot = nullptr, *elapsed_out = nullptr;
    unsigned long long *timing = nullptr;   // diagnostic (CSGPU_SCAN_TIMING=1): [gridDim.x + 1][4] globaltimer stamps — per CTA
                                            // start / streaming done / CTA list written; last CTA: ticket taken / merged / done
};

// LD selects the load flavour (tuned on B200, see profiles/): 0 = ld.global.nc.L1::no_allocate,
// 1 = plain ld.global.nc (__ldg), 2 = ld.global.cs (streaming / evict-first), 3-5 = flavour 0 with L2 hints.
template <int LD>
__device__ __forceinline__ float4 ldg_stream(const float4 *p)
{
    float4 v;
    if constexpr (LD == 0) {
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    } else if constexpr (LD == 1) {
        v = __ldg(p);
    } else if constexpr (LD == 2) {
        v = __ldcs(p);
    } else if constexpr (LD == 3) {   // + 256-byte L2 prefetch granularity
        asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    } else {   // 4: + evict-first L2 policy (the corpus is streamed once per query); 5: policy + 256-byte prefetch
        uint64_t pol;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        if constexpr (LD == 4)
            asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                         : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
        else
            asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.L2::256B.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                         : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
    }
    return v;
}

__device__ __forceinline__ bool id_allowed(const uint64_t *bitmap, uint64_t n_bits, uint32_t id)
{
    if (bitmap == nullptr) return true;
    if ((uint64_t)id >= n_bits) return false;
    return (bitmap[id >> 6] >> (id & 63)) & 1ull;
}

// Row-tag predicate (csgpu_predicate_t): language bit and file range from the tag alone; the optional per-file
// bitmap word is fetched separately (tag_word_needed) so its load can be pipelined like the id bitmap's.
__device__ __forceinline__ bool tag_pass_static(uint32_t tag, uint32_t lang_mask, uint32_t file_lo, uint32_t file_hi)
{
    const uint32_t lang = tag >> 27, file = tag & 0x07FFFFFFu;
    return ((lang_mask >> lang) & 1u) && file >= file_lo && file <= file_hi;
}
// zero-norm row i of the side list: allowed under whichever filter the launch carries
__device__ __forceinline__ bool zero_row_allowed(const uint64_t *bitmap, uint64_t n_bits, const uint32_t *tags_on,
                                                 uint32_t lang_mask, uint32_t file_lo, uint32_t file_hi,
                                                 const uint32_t *zero_ids, uint32_t n_zero, uint32_t i)
{
    if (tags_on == nullptr) return id_allowed(bitmap, n_bits, zero_ids[i]);
    const uint32_t tag = zero_ids[n_zero + i];
    if (!tag_pass_static(tag, lang_mask, file_lo, file_hi)) return false;
    return id_allowed(bitmap, n_bits, tag & 0x07FFFFFFu);
}

__device__ __forceinline__ float warp_sum_tree(float v)
{
    v += __shfl_xor_sync(FULL, v, 16);
    v += __shfl_xor_sync(FULL, v, 8);
    v += __shfl_xor_sync(FULL, v, 4);
    v += __shfl_xor_sync(FULL, v, 2);
    v += __shfl_xor_sync(FULL, v, 1);
    return v;
}

// Offer each lane's key (KEY_EMPTY = nothing) to the warp's selector.
template <class Sel>
__device__ __forceinline__ void offer_lane_keys(Sel &sel, uint64_t key, int lane)
{
    unsigned m = __ballot_sync(FULL, key < sel.thr);
    while (m) {
        int src = __ffs(m) - 1;
        m &= m - 1;
        uint64_t kk = shfl64(key, src);
        if (kk < sel.thr) sel.insert(kk, lane);
    }
}

// Warp selectors -> the CTA's top-k, ascending, into dst[0..k) (global or shared).
// smem: BIG: [2][SCAN_WARPS][kpad] (region A then B); !BIG: [SCAN_WARPS*32].
template <bool BIG, class Sel>
__device__ __forceinline__ void cta_reduce(Sel &sel, uint64_t *smem, uint32_t k, uint32_t kpad,
                                           uint64_t *dst, int warp, int lane)
{
    sel.flush(lane);
    uint32_t n;
    if constexpr (BIG) {
        uint64_t *mine = smem + (size_t)warp * kpad;
        if (sel.cur != mine)
            for (uint32_t j = lane; j < kpad; j += 32) mine[j] = sel.cur[j];
        n = SCAN_WARPS * kpad;
    } else {
        smem[warp * 32 + lane] = sel.v;
        n = SCAN_WARPS * 32;
    }
    cta_sort(smem, n);
    for (uint32_t j = threadIdx.x; j < k; j += blockDim.x) dst[j] = smem[j];
    __syncthreads();
}

// k <= 32, round 2. Every warp holds a sorted list of 32 keys (one per lane, ascending, KEY_EMPTY padded): the CTA's best
// 32 by a three-level tree of warp merges through shared memory (smem: [SCAN_WARPS * 32] keys) — five barriers and three
// 5-stage shuffle networks instead of the 36 barrier-separated stages of a 256-key bitonic sort (measured with
// CSGPU_SCAN_TIMING: 4.2 us -> the tail of EVERY query pays it twice). dst[0..k) and smem[0..32) receive the result.
__device__ __forceinline__ void cta_merge_lists32(uint64_t v, uint64_t *smem, uint32_t k, uint64_t *dst, int warp, int lane)
{
    smem[warp * 32 + lane] = v;
    __syncthreads();
#pragma unroll
    for (int half = SCAN_WARPS / 2; half >= 1; half >>= 1) {
        if (warp < half) v = warp_merge_low32(v, smem[(warp + half) * 32 + lane], lane);
        __syncthreads();                       // everyone has read its partner's list ...
        if (warp < half) smem[warp * 32 + lane] = v;
        __syncthreads();                       // ... before the survivors overwrite theirs
    }
    if (warp == 0 && dst != nullptr && (uint32_t)lane < k) dst[lane] = v;
    __syncthreads();
}

// This warp's share of n keys (key = load(i); batches of 32 dealt round-robin over the CTA's warps) folded into its
// running sorted list: each batch is sorted across the warp (15 shuffle stages) and merged in (5) — no data-dependent
// insertion loop. Eight batches per step: their loads are all in flight before the first one is used (one L2 round trip
// per 2048 keys of the CTA instead of one per batch) and the eight independent sorts overlap their shuffle latencies.
template <int NB = 8, class F>
__device__ __forceinline__ uint64_t warp_fold_keys32(uint64_t run, uint64_t n, F load, int warp, int lane)
{
    for (uint64_t b0 = (uint64_t)warp * 32; b0 < n; b0 += (uint64_t)NB * SCAN_WARPS * 32) {
        uint64_t kk[NB];
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            const uint64_t i = b0 + (uint64_t)u * SCAN_WARPS * 32 + lane;
            kk[u] = i < n ? load(i) : KEY_EMPTY;
        }
#pragma unroll
        for (int u = 0; u < NB; ++u) kk[u] = warp_sort32(kk[u], lane);
#pragma unroll
        for (int s = 1; s < NB; s <<= 1)
#pragma unroll
            for (int u = 0; u + s < NB; u += 2 * s) kk[u] = warp_merge_low32(kk[u], kk[u + s], lane);
        run = warp_merge_low32(run, kk[0], lane);
    }
    return run;
}

// The last CTA's job for k <= 32: L sorted lists of k keys (cand[c * k + j]) -> the best k, ascending, in dst[0..k) and
// smem[0..32). Sorting every key costs ~2 shuffles per key and stage on the ONE SM this CTA runs on (measured: 7.6 us for
// 296 x 10 keys, 19 us for 296 x 32 — shuffle-throughput bound), so the keys are filtered first: T = the k-th smallest of
// the lists' MINIMA is an upper bound of the global k-th best (the minima are L distinct keys), and only keys <= T —
// at least k, typically a few dozen, at most k per list of the k lists with the smallest minima — are collected
// (ballot-compacted into shared memory) and sorted. More than 256 survivors (never seen; possible when fewer than k
// lists are non-empty) fall back to folding everything. `extra` = this thread's optional extra key (zero-norm ids).
// smem: [2 * SCAN_WARPS * 32] keys (SCAN_SMALL_SMEM).
__device__ __forceinline__ void cta_topk_of_lists32(const volatile uint64_t *cand, uint32_t L, uint32_t k, uint64_t extra_run,
                                                    uint64_t *smem, uint64_t *dst, int warp, int lane)
{
    __shared__ unsigned s_nsurv;
    const uint64_t total = (uint64_t)L * k;
    uint64_t *surv = smem + SCAN_WARPS * 32;
    if (threadIdx.x == 0) s_nsurv = 0;
    // 1. T: k-th smallest of the L minima
    const uint64_t mins = warp_fold_keys32<2>(KEY_EMPTY, L, [&](uint64_t c) { return cand[c * k]; }, warp, lane);
    cta_merge_lists32(mins, smem, 0, nullptr, warp, lane);
    const uint64_t T = smem[k - 1];
    __syncthreads();
    // 2. survivors: keys <= T, four loads per thread in flight
    for (uint64_t t0 = 0; t0 < total; t0 += 4 * SCAN_THREADS) {
        uint64_t key[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint64_t t = t0 + (uint64_t)u * SCAN_THREADS + threadIdx.x;
            key[u] = t < total ? cand[t] : KEY_EMPTY;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const bool pass = key[u] <= T && key[u] != KEY_EMPTY;
            const unsigned m = __ballot_sync(FULL, pass);
            if (m) {
                unsigned base = 0;
                if (lane == 0) base = atomicAdd(&s_nsurv, (unsigned)__popc(m));
                base = __shfl_sync(FULL, base, 0);
                const unsigned pos = base + __popc(m & ((1u << lane) - 1u));
                if (pass && pos < SCAN_WARPS * 32) surv[pos] = key[u];
            }
        }
    }
    __syncthreads();
    const unsigned n = s_nsurv;
    uint64_t run;
    if (n <= SCAN_WARPS * 32) {   // one batch per warp
        const unsigned i = warp * 32 + lane;
        run = warp_sort32(i < n ? surv[i] : KEY_EMPTY, lane);
    } else {
        run = warp_fold_keys32<8>(KEY_EMPTY, total, [&](uint64_t t) { return cand[t]; }, warp, lane);
    }
    run = warp_merge_low32(run, extra_run, lane);
    __syncthreads();
    cta_merge_lists32(run, smem, k, dst, warp, lane);
}

__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

template <bool BIG>
struct SelOf { using type = WarpSel32; };
template <>
struct SelOf<true> { using type = WarpSelBig; };
// selectors of the scan kernels proper: per-warp registers for k <= 32, one CTA-shared buffer above
template <bool BIG>
struct ScanSelOf { using type = WarpSel32; };
template <>
struct ScanSelOf<true> { using type = CtaSel; };

// Tail of the last CTA when a.xchg is set. Precondition: smem[0..k) holds this rank's sorted local top-k
// (left there by cta_reduce) and all threads are past a barrier.
template <bool BIG, class Sel>
__device__ __forceinline__ void exchange_and_merge(const ScanArgs &a, Sel &sel, uint64_t *smem, int warp, int lane)
{
    const ExchangeDev &x = *a.xchg;   // pointer table stays in global memory (no local copy)
    const uint32_t par = a.seq & 1u, W = x.world, R = x.rank, KM = x.kmax, k = a.k;
    const bool gather = x.root >= 0;
    if (gather && R != (uint32_t)x.root) {
        // gather mode, not the root: k keys -> my slot in the root's block (peer stores over NVLink), flag, done
        uint64_t *dst = x.slots[x.root] + ((size_t)par * W + R) * KM;
        for (uint32_t j = threadIdx.x; j < k; j += blockDim.x) dst[j] = smem[j];
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            *reinterpret_cast<volatile unsigned *>(x.flags[x.root] + par * W + R) = a.seq;
        }
        return;
    }
    // my k keys -> slot [par][R] of every rank (own copy included; gather mode: own copy only)
    const uint32_t n_dst = gather ? 1u : W;
    for (uint32_t t = threadIdx.x; t < n_dst * k; t += blockDim.x) {
        const uint32_t p = gather ? R : t / k, j = gather ? t : t - p * k;
        x.slots[p][((size_t)par * W + R) * KM + j] = smem[j];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < W) {
        __threadfence_system();
        if (!gather || threadIdx.x == R)
            *reinterpret_cast<volatile unsigned *>(x.flags[threadIdx.x] + par * W + R) = a.seq;   // publish to rank threadIdx.x
        // wait until rank threadIdx.x's keys of this query have landed in MY memory
        const volatile unsigned *f = x.flags[R] + par * W + threadIdx.x;
        const unsigned long long t0 = global_timer_ns(), limit = x.timeout_ns;
        unsigned long long waited = 0;
        while (*f != a.seq) {
            waited = global_timer_ns() - t0;
            if (waited > limit) {   // a peer died or never launched: results are undefined, the host is told
                *reinterpret_cast<volatile unsigned *>(x.status) = 1u;
                break;
            }
        }
        if (x.wait_ring != nullptr) x.wait_ring[(size_t)((a.seq - 1u) % XCHG_STATS_RING) * XCHG_MAX_WORLD + threadIdx.x] = waited;
        __threadfence_system();
    }
    __syncthreads();
    const volatile uint64_t *mine = x.slots[R] + (size_t)par * W * KM;   // volatile: peers wrote it, L1 may be stale
    const uint32_t total = W * k;
    if constexpr (BIG) {
        sel.reset();
        cta_buf_stream(sel.cb, sel.cap, k, total, [&](uint64_t t) { const uint32_t r = (uint32_t)t / k, j = (uint32_t)t - r * k; return mine[(size_t)r * KM + j]; });
        sel.finish();
        for (uint32_t j = threadIdx.x; j < k; j += blockDim.x) a.out_keys[j] = smem[j];
        __syncthreads();
    } else {
        const uint64_t run = warp_fold_keys32(KEY_EMPTY, total, [&](uint64_t t) {
            const uint32_t r = (uint32_t)t / k, j = (uint32_t)t - r * k;
            return mine[(size_t)r * KM + j];
        }, warp, lane);
        cta_merge_lists32(run, smem, a.k, a.out_keys, warp, lane);
    }
}

// V = float4 per lane (ceil(dim4/32)); EXACT: dim4 == 32*V; R = rows in flight per warp.
// Filtered scan body (csgpu_search_filtered): the allow bitmap is tested BEFORE a row is read, so masked rows cost
// no HBM traffic at all — bytes scanned = passing rows x dim x 4 (+ 4 B of id per row). A warp walks blocks of 32
// consecutive rows: lane l owns row 32b + l, loads its chunk id (one coalesced 128 B line per block) and its bitmap
// word (id-indexed, L2-resident: 1.25 MB per 10M ids); the ballot of allowed lanes is then consumed R rows at a time
// with the same per-row FMA chain + shuffle tree as the unfiltered body (bit-identical scores). The id and bitmap
// loads run two / one blocks ahead so their latency stays off the critical path.
// Which 32-row block a warp takes next. Fixed stride (a.static_split, A/B runs): block gw + i * n_warps. Otherwise from the
// global work counter (a.ticket + 1), in chunks of up to SCAN_CHUNK consecutive blocks while plenty is left and single
// blocks near the end, the next grab issued before the current chunk is consumed — SMs stream at visibly different
// rates, and the fixed split leaves bandwidth idle at the end of every query (the unfiltered scan's DYN path, below,
// measured 1.5-2.8 %). Per-warp grabs for the per-warp register selector; for the CTA-shared selector (k > 32) the CTA
// grabs "CTA iterations" (one block per warp) so that every warp runs the same number of iterations — the selector
// synchronises inside the loop. `live` is false once there is nothing left (CTA-uniform for the CTA-shared selector).
constexpr int SCAN_CHUNK = 8;
template <bool CTA_SHARED>
struct BlockCursor {
    uint64_t n_blocks, gw, n_warps, i = 0, n_iters;
    uint32_t cur = 0, end = 0, c_next, nxt = 0, n_units, div;
    unsigned *work;
    uint32_t *s_start;
    bool dynamic, done = false;
    int lane, warp;

    __device__ __forceinline__ void init(const void *, unsigned *work_, uint32_t *s_start_, uint64_t n_blocks_, uint64_t gw_,
                                         uint64_t n_warps_, int lane_, int warp_, bool dynamic_)
    {
        n_blocks = n_blocks_; gw = gw_; n_warps = n_warps_; work = work_; s_start = s_start_; lane = lane_; warp = warp_; dynamic = dynamic_;
        const uint64_t b_first = gw - (gw % SCAN_WARPS);   // same trip count for every warp of the CTA
        n_iters = b_first < n_blocks ? (n_blocks - b_first + n_warps - 1) / n_warps : 0;
        n_units = CTA_SHARED ? (uint32_t)((n_blocks + SCAN_WARPS - 1) / SCAN_WARPS) : (uint32_t)n_blocks;
        div = 2u * (uint32_t)(CTA_SHARED ? n_warps / SCAN_WARPS : n_warps);
        c_next = max(1u, min((uint32_t)SCAN_CHUNK, n_units / div));
        if (dynamic && (CTA_SHARED ? threadIdx.x == 0 : lane == 0)) nxt = atomicAdd(work, c_next);
    }
    // returns the block index (>= n_blocks: nothing to do for this warp) and sets live
    __device__ __forceinline__ uint64_t next(bool &live)
    {
        if (!dynamic) {
            live = i < n_iters;
            return live ? gw + (i++) * n_warps : n_blocks;
        }
        if (cur >= end) {
            if (done) { live = false; return n_blocks; }
            uint32_t s;
            if constexpr (CTA_SHARED) {
                if (threadIdx.x == 0) *s_start = nxt;
                __syncthreads();
                s = *s_start;
                __syncthreads();   // nobody still reads s_start when thread 0 rewrites it at the next boundary
            } else {
                s = __shfl_sync(FULL, nxt, 0);
            }
            if (s >= n_units) { done = true; live = false; return n_blocks; }
            cur = s;
            end = min(s + c_next, n_units);
            c_next = max(1u, min((uint32_t)SCAN_CHUNK, (n_units - s) / div));
            if (CTA_SHARED ? threadIdx.x == 0 : lane == 0) nxt = atomicAdd(work, c_next);
        }
        live = true;
        const uint32_t u = cur++;
        return CTA_SHARED ? (uint64_t)u * SCAN_WARPS + warp : (uint64_t)u;
    }
};

template <int V, bool EXACT, int R, int LD, class Sel>
__device__ __forceinline__ void scan_rows_filtered(const ScanArgs &a, const float4 (&qv)[V], bool qzero, Sel &sel, int lane, int warp,
                                                   uint64_t gw, uint64_t n_warps, uint32_t *s_start)
{
    const uint64_t n = a.n_rows;
    const uint64_t n_blocks = (n + 31) / 32;
    // Two filter flavours share the pipeline (warp-uniform branch): the id bitmap (lane loads ids[row], then the
    // bitmap word of that id) and the row-tag predicate (lane loads tags[row]; language bit + file range are tested
    // on the tag, the optional per-file bitmap word is the second-stage load). In tag mode `id_*` holds the FILE id
    // with bit 31 set when the static part of the predicate already failed; chunk ids are fetched at insert time.
    const bool tagmode = a.tags != nullptr;
    const bool have_bm = a.bitmap != nullptr;
    auto load_id = [&](uint64_t b, bool &valid) -> uint32_t {
        const uint64_t row = b * 32 + lane;
        valid = b < n_blocks && row < n;
        if (!valid) return 0u;
        if (!tagmode) return __ldg(a.ids + row);
        const uint32_t tag = __ldg(a.tags + row);
        valid = tag_pass_static(tag, a.lang_mask, a.file_lo, a.file_hi);
        return tag & 0x07FFFFFFu;
    };
    auto load_word = [&](uint32_t id, bool valid) -> uint64_t {
        if (tagmode && !have_bm) return valid ? ~0ull : 0ull;
        return (valid && (uint64_t)id < a.n_bits) ? __ldg(reinterpret_cast<const unsigned long long *>(a.bitmap) + (id >> 6)) : 0ull;
    };
    BlockCursor<Sel::CTA_SHARED> cursor;
    cursor.init(nullptr, a.ticket + 1, s_start, n_blocks, gw, n_warps, lane, warp, a.static_split == 0);
    // software pipeline: the id / tag of a block is loaded two blocks ahead of its use, the bitmap word one block ahead
    bool v_cur, v_nxt, v_nx2, l_cur, l_nxt, l_nx2;
    uint64_t b_cur = cursor.next(l_cur);
    uint32_t id_cur = load_id(b_cur, v_cur);
    uint64_t w_cur = load_word(id_cur, v_cur);
    uint64_t b_nxt = cursor.next(l_nxt);
    uint32_t id_nxt = load_id(b_nxt, v_nxt);
    for (uint32_t it = 0; l_cur; ++it) {
        const uint64_t w_nxt = load_word(id_nxt, v_nxt);
        const uint64_t b_nx2 = cursor.next(l_nx2);
        const uint32_t id_nx2 = load_id(b_nx2, v_nx2);
        const uint64_t b = b_cur;
        unsigned m = __ballot_sync(FULL, (w_cur >> (id_cur & 63)) & 1ull);
        const float4 *blk = a.rows + b * 32 * a.dim4 + lane;
        while (m) {
            int rl[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                rl[r] = m ? (__ffs(m) - 1) : -1;
                m &= m - 1;   // 0 stays 0
            }
            float4 x[R][V];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float4 *p = blk + (size_t)(rl[r] < 0 ? 0 : rl[r]) * a.dim4;
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    if (rl[r] >= 0 && (EXACT || lane + 32 * j < a.dim4)) x[r][j] = ldg_stream<LD>(p + 32 * j);
                    else x[r][j] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float acc = 0.f;
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    acc = fmaf(x[r][j].x, qv[j].x, acc); acc = fmaf(x[r][j].y, qv[j].y, acc);
                    acc = fmaf(x[r][j].z, qv[j].z, acc); acc = fmaf(x[r][j].w, qv[j].w, acc);
                }
                acc = warp_sum_tree(acc);
                const float dist = qzero ? 0.f : fmaf(-0.5f, acc, 0.5f);
                if (rl[r] >= 0 && okey(dist) <= (uint32_t)(sel.thr >> 32)) {   // warp-uniform
                    const uint32_t id = tagmode ? __ldg(a.ids + b * 32 + rl[r]) : __shfl_sync(FULL, id_cur, rl[r]);
                    const uint64_t key = make_key(dist, id);
                    if (key < sel.thr) sel.insert(key, lane);
                }
            }
        }
        b_cur = b_nxt; l_cur = l_nxt; id_cur = id_nxt; v_cur = v_nxt; w_cur = w_nxt;
        b_nxt = b_nx2; l_nxt = l_nx2; id_nxt = id_nx2; v_nxt = v_nx2;
        if constexpr (Sel::CTA_SHARED) { if (it & 1) sel.sync_point(2 * 32 * SCAN_WARPS); }
    }
}

// DYN (unfiltered scans): warps take their row groups from a global work counter instead of a fixed stride. SMs
// stream at visibly different rates (per-CTA timestamps: with the fixed stride the first CTA is done 25 % before the
// last), so the fixed split leaves bandwidth idle at the end of every query; chunks of up to SCAN_CHUNK groups, single
// groups near the end, next grab issued before the current chunk is processed. Which warp scans a row does not change
// its score, and selection is a total order, so results are bit-identical either way.
template <int V, bool EXACT, int R, bool BIG, int OCC = (BIG ? 1 : 2), int LD = 0, bool FILT = false, bool DYN = !FILT>
__global__ void __launch_bounds__(SCAN_THREADS, OCC) scan_topk_kernel(const ScanArgs a)
{
    extern __shared__ __align__(16) uint64_t smem[];
    __shared__ bool is_last;
    if (a.run_if != nullptr && *reinterpret_cast<const volatile unsigned *>(a.run_if) == 0) return;   // CTA-uniform
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (a.timing && threadIdx.x == 0) a.timing[blockIdx.x * 4 + 0] = global_timer_ns();
    if (a.t0_slot != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *a.t0_slot = global_timer_ns();
    // k <= 32: per-warp register selector. k > 32: ONE candidate buffer per CTA (a.kpad = its capacity) + threshold
    using Sel = typename ScanSelOf<BIG>::type;
    __shared__ unsigned cb_cnt;
    __shared__ uint64_t cb_thr;
    Sel sel;
    if constexpr (BIG) {
        sel.cb.buf = smem; sel.cb.cnt = &cb_cnt; sel.cb.thr = &cb_thr; sel.cap = a.kpad; sel.k = a.k;
        sel.reset();
    } else {
        sel.init(a.k);
    }

    // ---- query -> registers, scaled to unit length -------------------------------------
    float4 qv[V];
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) {
        const uint32_t c = lane + 32 * j;
        if (EXACT || c < a.dim4) qv[j] = reinterpret_cast<const float4 *>(a.q)[c];
        else qv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        ss = fmaf(qv[j].x, qv[j].x, ss); ss = fmaf(qv[j].y, qv[j].y, ss);
        ss = fmaf(qv[j].z, qv[j].z, ss); ss = fmaf(qv[j].w, qv[j].w, ss);
    }
    ss = warp_sum_tree(ss);
    const bool qzero = !(ss > 0.f);  // zero-norm query: every distance is 0.0 (arroy pn*qn == 0)
    const float qinv = qzero ? 0.f : 1.0f / sqrtf(ss);
#pragma unroll
    for (int j = 0; j < V; ++j) { qv[j].x *= qinv; qv[j].y *= qinv; qv[j].z *= qinv; qv[j].w *= qinv; }

    // ---- stream the rows ---------------------------------------------------------------
    const uint64_t n = a.n_rows;
    const uint64_t gw = (uint64_t)blockIdx.x * SCAN_WARPS + warp;
    const uint64_t n_warps = (uint64_t)gridDim.x * SCAN_WARPS;
    // same trip count for every warp of the CTA (rows past the end are predicated off): the CTA-shared selector
    // synchronises every 8 iterations
    const uint64_t n_groups = (n + R - 1) / R, g_first = (uint64_t)blockIdx.x * SCAN_WARPS;
    const uint64_t n_iters = g_first < n_groups ? (n_groups - g_first + n_warps - 1) / n_warps : 0;
    // one group = R consecutive rows of this warp: R x V independent 128-bit loads, then R fixed FMA chains + trees
    auto scan_group = [&](uint64_t base) {
        float4 x[R][V];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const uint64_t row = base + r;
            const float4 *p = a.rows + row * a.dim4 + lane;
#pragma unroll
            for (int j = 0; j < V; ++j) {
                if (row < n && (EXACT || lane + 32 * j < a.dim4)) x[r][j] = ldg_stream<LD>(p + 32 * j);
                else x[r][j] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < V; ++j) {
                acc = fmaf(x[r][j].x, qv[j].x, acc); acc = fmaf(x[r][j].y, qv[j].y, acc);
                acc = fmaf(x[r][j].z, qv[j].z, acc); acc = fmaf(x[r][j].w, qv[j].w, acc);
            }
            acc = warp_sum_tree(acc);
            const float dist = qzero ? 0.f : fmaf(-0.5f, acc, 0.5f);  // (1 - cos)/2
            const uint64_t row = base + r;
            if (row < n && okey(dist) <= (uint32_t)(sel.thr >> 32)) {  // warp-uniform, rare
                const uint32_t id = a.ids[row];
                const uint64_t key = make_key(dist, id);
                if (key < sel.thr) sel.insert(key, lane);
            }
        }
    };
    __shared__ uint32_t s_start;   // CTA-level work grabs (dynamic split with the CTA-shared selector)
    if constexpr (FILT) scan_rows_filtered<V, EXACT, R, LD>(a, qv, qzero, sel, lane, warp, gw, n_warps, &s_start);
    else if constexpr (DYN && !BIG) {
        // per-warp grabs of up to SCAN_CHUNK groups
        const uint32_t n_grp = (uint32_t)n_groups, nw2 = 2u * (uint32_t)n_warps;   // n_rows < 2^32, R >= 2
        unsigned *work = a.ticket + 1;
        uint32_t c_next = max(1u, min((uint32_t)SCAN_CHUNK, n_grp / nw2)), nxt = 0;
        if (lane == 0) nxt = atomicAdd(work, c_next);
        for (;;) {
            const uint32_t c_start = __shfl_sync(FULL, nxt, 0);
            if (c_start >= n_grp) break;
            const uint32_t c_end = min(c_start + c_next, n_grp);
            c_next = max(1u, min((uint32_t)SCAN_CHUNK, (n_grp - c_start) / nw2));
            if (lane == 0) nxt = atomicAdd(work, c_next);
            for (uint32_t grp = c_start; grp < c_end; ++grp) scan_group((uint64_t)grp * R);
        }
    } else if constexpr (DYN) {
        // CTA-shared selector: the CTA synchronises at every sync point anyway, so the whole CTA grabs up to SCAN_CHUNK
        // "CTA iterations" (one group per warp) at a time; thread 0 issues the next grab before the chunk is processed
        const uint32_t n_cit = (uint32_t)((n_groups + SCAN_WARPS - 1) / SCAN_WARPS), g2 = 2u * gridDim.x;
        unsigned *work = a.ticket + 1;
        uint32_t c_next = max(1u, min((uint32_t)SCAN_CHUNK, n_cit / g2)), nxt = 0;
        if (threadIdx.x == 0) nxt = atomicAdd(work, c_next);
        for (;;) {
            if (threadIdx.x == 0) s_start = nxt;
            __syncthreads();
            const uint32_t c_start = s_start;
            if (c_start >= n_cit) break;
            const uint32_t c_end = min(c_start + c_next, n_cit);
            c_next = max(1u, min((uint32_t)SCAN_CHUNK, (n_cit - c_start) / g2));
            if (threadIdx.x == 0) nxt = atomicAdd(work, c_next);
            for (uint32_t it = c_start; it < c_end; ++it) scan_group(((uint64_t)it * SCAN_WARPS + warp) * R);
            sel.sync_point(SCAN_CHUNK * R * SCAN_WARPS);   // barrier inside: s_start is free to be rewritten
        }
    } else {
        for (uint64_t it = 0; it < n_iters; ++it) {
            scan_group((gw + it * n_warps) * R);
            if constexpr (BIG) { if ((it & 7) == 7) sel.sync_point(8 * R * SCAN_WARPS); }
        }
    }

    // ---- CTA top-k -> cand[blockIdx.x] ---------------------------------------------------
    if (a.timing) { __syncthreads(); if (threadIdx.x == 0) a.timing[blockIdx.x * 4 + 1] = global_timer_ns(); }
    if constexpr (BIG) {
        sel.finish();
        uint64_t *dst = a.cand + (size_t)blockIdx.x * a.k;
        for (uint32_t j = threadIdx.x; j < a.k; j += blockDim.x) dst[j] = smem[j];
        __syncthreads();
    } else {
        cta_merge_lists32(sel.v, smem, a.k, a.cand + (size_t)blockIdx.x * a.k, warp, lane);   // the selector keeps v sorted by lane
    }

    // ---- last CTA merges all CTAs' results (threadfence-reduction pattern) ---------------
    __threadfence();
    if (threadIdx.x == 0) {
        if (a.timing) a.timing[blockIdx.x * 4 + 2] = global_timer_ns();
        unsigned t = atomicAdd(a.ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (a.timing && threadIdx.x == 0) a.timing[gridDim.x * 4 + 0] = global_timer_ns();

    const volatile uint64_t *cand = a.cand;
    if constexpr (BIG) {
        sel.reset();
        // Every CTA's list is sorted ascending, so the lists are consumed COLUMN-wise: column j = the j-th key of every
        // list, one load latency per column. Column 0 alone (one key per CTA) usually holds >= k keys, so the first
        // compaction already yields a tight threshold (the k-th smallest of the CTAs' minima), and the walk stops at the
        // first column in which no list beats the threshold — a handful of columns instead of gridDim.x * k / 256
        // dependent rounds (0.2 ms of tail at k = 200).
        {
            const uint32_t L = gridDim.x;   // host guarantees k + L <= cap (L <= 2 * SMs, cap >= 2k + 512)
            for (uint32_t j = 0; j < a.k; ++j) {
                bool passed = false;
                for (uint32_t c = threadIdx.x; c < L; c += blockDim.x) {
                    const uint64_t key = cand[(size_t)c * a.k + j];
                    if (key < *sel.cb.thr) { sel.cb.buf[atomicAdd(sel.cb.cnt, 1u)] = key; passed = true; }
                }
                if (!__syncthreads_or(passed)) break;
                const unsigned cnt = *reinterpret_cast<volatile unsigned *>(sel.cb.cnt);
                const bool need = cnt + L > sel.cap || (*sel.cb.thr == KEY_EMPTY && cnt >= a.k);
                if (__syncthreads_or(need)) cta_buf_compact(sel.cb.buf, sel.cb.cnt, sel.cb.thr, sel.cap, a.k);
            }
        }
        // zero-norm rows: distance 0.0, ascending id; the first k allowed ones suffice
        uint32_t found = 0;
        for (uint32_t b = 0; b < a.n_zero && found < a.k; b += blockDim.x) {
            uint64_t key = KEY_EMPTY;
            if (b + threadIdx.x < a.n_zero) {
                const uint32_t id = a.zero_ids[b + threadIdx.x];
                if (zero_row_allowed(a.bitmap, a.n_bits, a.tags, a.lang_mask, a.file_lo, a.file_hi, a.zero_ids, a.n_zero, b + threadIdx.x))
                    key = make_key(0.f, id);
            }
            found += __syncthreads_count(key != KEY_EMPTY);
            cta_buf_stream(sel.cb, sel.cap, a.k, blockDim.x, [&](uint64_t) { return key; });
        }
        sel.finish();   // smem[0..k) = this rank's top-k
        if (a.xchg == nullptr) {
            for (uint32_t j = threadIdx.x; j < a.k; j += blockDim.x) a.out_keys[j] = smem[j];
        } else {
            exchange_and_merge<true>(a, sel, smem, warp, lane);
        }
    } else {
        // zero-norm rows: distance 0.0, ascending id; the first k allowed ones suffice (CTA-uniform loop) -> one sorted list per warp
        uint64_t zrun = KEY_EMPTY;
        if (a.n_zero) {
            uint32_t found = 0;
            for (uint32_t b = 0; b < a.n_zero && found < a.k; b += SCAN_WARPS * 32) {
                uint64_t key = KEY_EMPTY;
                const uint32_t i = b + threadIdx.x;
                if (i < a.n_zero && zero_row_allowed(a.bitmap, a.n_bits, a.tags, a.lang_mask, a.file_lo, a.file_hi, a.zero_ids, a.n_zero, i))
                    key = make_key(0.f, a.zero_ids[i]);
                found += __syncthreads_count(key != KEY_EMPTY);
                zrun = warp_merge_low32(zrun, warp_sort32(key, lane), lane);
            }
        }
        // 2 x SMs lists of k keys -> the best k (threshold from the lists' minima, survivors sorted; cta_topk_of_lists32)
        if (a.xchg == nullptr) {
            cta_topk_of_lists32(cand, gridDim.x, a.k, zrun, smem, a.out_keys, warp, lane);
            if (a.timing && threadIdx.x == 0) a.timing[gridDim.x * 4 + 1] = global_timer_ns();
        } else {
            // local top-k stays in smem[0..k) (cand[0] is a scratch destination), then exchange + global merge
            cta_topk_of_lists32(cand, gridDim.x, a.k, zrun, smem, a.cand, warp, lane);
            if (a.timing && threadIdx.x == 0) a.timing[gridDim.x * 4 + 1] = global_timer_ns();
            exchange_and_merge<false>(a, sel, smem, warp, lane);
        }
    }
    if (threadIdx.x == 0) {
        a.ticket[0] = 0; a.ticket[1] = 0;
        if (a.timing) a.timing[gridDim.x * 4 + 2] = global_timer_ns();
        // CTA 0 stored t0 before its __threadfence() + ticket; this (last) CTA read every ticket after them
        if (a.elapsed_out != nullptr) *a.elapsed_out = global_timer_ns() - *reinterpret_cast<volatile unsigned long long *>(a.t0_slot);
    }
}

// The fused exchange on its own (one CTA): this rank's sorted local top-k `local[0..k)` -> peer stores, flags, wait,
// global merge -> a.out_keys. Used when the local keys come from a kernel without the exchange in its tail (the byte
// prefilter's scan_i8_kernel, possibly followed by the conditional fp32 scan).
template <bool BIG>
__global__ void __launch_bounds__(SCAN_THREADS, 1) exchange_keys_kernel(const ScanArgs a, const uint64_t *__restrict__ local)
{
    extern __shared__ __align__(16) uint64_t smem[];
    __shared__ unsigned cb_cnt;
    __shared__ uint64_t cb_thr;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    using Sel = typename ScanSelOf<BIG>::type;
    Sel sel;
    if constexpr (BIG) { sel.cb.buf = smem; sel.cb.cnt = &cb_cnt; sel.cb.thr = &cb_thr; sel.cap = a.kpad; sel.k = a.k; }
    else sel.init(a.k);
    for (uint32_t j = threadIdx.x; j < a.k; j += blockDim.x) smem[j] = local[j];
    __syncthreads();
    exchange_and_merge<BIG>(a, sel, smem, warp, lane);
}

constexpr size_t MERGE_SMEM_MAX = (size_t)2 * SCAN_WARPS * 1024 * sizeof(uint64_t);   // merge_keys_kernel<true> at k = 1024

// Generic k-way merge: for query b = blockIdx.x, n_lists lists of k keys -> top-k. One CTA per query.
// keys layout [n_lists][nq = gridDim.x][k]; out [nq][k]. Used for the cross-GPU merge.
template <bool BIG>
__global__ void __launch_bounds__(SCAN_THREADS, 1)
merge_keys_kernel(const uint64_t *__restrict__ keys, uint32_t n_lists, uint32_t k, uint32_t kpad,
                  uint64_t *__restrict__ out)
{
    extern __shared__ __align__(16) uint64_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t total = (uint64_t)n_lists * k;
    const size_t list_stride = (size_t)gridDim.x * k, q_off = (size_t)blockIdx.x * k;
    auto load = [&](uint64_t t) {
        const uint64_t l = t / k, i = t - l * k;
        return keys[l * list_stride + q_off + i];
    };
    if constexpr (!BIG) {
        const uint64_t run = warp_fold_keys32(KEY_EMPTY, total, load, warp, lane);
        cta_merge_lists32(run, smem, k, out + q_off, warp, lane);
    } else {
        WarpSelBig sel;
        sel.init(smem + (size_t)warp * kpad, smem + (size_t)(SCAN_WARPS + warp) * kpad, k, kpad, lane);
        for (uint64_t b = (uint64_t)warp * 32; b < total; b += SCAN_WARPS * 32) {
            const uint64_t key = b + lane < total ? load(b + lane) : KEY_EMPTY;
            offer_lane_keys(sel, key, lane);
        }
        cta_reduce<true>(sel, smem, k, kpad, out + q_off, warp, lane);
    }
}

// Device-side dedup of the per-variant lists of one user query (the HashMap + BinaryHeap pass at
// /root/reference/src/search/mod.rs:513-590): keys [total] (any order, KEY_EMPTY = hole) -> per chunk id the entry
// with the smallest distance -> the best k_out of the union, ascending (distance, id). One CTA; two bitonic sorts
// in shared memory (by (id, distance), then by (distance, id)). npad = pow2 >= total.
static __global__ void __launch_bounds__(SCAN_THREADS, 1)
dedup_variants_kernel(const uint64_t *__restrict__ keys, uint32_t total, uint32_t npad, uint32_t k_out, uint64_t *__restrict__ out)
{
    extern __shared__ __align__(16) uint64_t smem[];
    for (uint32_t t = threadIdx.x; t < npad; t += blockDim.x) {
        const uint64_t key = t < total ? keys[t] : KEY_EMPTY;
        smem[t] = key == KEY_EMPTY ? KEY_EMPTY : ((key << 32) | (key >> 32));   // (id, okey(distance))
    }
    cta_sort(smem, npad);
    // keep the first (= closest) entry of every id run. Batches run from the top down: batch i reads element
    // i*B - 1, which only a LOWER batch may overwrite later.
    const uint32_t B = blockDim.x;
    for (int64_t base = (int64_t)((npad - 1) / B) * B; base >= 0; base -= B) {
        const uint32_t t = (uint32_t)base + threadIdx.x;
        uint64_t v = KEY_EMPTY;
        if (t < npad) {
            const uint64_t cur = smem[t];
            const bool dup = cur == KEY_EMPTY || (t > 0 && (smem[t - 1] >> 32) == (cur >> 32));
            v = dup ? KEY_EMPTY : ((cur << 32) | (cur >> 32));                   // back to (okey(distance), id)
        }
        __syncthreads();
        if (t < npad) smem[t] = v;
        __syncthreads();
    }
    cta_sort(smem, npad);
    for (uint32_t j = threadIdx.x; j < k_out; j += blockDim.x) out[j] = j < npad ? smem[j] : KEY_EMPTY;
}

// The same dedup without the two full sorts (round 2: they were ~30 us of a 220 us hybrid search over 100k rows): a hash table
// in shared memory keyed by chunk id keeps, per id, the smallest (distance, id) key (open addressing, 64-bit atomicCAS to claim
// a slot, atomicMin to improve it); the occupied slots — one per distinct id, typically a quarter of the keys — are compacted
// and only THEY are sorted. Same result as dedup_variants_kernel (a set operation followed by a sort on a total order).
// smem: table[tbl] | compact[cpad] keys; tbl = pow2 >= 2 * total, cpad = pow2 >= total. total <= 8192.
static __global__ void __launch_bounds__(SCAN_THREADS, 1)
dedup_variants_hash_kernel(const uint64_t *__restrict__ keys, uint32_t total, uint32_t tbl, uint32_t cpad, uint32_t k_out, uint64_t *__restrict__ out)
{
    extern __shared__ __align__(16) uint64_t smem[];
    __shared__ unsigned s_n;
    unsigned long long *table = reinterpret_cast<unsigned long long *>(smem);
    uint64_t *compact = smem + tbl;
    for (uint32_t t = threadIdx.x; t < tbl; t += blockDim.x) table[t] = KEY_EMPTY;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    const uint32_t mask = tbl - 1;
    for (uint32_t t = threadIdx.x; t < total; t += blockDim.x) {
        const uint64_t key = keys[t];
        if (key == KEY_EMPTY) continue;
        const uint32_t id = (uint32_t)key;
        uint32_t slot = (id * 2654435761u) & mask;
        for (;;) {   // terminates: the table has at least twice as many slots as there are keys
            const unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(&table[slot]);
            if (cur == KEY_EMPTY) {
                const unsigned long long prev = atomicCAS(&table[slot], (unsigned long long)KEY_EMPTY, (unsigned long long)key);
                if (prev == KEY_EMPTY) break;                                           // claimed
                if ((uint32_t)prev == id) { atomicMin(&table[slot], (unsigned long long)key); break; }
            } else if ((uint32_t)cur == id) {
                atomicMin(&table[slot], (unsigned long long)key);                       // same id: keep the smaller distance
                break;
            }
            slot = (slot + 1) & mask;
        }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    for (uint32_t t0 = 0; t0 < tbl; t0 += blockDim.x) {   // compact the occupied slots (order is irrelevant: they are sorted next)
        const uint32_t t = t0 + threadIdx.x;
        const uint64_t v = t < tbl ? (uint64_t)table[t] : KEY_EMPTY;
        const unsigned m = __ballot_sync(FULL, v != KEY_EMPTY);
        unsigned base = 0;
        if (lane == 0 && m) base = atomicAdd(&s_n, (unsigned)__popc(m));
        base = __shfl_sync(FULL, base, 0);
        if (v != KEY_EMPTY) compact[base + __popc(m & ((1u << lane) - 1u))] = v;
    }
    __syncthreads();
    const uint32_t n = s_n;
    const uint32_t npad = pow2_at_least(n, 32);           // <= cpad
    for (uint32_t t = n + threadIdx.x; t < npad; t += blockDim.x) compact[t] = KEY_EMPTY;
    cta_sort(compact, npad);
    for (uint32_t j = threadIdx.x; j < k_out; j += blockDim.x) out[j] = j < npad ? compact[j] : KEY_EMPTY;
}

}  // namespace csgpu
