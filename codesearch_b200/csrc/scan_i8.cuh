// scan_i8.cuh — the opt-in "byte prefilter" of an fp32 index (csgpu_set_byte_prefilter): exact fp32 single-query
// results from ONE pass over a 1-byte-per-element shadow of the corpus instead of the 4-byte rows.
//
// Same contract as the tensor prefilter (rescore.cuh), aimed at the single-query path of
// /root/reference/src/vectordb/store.rs:446-459: the shadow is only a FILTER with a PROVEN per-row error bound; the
// rows that survive it (a few hundred of 10M) are rescored from the fp32 rows with exactly the arithmetic of
// scan_topk_kernel (same query normalisation, FMA chain, xor-shuffle tree, distance = fma(-0.5, cos, 0.5)) and the
// top-k is selected on those exact keys. Ids AND distances are therefore bit-identical to csgpu_search, while the
// kernel moves dim + 4 bytes per row instead of 4 * dim: the HBM floor drops from 2.14 ms to ~0.55 ms at 10M x 384.
//
// Shadow layout (HBM): row r = d8 = 128 * ceil(dim / 128) signed bytes (zero padded), xi = rint(x / s_r) in
// [-127, 127] with a per-row scale s_r; meta[r] = half2(s_r, e_r), e_r >= ||x - s_r * xi||_2 (computed in f64 at
// build, rounded UP to half). A row is 3 full 128-byte lines at dim = 384.
//
// Error bound. q = the fp32 unit query exactly as scan_topk_kernel holds it; q^ = s_q * qi its int8 image with
// e_q >= ||q - q^||_2. For a row x with image x^ = s_r * xi:
//     q.x - q^.x^ = (q - q^).x + q^.(x - x^)        |.| <= e_q ||x|| + ||q^|| e_r  =: M_r           (Cauchy-Schwarz)
// q^.x^ = s_q s_r * (integer dot product, exact in int32: 1024 * 127^2 < 2^24). Stored rows are unit vectors to
// within 1e-6 (normalise_rows_kernel), covered by the factor folded into e_q. The scan kernel's own fp32 result differs
// from the real-number q.x by <= (4V + 5) 2^-24 (its longest FMA/add chain) < 2.3e-6 at V = 8, i.e. 1.2e-6 in
// distance; I8_SLACK = 3e-6 (distance units) covers that plus the handful of roundings below. Hence for every row
//     lb_r = d^_r - M_r/2 - SLACK  <=  d_fp32(r)  <=  d^_r + M_r/2 + SLACK = ub_r.
//
// Threshold without a global selector. Every streaming warp publishes the smallest ub it has seen (one u32 per warp
// in HBM/L2). A helper warp per CTA keeps recomputing G = the k-th smallest of those per-warp minima: k DISTINCT rows
// have d_fp32 <= ub <= G, so a row with lb > G is strictly behind k rows and can never be in the top-k. Because the
// top-k rows of 10M almost surely sit in k different warps' slices (2368 warps), G converges to the true k-th best ub.
// Rows with lb <= G are appended to the CTA's candidate region (a burst while G is still +inf, then a trickle); at its
// end each CTA re-filters its region against the now almost final G, RESCORES the survivors (a handful per CTA, all
// CTAs in parallel) from the fp32 rows and appends them to one global list as exact (distance, id) keys. The last CTA
// adds the zero-norm ids and selects the top-k of that list (one rank sort for the usual ~100 keys; the CTA-shared
// streaming selector of the big-k scan for longer lists). One kernel launch per query.
//
// Anything the fast path cannot bound — zero-norm query, candidate overflow (adversarial order / massive near-ties) —
// raises the status word next to the result; the host then answers the query with scan_topk_kernel (still the GPU:
// there is no CPU fallback). tests/test_gpu_byte_prefilter.py asserts bit-equality with csgpu_search.
#pragma once
#include <cuda_fp16.h>

#include "scan.cuh"

namespace csgpu {

constexpr int I8_WARPS = 8;                          // streaming warps; warp I8_WARPS is the helper
constexpr int I8_THREADS = (I8_WARPS + 1) * 32;
constexpr uint32_t I8_REGION = 4096;                 // candidate slots per CTA
constexpr uint32_t I8_TAIL_CAP = 4096;               // keys the last CTA sorts (32 KB of shared memory)
constexpr uint32_t I8_MAX_K = 256;
constexpr uint32_t I8_CTA_CAP = 2048;                // survivors one CTA rescoring at its end can hold (shared memory)
constexpr uint32_t I8_MAX_WARPS = 148 * 2 * I8_WARPS;   // per-warp minima staged in shared memory by the helper
constexpr float I8_SLACK = 3e-6f;
constexpr int I8_CHUNK = 8;                          // row groups per grab of the work counter (shrinks near the end)
constexpr uint32_t I8_FINAL_CAP = 65536;             // slots of the global candidate list (the tail re-filters it when long)
constexpr int I8_HELPER_BATCH = 38;                  // warp-minima loads a helper lane keeps in flight (2 batches cover 2432)
constexpr int I8_SEARCH_BITS = 20;                   // bits of the k-th-smallest search (the rest is rounded up)

struct I8Args {
    const uint8_t *shadow;      // [n_rows][d8] int8
    const uint32_t *meta;       // [n_rows] half2: lo = s_r, hi = e_r
    const float4 *rows;         // [n_rows][dim4] fp32 unit rows (rescoring only)
    const uint32_t *ids;        // [n_rows]
    uint64_t n_rows;
    uint32_t dim4, d8;
    const float *q;             // [dim4*4] raw query (device)
    uint32_t k;
    const uint32_t *zero_ids;
    uint32_t n_zero;
    uint32_t *warp_min;         // [gridDim.x * I8_WARPS] okey(ub); 0xFFFFFFFF between launches
    uint64_t *region;           // [gridDim.x][I8_REGION]
    uint64_t *final_list;       // [I8_FINAL_CAP]
    unsigned *counters;         // [0] ticket, [1] final count, [2] overflow flag, [3] next row group — all 0 between launches;
                                // [4] device copy of status[0], rewritten by every launch (ScanArgs::run_if reads it)
    uint64_t *out_keys;         // [k] result keys
    unsigned long long *timing; // nullptr, or [gridDim.x + 1][4] globaltimer stamps (CSGPU_I8_TIMING=1: where the time goes)
    uint64_t *status;           // [0]: 0 = out_keys valid, 1 = answer with the exact scan instead;
                                // [1]: (fp32 rows rescored << 32) | candidates on the final list
    // FILT instantiation only (round 2): the same filters as the fp32 filtered scan (scan.cuh scan_rows_filtered) —
    // id-indexed allow bitmap (tags == nullptr) or row-tag predicate with an optional per-FILE bitmap (tags != nullptr)
    const uint32_t *tags = nullptr;
    const uint64_t *bitmap = nullptr;
    uint64_t n_bits = 0;
    uint32_t lang_mask = 0xFFFFFFFFu, file_lo = 0, file_hi = 0xFFFFFFFFu;
};

__device__ __forceinline__ uint4 ldg_stream_u4(const uint8_t *p)
{
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

__device__ __forceinline__ uint32_t pack_i8x4(int a, int b, int c, int d)
{
    return (uint32_t)(a & 0xFF) | ((uint32_t)(b & 0xFF) << 8) | ((uint32_t)(c & 0xFF) << 16) | ((uint32_t)(d & 0xFF) << 24);
}

// k-th smallest (1-based) of vals[0..n4*4) held in shared memory (padded with 0xFFFFFFFF), searched on the top
// I8_SEARCH_BITS bits and rounded UP (so the result is >= the true k-th smallest: still a valid bound). One warp.
__device__ __forceinline__ uint32_t warp_kth_smallest(const uint4 *vals4, uint32_t n4, uint32_t k, int lane,
                                                   const volatile unsigned *stop = nullptr)
{
    uint32_t ans = 0;
    for (int b = 31; b >= 32 - I8_SEARCH_BITS; --b) {
        if (stop != nullptr && *stop) return 0xFFFFFFFFu;   // the CTA is done: do not hold it up
        const uint32_t cand = ans | ((1u << b) - 1u);   // bit b clear, everything below set
        unsigned c = 0;
        for (uint32_t t = lane; t < n4; t += 32) {
            const uint4 v = vals4[t];
            c += (v.x <= cand ? 1u : 0u) + (v.y <= cand ? 1u : 0u) + (v.z <= cand ? 1u : 0u) + (v.w <= cand ? 1u : 0u);
        }
        c = __reduce_add_sync(FULL, c);
        if (c < k) ans |= 1u << b;
    }
    return ans | ((1u << (32 - I8_SEARCH_BITS)) - 1u);
}

// CTA-wide ascending sort of n (pow2) keys in shared memory. Small lists (the common case: a few hundred candidates)
// are sorted by rank — every element counts the keys below it and is scattered to that position: one pass of
// broadcast reads and two barriers instead of log^2(n) barrier-separated bitonic steps. Equal keys (only KEY_EMPTY
// padding) are ordered by position.
__device__ __forceinline__ void cta_sort_fast(uint64_t *s, uint32_t n)
{
    if (n > 512) { cta_sort(s, n); return; }
    uint64_t mine[2];
    uint32_t rank[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const uint32_t i = threadIdx.x + e * blockDim.x;
        mine[e] = KEY_EMPTY; rank[e] = 0;
        if (i < n) {
            const uint64_t v = s[i];
            uint32_t r = 0;
            for (uint32_t j = 0; j < n; ++j) {
                const uint64_t o = s[j];
                r += (o < v || (o == v && j < i)) ? 1u : 0u;
            }
            mine[e] = v; rank[e] = r;
        }
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 2; ++e)
        if (threadIdx.x + e * blockDim.x < n) s[rank[e]] = mine[e];
    __syncthreads();
}

__device__ __forceinline__ uint32_t ld_cg_u32(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// The fp32 query exactly as scan_topk_kernel holds it: lane owns float4 lane + 32 j, scaled by 1/||q||. Returns ss.
template <int V, bool EXACT>
__device__ __forceinline__ float load_unit_query(const float *q, uint32_t dim4, float4 (&qv)[V], int lane)
{
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) {
        const uint32_t c = lane + 32 * j;
        if (EXACT || c < dim4) qv[j] = reinterpret_cast<const float4 *>(q)[c];
        else qv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        ss = fmaf(qv[j].x, qv[j].x, ss); ss = fmaf(qv[j].y, qv[j].y, ss);
        ss = fmaf(qv[j].z, qv[j].z, ss); ss = fmaf(qv[j].w, qv[j].w, ss);
    }
    ss = warp_sum_tree(ss);
    if (ss > 0.f) {
        const float qinv = 1.0f / sqrtf(ss);
#pragma unroll
        for (int j = 0; j < V; ++j) { qv[j].x *= qinv; qv[j].y *= qinv; qv[j].z *= qinv; qv[j].w *= qinv; }
    }
    return ss;
}

// Exact keys for buf[0, n) in place: entries carry ROW indices in their low word and come back as (distance, chunk id)
// keys computed with scan_topk_kernel's arithmetic, operation for operation (see rescore.cuh: rescore_range). Streaming
// warps only; the caller synchronises.
template <int V, bool EXACT>
__device__ __forceinline__ void i8_rescore(const I8Args &a, const float4 (&qv)[V], uint64_t *buf, uint32_t n, int warp, int lane)
{
    constexpr int RR = V <= 4 ? 4 : (V <= 6 ? 2 : 1);
    if (warp >= I8_WARPS) return;
    for (uint32_t i0 = (uint32_t)warp * RR; i0 < n; i0 += I8_WARPS * RR) {
        uint32_t rowi[RR];
        bool lv[RR];
#pragma unroll
        for (int r = 0; r < RR; ++r) {
            lv[r] = i0 + r < n;
            rowi[r] = lv[r] ? (uint32_t)buf[i0 + r] : 0u;
        }
        __syncwarp();      // every lane has read its slots before lane 0 overwrites them below
        uint32_t idv[RR];  // chunk ids ride along with the row loads
#pragma unroll
        for (int r = 0; r < RR; ++r) idv[r] = lv[r] ? __ldg(a.ids + rowi[r]) : 0u;
        float4 xr[RR][V];
#pragma unroll
        for (int r = 0; r < RR; ++r) {
            const float4 *p = a.rows + (size_t)rowi[r] * a.dim4 + lane;
#pragma unroll
            for (int j = 0; j < V; ++j) {
                if (lv[r] && (EXACT || lane + 32 * j < a.dim4)) xr[r][j] = __ldg(p + 32 * j);
                else xr[r][j] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int r = 0; r < RR; ++r) {
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < V; ++j) {
                acc = fmaf(xr[r][j].x, qv[j].x, acc); acc = fmaf(xr[r][j].y, qv[j].y, acc);
                acc = fmaf(xr[r][j].z, qv[j].z, acc); acc = fmaf(xr[r][j].w, qv[j].w, acc);
            }
            acc = warp_sum_tree(acc);
            const float dist = fmaf(-0.5f, acc, 0.5f);
            if (lv[r] && lane == 0) buf[i0 + r] = make_key(dist, idv[r]);
        }
    }
}

// V = float4 per lane of the fp32 query (ceil(dim4/32)) = 128-byte lines per shadow row; EXACT: dim4 == 32 V;
// R = 4-row groups in flight per warp iteration (a warp streams 4 R rows = R x V x 512 B at a time).
// FILT (round 2): the byte prefilter under a filter. A row group (4 R consecutive rows) is only streamed if the filter
// allows at least one of its rows, and only allowed rows may publish a bound or become candidates — so, like the fp32
// filtered scan, masked rows cost no shadow bytes (a file / language predicate masks runs of ~37 chunks, i.e. whole
// groups), and G stays the k-th smallest upper bound of k distinct ALLOWED rows. The unit of work is a 32-row block (one
// filter word — tag or chunk id — per lane, loaded two blocks ahead of its use, the bitmap word one block ahead); the
// block's 32 / (4 R) row groups are streamed only where the filter allows a row (FILT needs 4 R | 32: R = 4 at dim 384).
template <int V, bool EXACT, int R, bool FILT = false>
__global__ void __launch_bounds__(I8_THREADS, 2) scan_i8_kernel(const I8Args a)
{
    constexpr int ROWS_PER_ITER = 4 * R;
    extern __shared__ __align__(16) uint64_t smem[];          // helper: per-warp minima (u32); tail: candidate keys
    __shared__ unsigned s_cnt, s_done, s_last, s_end, s_ns;
    __shared__ uint64_t s_thr;
    __shared__ volatile uint32_t s_G;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane & 7;
    const uint32_t k = a.k;
    if (threadIdx.x == 0) { s_cnt = 0; s_done = 0; s_ns = 0; s_G = 0xFFFFFFFFu; }
    if (a.timing && threadIdx.x == 0) a.timing[blockIdx.x * 4 + 0] = global_timer_ns();

    float s_q, EQ, QN;
    uint4 qw[V];
    {
    // ---- query -> registers, scaled to unit length: scan_topk_kernel's prologue, operation for operation ----
    float4 qv[V];
    const float ss = load_unit_query<V, EXACT>(a.q, a.dim4, qv, lane);
    if (!(ss > 0.f) || !(ss < 3.0e38f)) {   // zero-norm query (every distance is 0.0), or |q|^2 overflows fp32 (the scan
                                            // kernel then scores every row 0.5): the exact kernel answers it
        if (blockIdx.x == 0 && threadIdx.x == 0) { a.status[0] = 1ull; a.status[1] = 0ull; a.counters[4] = 1u; }
        return;
    }

    // ---- int8 image of the unit query + its error terms (every warp computes the same values) ----
    float amax = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) amax = fmaxf(amax, fmaxf(fmaxf(fabsf(qv[j].x), fabsf(qv[j].y)), fmaxf(fabsf(qv[j].z), fabsf(qv[j].w))));
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) amax = fmaxf(amax, __shfl_xor_sync(FULL, amax, m));
    s_q = amax * (1.0f / 127.0f);   // amax > 0 (ss > 0)
    const float inv_sq = 1.0f / s_q;
    uint32_t w[V];
    float res2 = 0.f;
    int qq = 0;
#pragma unroll
    for (int j = 0; j < V; ++j) {
        const int ix = max(-127, min(127, __float2int_rn(qv[j].x * inv_sq))), iy = max(-127, min(127, __float2int_rn(qv[j].y * inv_sq)));
        const int iz = max(-127, min(127, __float2int_rn(qv[j].z * inv_sq))), iw = max(-127, min(127, __float2int_rn(qv[j].w * inv_sq)));
        w[j] = pack_i8x4(ix, iy, iz, iw);
        const float rx = fmaf(-s_q, (float)ix, qv[j].x), ry = fmaf(-s_q, (float)iy, qv[j].y);
        const float rz = fmaf(-s_q, (float)iz, qv[j].z), rw = fmaf(-s_q, (float)iw, qv[j].w);
        res2 = fmaf(rx, rx, res2); res2 = fmaf(ry, ry, res2); res2 = fmaf(rz, rz, res2); res2 = fmaf(rw, rw, res2);
        qq += ix * ix + iy * iy + iz * iz + iw * iw;
    }
    res2 = warp_sum_tree(res2);
    qq = __reduce_add_sync(FULL, qq);
    // e_q (x 1 + 1e-3: fp32 summation error, ||x|| <= 1 + 1e-6) and ||q^|| (x 1 + 1e-5), both rounded generously up
    EQ = sqrtf(res2) * 1.001f + 1e-9f;
    QN = s_q * sqrtf((float)qq) * 1.00001f;
    // lane layout of the streamed rows: 8 lanes per row, lane `sub` owns bytes [128 i + 16 sub, +16) of every line i —
    // i.e. the packed words of lanes 4 sub .. 4 sub + 3 (float4 index lane + 32 j <-> byte 16 (lane + 32 j))
#pragma unroll
    for (int j = 0; j < V; ++j) {
        qw[j].x = __shfl_sync(FULL, w[j], 4 * sub + 0); qw[j].y = __shfl_sync(FULL, w[j], 4 * sub + 1);
        qw[j].z = __shfl_sync(FULL, w[j], 4 * sub + 2); qw[j].w = __shfl_sync(FULL, w[j], 4 * sub + 3);
    }
    }
    __syncthreads();

    const uint64_t n = a.n_rows;
    const uint32_t n_warps = gridDim.x * I8_WARPS;
    uint64_t *region = a.region + (size_t)blockIdx.x * I8_REGION;

    if (warp == I8_WARPS) {
        // ---- helper warp: G = k-th smallest of the per-warp minima, refreshed until the CTA's rows are done ----
        uint32_t *vals = reinterpret_cast<uint32_t *>(smem);
        const uint32_t n_pad = (n_warps + 127u) & ~127u;   // whole uint4 per lane in the search
        unsigned rounds = 0;
        while (*reinterpret_cast<volatile unsigned *>(&s_done) == 0) {
            // all of a lane's loads are issued before the first one is consumed: under the streaming load an L2 round
            // trip is 2-3 us, and ten dependent batches made one refresh ~25 us (short scans then ended with G still +inf)
            for (uint32_t t0 = 0; t0 < n_pad; t0 += 32 * I8_HELPER_BATCH) {
                uint32_t v[I8_HELPER_BATCH];
#pragma unroll
                for (int u = 0; u < I8_HELPER_BATCH; ++u) {
                    const uint32_t t = t0 + u * 32 + lane;
                    v[u] = t < n_warps ? ld_cg_u32(a.warp_min + t) : 0xFFFFFFFFu;
                }
#pragma unroll
                for (int u = 0; u < I8_HELPER_BATCH; ++u) {
                    const uint32_t t = t0 + u * 32 + lane;
                    if (t < n_pad) vals[t] = v[u];
                }
            }
            __syncwarp();
            // once a finite G is out, a refresh gives way as soon as the CTA's rows are done; the first one always completes
            const uint32_t g = warp_kth_smallest(reinterpret_cast<const uint4 *>(vals), n_pad / 4, k, lane,
                                                 s_G != 0xFFFFFFFFu ? &s_done : nullptr);
            if (lane == 0) atomicMin(const_cast<uint32_t *>(&s_G), g);
            __syncwarp();
            // G settles within the first few rows of every warp: refresh back to back at first, then back off
            // geometrically (4 us, 8 us, ... 64 us between refreshes) — a refresh is ~4k instructions on a scheduler the
            // streaming warps share, and a slightly stale G only lets a few more candidates through
            if (++rounds < 16) __nanosleep(100);
            else {
                const int naps = min(256, 16 << min(rounds - 16, 4u));
                for (int w = 0; w < naps && *reinterpret_cast<volatile unsigned *>(&s_done) == 0; ++w) __nanosleep(250);
            }
        }
    } else {
        // ---- streaming warps ----
        const uint64_t gw = (uint64_t)blockIdx.x * I8_WARPS + warp;
        const uint32_t n_groups = (uint32_t)((n + ROWS_PER_ITER - 1) / ROWS_PER_ITER);
        uint32_t wmin = 0xFFFFFFFFu;
        const float sq_scale = s_q;
        const int my_g = sub % R, my_r4 = lane >> 3;          // lanes with sub < R each finish one of the 4 R rows
        // Dynamic work distribution: SMs stream at visibly different rates (a static split leaves the first CTA idle
        // for the last 25 % of the kernel), so warps grab chunks of row groups from one global counter — up to
        // I8_CHUNK groups while plenty is left, single groups near the end. The next grab is issued before the
        // current chunk is processed, so its latency never shows.
        int patience = 64;   // x 500 ns per warp, in total
        if constexpr (FILT) {
            const bool tagmode = a.tags != nullptr, have_bm = a.bitmap != nullptr;
            static_assert(32 % ROWS_PER_ITER == 0, "FILT: a 32-row block must be a whole number of row groups");
            const uint64_t n_blocks = (n + 31) / 32;
            auto load_fw = [&](uint64_t blk, bool &valid) -> uint32_t {   // filter word of row blk * 32 + lane
                const uint64_t row = blk * 32 + lane;
                valid = blk < n_blocks && row < n;
                if (!valid) return 0u;
                if (!tagmode) return __ldg(a.ids + row);
                const uint32_t tag = __ldg(a.tags + row);
                valid = tag_pass_static(tag, a.lang_mask, a.file_lo, a.file_hi);
                return tag & 0x07FFFFFFu;
            };
            auto load_bw = [&](uint32_t id, bool valid) -> uint64_t {
                if (tagmode && !have_bm) return valid ? ~0ull : 0ull;
                return (valid && (uint64_t)id < a.n_bits) ? __ldg(reinterpret_cast<const unsigned long long *>(a.bitmap) + (id >> 6)) : 0ull;
            };
            BlockCursor<false> cursor;
            cursor.init(nullptr, a.counters + 3, nullptr, n_blocks, gw, n_warps, lane, warp, true);
            bool v_cur, v_nxt, v_nx2, l_cur, l_nxt, l_nx2;
            uint64_t g_cur = cursor.next(l_cur);
            uint32_t f_cur = load_fw(g_cur, v_cur);
            uint64_t w_cur = load_bw(f_cur, v_cur);
            uint64_t g_nxt = cursor.next(l_nxt);
            uint32_t f_nxt = load_fw(g_nxt, v_nxt);
            while (l_cur) {
                const uint64_t w_nxt = load_bw(f_nxt, v_nxt);
                const uint64_t g_nx2 = cursor.next(l_nx2);
                const uint32_t f_nx2 = load_fw(g_nx2, v_nx2);
                const unsigned am32 = __ballot_sync(FULL, (w_cur >> (f_cur & 63)) & 1ull);   // bit i: row 32 blk + i is allowed
#pragma unroll 1
                for (int w = 0; w < 32 / ROWS_PER_ITER; ++w) {
                    const unsigned am = ROWS_PER_ITER == 32 ? am32 : ((am32 >> (w * ROWS_PER_ITER)) & ((1u << (ROWS_PER_ITER & 31)) - 1u));
                    if (!am) continue;   // bit i: row base + i is allowed
                    while (patience > 0 && s_G == 0xFFFFFFFFu && *reinterpret_cast<volatile unsigned *>(&s_cnt) > I8_REGION / 4) {
                        __nanosleep(500);
                        --patience;
                    }
                    const uint64_t base = g_cur * 32 + (uint64_t)w * ROWS_PER_ITER;
                    uint4 x[R][V];
#pragma unroll
                    for (int g = 0; g < R; ++g) {
                        const uint64_t row = base + g * 4 + (lane >> 3);
                        const uint8_t *p = a.shadow + row * a.d8 + sub * 16;
                        const bool on = row < n && ((am >> (g * 4 + (lane >> 3))) & 1u);
#pragma unroll
                        for (int j = 0; j < V; ++j) {
                            if (on) x[g][j] = ldg_stream_u4(p + 128 * j);
                            else x[g][j] = make_uint4(0u, 0u, 0u, 0u);
                        }
                    }
                    uint32_t mt = 0;
                    if (lane < ROWS_PER_ITER && base + lane < n) mt = __ldg(a.meta + base + lane);
                    const uint32_t G = s_G;
                    int mine = 0;
#pragma unroll
                    for (int g = 0; g < R; ++g) {
                        int acc = 0;
#pragma unroll
                        for (int j = 0; j < V; ++j) {
                            acc = __dp4a((int)x[g][j].x, (int)qw[j].x, acc); acc = __dp4a((int)x[g][j].y, (int)qw[j].y, acc);
                            acc = __dp4a((int)x[g][j].z, (int)qw[j].z, acc); acc = __dp4a((int)x[g][j].w, (int)qw[j].w, acc);
                        }
                        acc += __shfl_xor_sync(FULL, acc, 4);
                        acc += __shfl_xor_sync(FULL, acc, 2);
                        acc += __shfl_xor_sync(FULL, acc, 1);
                        if (my_g == g) mine = acc;
                    }
                    const uint32_t m = __shfl_sync(FULL, mt, 4 * my_g + my_r4);
                    const uint64_t row = base + 4 * my_g + my_r4;
                    const bool live = sub < R && row < n && ((am >> (4 * my_g + my_r4)) & 1u);
                    const float s_r = __half2float(__ushort_as_half((unsigned short)(m & 0xFFFFu)));
                    const float e_r = __half2float(__ushort_as_half((unsigned short)(m >> 16)));
                    const float c_hat = (float)mine * (sq_scale * s_r);
                    const float M = fmaf(QN, e_r, EQ);
                    const float lb = fmaf(-0.5f, c_hat + M, 0.5f) - I8_SLACK;
                    const float ub = fmaf(-0.5f, c_hat - M, 0.5f) + I8_SLACK;
                    const uint32_t lbk = okey(lb), ubk = live ? okey(ub) : 0xFFFFFFFFu;
                    const bool hit = live && lbk <= G;
                    const unsigned hm = __ballot_sync(FULL, hit);
                    if (hm) {
                        unsigned pos0 = 0;
                        if (lane == 0) pos0 = atomicAdd(&s_cnt, (unsigned)__popc(hm));
                        pos0 = __shfl_sync(FULL, pos0, 0);
                        if (hit) {
                            const unsigned pos = pos0 + __popc(hm & ((1u << lane) - 1u));
                            if (pos < I8_REGION) region[pos] = ((uint64_t)lbk << 32) | (uint32_t)row;
                        }
                    }
                    if (__any_sync(FULL, ubk < wmin)) {
                        wmin = min(wmin, __reduce_min_sync(FULL, ubk));
                        if (lane == 0) *reinterpret_cast<volatile uint32_t *>(a.warp_min + gw) = wmin;
                    }
                }
                g_cur = g_nxt; l_cur = l_nxt; f_cur = f_nxt; v_cur = v_nxt; w_cur = w_nxt;
                g_nxt = g_nx2; l_nxt = l_nx2; f_nxt = f_nx2; v_nxt = v_nx2;
            }
        } else {
        uint32_t c_next = max(1u, min((uint32_t)I8_CHUNK, n_groups / (2u * n_warps))), nxt = 0;
        if (lane == 0) nxt = atomicAdd(a.counters + 3, c_next);
        for (;;) {
        const uint32_t c_start = __shfl_sync(FULL, nxt, 0);
        if (c_start >= n_groups) break;
        // While G is still +inf every row is a candidate. If the grid starts staggered (first launch, SMs busy with
        // another stream) the CTAs that run first must not swallow chunk after chunk in that mode and overflow their
        // region: with the region a quarter full they wait — briefly, bounded — for the threshold to come alive.
        while (patience > 0 && s_G == 0xFFFFFFFFu && *reinterpret_cast<volatile unsigned *>(&s_cnt) > I8_REGION / 4) {
            __nanosleep(500);
            --patience;
        }
        const uint32_t c_end = min(c_start + c_next, n_groups);
        c_next = max(1u, min((uint32_t)I8_CHUNK, (n_groups - c_start) / (2u * n_warps)));
        if (lane == 0) nxt = atomicAdd(a.counters + 3, c_next);
        for (uint32_t grp = c_start; grp < c_end; ++grp) {
            const uint64_t base = (uint64_t)grp * ROWS_PER_ITER;
            uint4 x[R][V];
#pragma unroll
            for (int g = 0; g < R; ++g) {
                const uint64_t row = base + g * 4 + (lane >> 3);
                const uint8_t *p = a.shadow + row * a.d8 + sub * 16;
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    if (row < n) x[g][j] = ldg_stream_u4(p + 128 * j);
                    else x[g][j] = make_uint4(0u, 0u, 0u, 0u);
                }
            }
            uint32_t mt = 0;
            if (lane < ROWS_PER_ITER && base + lane < n) mt = __ldg(a.meta + base + lane);
            const uint32_t G = s_G;
            int mine = 0;
#pragma unroll
            for (int g = 0; g < R; ++g) {
                int acc = 0;
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    acc = __dp4a((int)x[g][j].x, (int)qw[j].x, acc); acc = __dp4a((int)x[g][j].y, (int)qw[j].y, acc);
                    acc = __dp4a((int)x[g][j].z, (int)qw[j].z, acc); acc = __dp4a((int)x[g][j].w, (int)qw[j].w, acc);
                }
                acc += __shfl_xor_sync(FULL, acc, 4);
                acc += __shfl_xor_sync(FULL, acc, 2);
                acc += __shfl_xor_sync(FULL, acc, 1);
                if (my_g == g) mine = acc;
            }
            // lane (r4, sub < R) now owns row base + 4 sub + r4; its meta word sits in lane 4 sub + r4
            const uint32_t m = __shfl_sync(FULL, mt, 4 * my_g + my_r4);
            const uint64_t row = base + 4 * my_g + my_r4;
            const bool live = sub < R && row < n;
            const float s_r = __half2float(__ushort_as_half((unsigned short)(m & 0xFFFFu)));
            const float e_r = __half2float(__ushort_as_half((unsigned short)(m >> 16)));
            const float c_hat = (float)mine * (sq_scale * s_r);
            const float M = fmaf(QN, e_r, EQ);
            const float lb = fmaf(-0.5f, c_hat + M, 0.5f) - I8_SLACK;
            const float ub = fmaf(-0.5f, c_hat - M, 0.5f) + I8_SLACK;
            const uint32_t lbk = okey(lb), ubk = live ? okey(ub) : 0xFFFFFFFFu;
            const bool hit = live && lbk <= G;
            const unsigned hm = __ballot_sync(FULL, hit);
            if (hm) {   // warp-uniform, rare once G is live
                unsigned pos0 = 0;
                if (lane == 0) pos0 = atomicAdd(&s_cnt, (unsigned)__popc(hm));
                pos0 = __shfl_sync(FULL, pos0, 0);
                if (hit) {
                    const unsigned pos = pos0 + __popc(hm & ((1u << lane) - 1u));
                    if (pos < I8_REGION) region[pos] = ((uint64_t)lbk << 32) | (uint32_t)row;
                }
            }
            if (__any_sync(FULL, ubk < wmin)) {   // rare: ~ln(rows per warp) times
                wmin = min(wmin, __reduce_min_sync(FULL, ubk));
                if (lane == 0) *reinterpret_cast<volatile uint32_t *>(a.warp_min + gw) = wmin;
            }
        }
        }
        }
        asm volatile("bar.sync 1, %0;" :: "n"(I8_WARPS * 32));
        if (threadIdx.x == 0) atomicExch(&s_done, 1u);
    }
    __syncthreads();
    if (a.timing && threadIdx.x == 0) a.timing[blockIdx.x * 4 + 1] = global_timer_ns();

    // ---- CTA end: re-filter the region against the (almost final) G; the survivors (a handful per CTA) are rescored
    //      HERE, by every CTA for its own rows in parallel, and go to the global list as exact keys ----
    uint64_t *S = smem + I8_TAIL_CAP - I8_CTA_CAP;   // behind the helper's staging area
    if (s_G == 0xFFFFFFFFu) {   // CTA-uniform (read after the barrier)
        // A short scan can end before the helper has produced a finite G. Every warp of the grid that had rows has
        // published its minimum by now, so the whole CTA fetches them in one round trip and warp 0 searches.
        uint32_t *vals = reinterpret_cast<uint32_t *>(smem);
        const uint32_t n_pad = (n_warps + 127u) & ~127u;
        for (uint32_t t = threadIdx.x; t < n_pad; t += blockDim.x) vals[t] = t < n_warps ? ld_cg_u32(a.warp_min + t) : 0xFFFFFFFFu;
        __syncthreads();
        if (warp == 0) {
            const uint32_t g = warp_kth_smallest(reinterpret_cast<const uint4 *>(vals), n_pad / 4, k, lane);
            if (lane == 0) s_G = g;
        }
        __syncthreads();
    }
    {
        const unsigned cnt = s_cnt;
        if (cnt > I8_REGION && threadIdx.x == 0) atomicExch(a.counters + 2, 1u);
        const unsigned nreg = min(cnt, I8_REGION);
        const uint32_t G = s_G;
        for (unsigned t = threadIdx.x; t < nreg; t += blockDim.x) {
            const uint64_t key = region[t];
            if ((uint32_t)(key >> 32) <= G) {
                const unsigned pos = atomicAdd(&s_ns, 1u);
                if (pos < I8_CTA_CAP) S[pos] = key;
            }
        }
    }
    __syncthreads();
    {
        unsigned n_s = s_ns;   // CTA-uniform
        if (n_s > I8_CTA_CAP) { if (threadIdx.x == 0) atomicExch(a.counters + 2, 1u); n_s = I8_CTA_CAP; }
        if (n_s) {
            float4 qv[V];   // the unit query again, for the exact arithmetic (not kept live through the streaming loop)
            load_unit_query<V, EXACT>(a.q, a.dim4, qv, lane);
            i8_rescore<V, EXACT>(a, qv, S, n_s, warp, lane);
            __syncthreads();
            if (threadIdx.x == 0) s_end = atomicAdd(a.counters + 1, n_s);
            __syncthreads();
            const unsigned base = s_end;
            for (unsigned t = threadIdx.x; t < n_s; t += blockDim.x)
                if (base + t < I8_FINAL_CAP) a.final_list[base + t] = S[t];
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (a.timing) { a.timing[blockIdx.x * 4 + 2] = global_timer_ns(); a.timing[blockIdx.x * 4 + 3] = ((unsigned long long)s_ns << 32) | s_G; }
        const unsigned t = atomicAdd(a.counters + 0, 1u);
        s_last = (t == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();

    // ---- last CTA: the global list holds exact (distance, id) keys; add the zero-norm ids and select ----
    uint64_t *C = smem;
    const unsigned total = *reinterpret_cast<volatile unsigned *>(a.counters + 1);
    const bool overflow = total > I8_FINAL_CAP || *reinterpret_cast<volatile unsigned *>(a.counters + 2) != 0;
    __syncthreads();
    // leave the scratch clean for the next launch (all other CTAs have retired their use of it)
    for (uint32_t t = threadIdx.x; t < n_warps; t += blockDim.x) a.warp_min[t] = 0xFFFFFFFFu;
    if (threadIdx.x == 0) { a.counters[0] = 0; a.counters[1] = 0; a.counters[2] = 0; a.counters[3] = 0; }
    if (overflow) {
        if (threadIdx.x == 0) { a.status[0] = 1ull; a.status[1] = (uint64_t)min(total, 0xFFFFFFFFu); a.counters[4] = 1u; }
        return;
    }
    const volatile uint64_t *fl = a.final_list;
    // zero-norm rows: distance 0.0 (arroy pn*qn == 0), ascending id; the first k suffice — under a filter the first k ALLOWED
    // ones, so every entry of the side list is looked at (disallowed ones become empty slots)
    const uint32_t nz = FILT ? a.n_zero : min(a.n_zero, k);
    auto zero_key = [&](uint32_t i) -> uint64_t {
        if constexpr (FILT) {
            if (!zero_row_allowed(a.bitmap, a.n_bits, a.tags, a.lang_mask, a.file_lo, a.file_hi, a.zero_ids, a.n_zero, i)) return KEY_EMPTY;
        }
        return make_key(0.f, a.zero_ids[i]);
    };
    if (total + nz <= 512) {
        // the common case (~115 keys at k = 10): one rank sort
        const uint32_t n_all = total + nz, fpad = pow2_at_least(n_all, 32);
        for (uint32_t t = threadIdx.x; t < fpad; t += blockDim.x)
            C[t] = t < total ? fl[t] : (t < n_all ? zero_key(t - total) : KEY_EMPTY);
        __syncthreads();
        cta_sort_fast(C, fpad);
        for (uint32_t j = threadIdx.x; j < k; j += blockDim.x) a.out_keys[j] = j < fpad ? C[j] : KEY_EMPTY;
    } else {
        // any number of keys through the CTA-shared streaming selector of the big-k scan (topk.cuh)
        CtaSel sel;
        sel.cb.buf = C; sel.cb.cnt = &s_cnt; sel.cb.thr = &s_thr; sel.cap = ctabuf_cap(k); sel.k = k;
        sel.reset();
        cta_buf_stream(sel.cb, sel.cap, k, total, [&](uint64_t t) { return fl[t]; });
        if (nz) cta_buf_stream(sel.cb, sel.cap, k, nz, [&](uint64_t t) { return zero_key((uint32_t)t); });
        sel.finish();
        for (uint32_t j = threadIdx.x; j < k; j += blockDim.x) a.out_keys[j] = C[j];
    }
    if (threadIdx.x == 0) {
        a.status[0] = 0ull; a.status[1] = ((uint64_t)total << 32) | total; a.counters[4] = 0u;
        if (a.timing) { a.timing[gridDim.x * 4 + 0] = global_timer_ns(); a.timing[gridDim.x * 4 + 1] = blockIdx.x; }
    }
}

// fp32 unit rows -> int8 shadow + meta (one warp per row). s_r = max|x| / 127 rounded UP to half (so |x / s_r| <= 127),
// xi = rint(x / s_r), e_r = ||x - s_r xi||_2 accumulated in f64 and rounded UP to half.
static __global__ void shadow_i8_from_rows_kernel(const float4 *__restrict__ rows, uint32_t *__restrict__ shadow_words,
                                                  uint32_t *__restrict__ meta, uint64_t first, uint64_t n, uint32_t dim4, uint32_t d8)
{
    const int lane = threadIdx.x & 31;
    const uint64_t w0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint32_t words = d8 / 4;
    for (uint64_t i = w0; i < n; i += nw) {
        const uint64_t row = first + i;
        const float4 *p = rows + row * dim4;
        float amax = 0.f;
        for (uint32_t c = lane; c < dim4; c += 32) {
            const float4 v = p[c];
            amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
        }
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xFFFFFFFFu, amax, m));
        __half sh = __float2half_ru(amax * (1.0f / 127.0f) * 1.0000002f);
        if (__half2float(sh) < 6.2e-5f) sh = __float2half_ru(6.2e-5f);   // keep the scale a normal half
        const float sf = __half2float(sh), inv = 1.0f / sf;
        double e2 = 0.0;
        for (uint32_t c = lane; c < words; c += 32) {
            uint32_t wv = 0;
            if (c < dim4) {
                const float4 v = p[c];
                const int ix = max(-127, min(127, __float2int_rn(v.x * inv))), iy = max(-127, min(127, __float2int_rn(v.y * inv)));
                const int iz = max(-127, min(127, __float2int_rn(v.z * inv))), iw = max(-127, min(127, __float2int_rn(v.w * inv)));
                wv = pack_i8x4(ix, iy, iz, iw);
                const double rx = (double)v.x - (double)sf * ix, ry = (double)v.y - (double)sf * iy;
                const double rz = (double)v.z - (double)sf * iz, rw = (double)v.w - (double)sf * iw;
                e2 += rx * rx + ry * ry + rz * rz + rw * rw;
            }
            shadow_words[row * words + c] = wv;
        }
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) e2 += __shfl_xor_sync(0xFFFFFFFFu, e2, m);
        if (lane == 0) {
            const float ef = (float)(sqrt(e2) * 1.000001) + 1e-12f;
            const __half eh = __float2half_ru(ef * 1.0000002f);
            meta[row] = ((uint32_t)__half_as_ushort(eh) << 16) | (uint32_t)__half_as_ushort(sh);
        }
    }
}

}  // namespace csgpu
