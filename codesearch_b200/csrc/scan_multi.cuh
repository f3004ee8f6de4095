// scan_multi.cuh — one pass over the corpus for MQ (4 or 8) queries at once.
//
// Serves the caller pattern at /root/reference/src/search/mod.rs:508-511 (<= 9 query-variant
// embeddings searched with the same limit) and csgpu_search_batch in general: the corpus is read
// from HBM ONCE per MQ queries, so queries/s scales ~MQ x while the scan stays HBM-bound
// (fp32 machine balance ~11 flop/B vs MQ/2 flop/B here).
//
// Arithmetic is bit-identical to scan.cuh: every (row, query) dot product uses the same per-lane
// FMA chain (j outer, xyzw inner) and the same pairing tree (offsets 16,8,4,2,1); the tree is just
// evaluated "transposed" — at each level a lane keeps half of its values and ships the other half
// to its partner — so R*MQ sums cost R*MQ-1 shuffles instead of 5*R*MQ, and every lane ends up
// owning exactly one (row, query) result: lane l -> row l/MQ, query l%MQ (MQ=8, R=4).
//
// Selection: per warp, per query, a sorted list of kpad keys in shared memory + a 32-entry pending
// buffer (also smem); merged in registers (E = kpad/32 keys per lane, in place). k <= 256.
#pragma once
#include "scan.cuh"

namespace csgpu {

struct MultiArgs {
    const float4 *rows;
    const uint32_t *ids;
    uint64_t n_rows;
    uint32_t dim4;
    const float *q;        // [nq, dim4*4] device, raw
    uint32_t nq;           // active queries (<= MQ)
    uint32_t k, kpad;
    const uint64_t *bitmap;
    uint64_t n_bits;
    const uint32_t *zero_ids;
    uint32_t n_zero;
    uint64_t *cand;        // [MQ][gridDim.x][k]
    unsigned *ticket;      // [0] arrival ticket, [1] work counter of the dynamic row split, [2] finishers done (all self-resetting)
    uint64_t *out_keys;    // [nq][k]
    uint32_t static_split = 0;   // 1 = fixed-stride row split (CSGPU_SCAN_STATIC=1, A/B runs)
    // row-tag predicate (csgpu_predicate_t; csgpu_search_variants_tagged): tags != nullptr switches it on, `bitmap` is then
    // the optional per-FILE bitmap. Evaluated on the few (row, query) results that beat a threshold — every row is still
    // read, which is the right trade for a handful of variants over a small corpus (the host routes by corpus size).
    const uint32_t *tags = nullptr;
    uint32_t lang_mask = 0, file_lo = 0, file_hi = 0;
    // in-kernel device time instead of CUDA events (see ScanArgs): CTA 0 stamps *t0_slot at its start, the last finisher
    // stores now - t0 (ns) into *elapsed_out (mapped host memory)
    unsigned long long *t0_slot = nullptr, *elapsed_out = nullptr;
};

// is row `row` (chunk id `id`) allowed under the launch's filter: id bitmap, or row-tag predicate (+ per-file bitmap)
__device__ __forceinline__ bool multi_row_allowed(const MultiArgs &a, uint64_t row, uint32_t id)
{
    if (a.tags == nullptr) return id_allowed(a.bitmap, a.n_bits, id);
    const uint32_t tag = __ldg(a.tags + row);
    if (!tag_pass_static(tag, a.lang_mask, a.file_lo, a.file_hi)) return false;
    return id_allowed(a.bitmap, a.n_bits, tag & 0x07FFFFFFu);
}

template <int NV, int O>
__device__ __forceinline__ void butterfly_level(float (&v)[32], int lane)
{
    if constexpr (NV > 1) {
        constexpr int H = NV / 2;
        const bool up = (lane & O) != 0;
#pragma unroll
        for (int i = 0; i < H; ++i) {
            const float send = up ? v[i] : v[i + H];
            const float keep = up ? v[i + H] : v[i];
            v[i] = keep + __shfl_xor_sync(FULL, send, O);
        }
    } else {
        v[0] += __shfl_xor_sync(FULL, v[0], O);
    }
}

// Per-warp, per-query selection state in shared memory.
template <int E>
struct MultiSel {
    static constexpr uint32_t KPAD = 32u * E;
    uint64_t *list;    // [MQ][KPAD]
    uint64_t *pend;    // [MQ][32]
    uint32_t *npend;   // [MQ]
    uint32_t km1;

    __device__ __forceinline__ uint64_t thr_of(int b) const { return list[(size_t)b * KPAD + km1]; }

    __device__ __forceinline__ void init(uint64_t *l, uint64_t *p, uint32_t *np, uint32_t k, int mq, int lane)
    {
        list = l; pend = p; npend = np; km1 = k - 1;
        for (int j = lane; j < mq * (int)KPAD; j += 32) list[j] = KEY_EMPTY;
        if (lane < mq) npend[lane] = 0;
        __syncwarp();
    }
    // warp-uniform (b, key)
    __device__ __forceinline__ void append(int b, uint64_t key, int lane)
    {
        const uint32_t np = npend[b];
        __syncwarp();
        if (lane == 0) { pend[b * 32 + np] = key; npend[b] = np + 1; }
        __syncwarp();
        if (np + 1 == 32) flush(b, lane);
    }
    __device__ __noinline__ void flush(int b, int lane)
    {
        const uint32_t np = npend[b];
        if (np == 0) return;
        uint64_t *L = list + (size_t)b * KPAD;
        uint64_t p = ((uint32_t)lane < np) ? pend[b * 32 + lane] : KEY_EMPTY;
        p = warp_sort32(p, lane);
        uint64_t l[E];
        uint32_t pos[E];
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const uint32_t j = lane + 32 * e;
            l[e] = L[j];
            int c = 0;
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) {
                const uint64_t pv = shfl64(p, c + s - 1);
                if (pv < l[e]) c += s;
            }
            const uint64_t pv = shfl64(p, c);
            if (pv < l[e]) c += 1;
            pos[e] = j + (uint32_t)c;
        }
        uint32_t c = 0;
#pragma unroll
        for (uint32_t s = KPAD >> 1; s > 0; s >>= 1)
            if (L[c + s - 1] <= p) c += s;
        if (L[c] <= p) c += 1;
        const uint32_t ppos = (uint32_t)lane + c;
        __syncwarp();
#pragma unroll
        for (int e = 0; e < E; ++e)
            if (pos[e] < KPAD) L[pos[e]] = l[e];
        if (ppos < KPAD) L[ppos] = p;
        if (lane == 0) npend[b] = 0;
        __syncwarp();
    }
};

// smem per CTA: queries MQT*dim4 float4 | per warp: MQT*KPAD + MQT*32 keys + MQT u32 | sort buffer WARPS*KPAD keys
template <int E>
__host__ __device__ constexpr size_t multi_warp_keys(int mq) { return (size_t)mq * (32 * E) + (size_t)mq * 32; }

// NG (round 2): query groups per pass. The rows a warp has loaded are scored against NG groups of MQ queries one after the
// other (same registers, same per-(row, query) arithmetic), so 9..16 queries cost ONE pass over HBM and twice the FMA
// work, at the same LDS : FFMA ratio as the 8-query kernel. (A 16-query x 2-row variant was built first and measured no
// faster than two 8-query passes — 5.4 ms: each query float4 read from shared memory fed only 8 FMAs.)
template <int V, int R, int MQ, int NG, int E>
__global__ void __launch_bounds__(SCAN_THREADS, (E <= 4) ? 2 : 1) scan_multi_topk_kernel(const MultiArgs a)
{
    constexpr bool EXACT = true;   // dim % 128 == 0 only (other dims: per-query loop of scan.cuh)
    constexpr int MQT = MQ * NG;   // queries per pass
    constexpr int NV = R * MQ;     // 8, 16 or 32 results per warp iteration and query group
    constexpr int SH = (NV == 32) ? 0 : (NV == 16 ? 1 : 2);
    static_assert(NV == 32 || NV == 16 || NV == 8, "R*MQ must be 8, 16 or 32");
    constexpr uint32_t KPAD = 32u * E;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned s_ticket;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t dim4 = a.dim4;
    if (a.t0_slot != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *a.t0_slot = global_timer_ns();

    float4 *qs = reinterpret_cast<float4 *>(smem_raw);                                   // [MQT][dim4]
    uint64_t *keys0 = reinterpret_cast<uint64_t *>(smem_raw + (size_t)MQT * dim4 * sizeof(float4));
    const size_t wkeys = multi_warp_keys<E>(MQT);
    uint64_t *wbase = keys0 + (size_t)warp * wkeys;
    uint64_t *sortbuf = keys0 + (size_t)SCAN_WARPS * wkeys;                              // [WARPS*KPAD]
    uint32_t *np_all = reinterpret_cast<uint32_t *>(sortbuf + (size_t)SCAN_WARPS * KPAD); // [WARPS][MQT]
    __shared__ float qflag[MQT];  // 1.0f if the query has zero norm (distance 0.0 everywhere)

    // ---- queries -> smem, unit-normalised exactly as scan.cuh does it (warp w handles query w) ----
    for (int b = warp; b < MQT; b += SCAN_WARPS) {
        float4 t[V];
        float ss = 0.f;
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const uint32_t c = lane + 32 * j;
            if ((uint32_t)b < a.nq && (EXACT || c < dim4)) t[j] = reinterpret_cast<const float4 *>(a.q)[(size_t)b * dim4 + c];
            else t[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            ss = fmaf(t[j].x, t[j].x, ss); ss = fmaf(t[j].y, t[j].y, ss);
            ss = fmaf(t[j].z, t[j].z, ss); ss = fmaf(t[j].w, t[j].w, ss);
        }
        ss = warp_sum_tree(ss);
        const bool qzero = !(ss > 0.f);
        const float qinv = qzero ? 0.f : 1.0f / sqrtf(ss);
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const uint32_t c = lane + 32 * j;
            if (EXACT || c < dim4) qs[(size_t)b * dim4 + c] = make_float4(t[j].x * qinv, t[j].y * qinv, t[j].z * qinv, t[j].w * qinv);
        }
        if (lane == 0) qflag[b] = qzero ? 1.f : 0.f;
    }
    MultiSel<E> sel;
    sel.init(wbase, wbase + (size_t)MQT * KPAD, np_all + warp * MQT, a.k, MQT, lane);
    __syncthreads();

    // lane -> (row slot, query) ownership after the butterfly, one query per group
    const int own = lane >> SH;
    const bool owner = (lane & ((1 << SH) - 1)) == 0;
    const int my_b = own % MQ, my_r = own / MQ;
    bool my_active[NG], my_qzero[NG];
    uint64_t thr[NG];
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        my_active[g] = owner && (uint32_t)(my_b + g * MQ) < a.nq;
        my_qzero[g] = qflag[my_b + g * MQ] != 0.f;
        thr[g] = my_active[g] ? KEY_EMPTY : 0ull;  // inactive lanes never pass
    }

    const uint64_t n = a.n_rows;
    const uint64_t gw = (uint64_t)blockIdx.x * SCAN_WARPS + warp;
    const uint64_t stride = (uint64_t)gridDim.x * SCAN_WARPS * R;
    auto scan_group = [&](uint64_t base) {
        float4 x[R][V];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const uint64_t row = base + r;
            const float4 *p = a.rows + row * dim4 + lane;
#pragma unroll
            for (int j = 0; j < V; ++j) {
                if (row < n && (EXACT || lane + 32 * j < dim4)) x[r][j] = ldg_stream<0>(p + 32 * j);
                else x[r][j] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            if (g > 0 && (uint32_t)(g * MQ) >= a.nq) break;   // warp-uniform: no query in this group
            float v[32];
#pragma unroll
            for (int i = 0; i < NV; ++i) v[i] = 0.f;
#pragma unroll
            for (int j = 0; j < V; ++j) {
#pragma unroll
                for (int b = 0; b < MQ; ++b) {
                    float4 qq = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (EXACT || lane + 32 * j < dim4) qq = qs[(size_t)(g * MQ + b) * dim4 + lane + 32 * j];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        float acc = v[r * MQ + b];
                        acc = fmaf(x[r][j].x, qq.x, acc); acc = fmaf(x[r][j].y, qq.y, acc);
                        acc = fmaf(x[r][j].z, qq.z, acc); acc = fmaf(x[r][j].w, qq.w, acc);
                        v[r * MQ + b] = acc;
                    }
                }
            }
            butterfly_level<NV, 16>(v, lane);
            butterfly_level<(NV / 2 > 0 ? NV / 2 : 1), 8>(v, lane);
            butterfly_level<(NV / 4 > 0 ? NV / 4 : 1), 4>(v, lane);
            butterfly_level<(NV / 8 > 0 ? NV / 8 : 1), 2>(v, lane);
            butterfly_level<(NV / 16 > 0 ? NV / 16 : 1), 1>(v, lane);
            const float dist = my_qzero[g] ? 0.f : fmaf(-0.5f, v[0], 0.5f);
            const uint64_t row = base + my_r;
            uint64_t key = KEY_EMPTY;
            if (my_active[g] && row < n && okey(dist) <= (uint32_t)(thr[g] >> 32)) {
                const uint32_t id = a.ids[row];
                const uint64_t kk = make_key(dist, id);
                if (kk < thr[g] && multi_row_allowed(a, row, id)) key = kk;
            }
            unsigned m = __ballot_sync(FULL, key != KEY_EMPTY);
            if (m) {
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const uint64_t kk = shfl64(key, src);
                    const int b = (src >> SH) % MQ + g * MQ;
                    if (kk < sel.thr_of(b)) sel.append(b, kk, lane);
                }
                if (my_active[g]) thr[g] = sel.thr_of(my_b + g * MQ);
            }
        }
    };
    if (a.static_split) {
        for (uint64_t base = gw * R; base < n; base += stride) scan_group(base);
    } else {
        // dynamic row split (see scan.cuh DYN): per-warp grabs of up to SCAN_CHUNK groups of R rows from a global counter
        const uint32_t n_grp = (uint32_t)((n + R - 1) / R), nw2 = 2u * gridDim.x * SCAN_WARPS;
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
    }

    // ---- per query: CTA top-k -> cand[b][cta][k]; the last CTAs to arrive merge across CTAs, one query each ----------------
    // k <= 32 (round 2): every warp's list of a query is 32 sorted keys, so warp w merges query w's (and w + 8's) eight lists
    // with seven 5-stage shuffle merges — all queries at once, two barriers — instead of one 256-key bitonic sort (36
    // barrier-separated stages) per query: that sort was ~3 us per query in EVERY CTA, 25 us of an 8-query pass.
    static_assert(E == 1, "per-warp sorted lists of 32 keys (k <= 32); larger k: scan_multi_cta_topk_kernel");
    for (int b = 0; b < (int)a.nq; ++b) sel.flush(b, lane);
    __syncthreads();
    for (int b = warp; b < (int)a.nq; b += SCAN_WARPS) {
        uint64_t run = keys0[(size_t)b * KPAD + lane];
#pragma unroll 1
        for (int w2 = 1; w2 < SCAN_WARPS; ++w2) run = warp_merge_low32(run, keys0[(size_t)w2 * wkeys + (size_t)b * KPAD + lane], lane);
        if ((uint32_t)lane < a.k) a.cand[((size_t)b * gridDim.x + blockIdx.x) * a.k + lane] = run;
    }
    __threadfence();
    __syncthreads();

    // The LAST nq CTAs to arrive each finish one query (round 2; one last CTA used to finish them one after the other,
    // ~6 us each — most of a small corpus's latency). A finisher waits until every CTA has delivered its lists: the CTAs
    // still running are resident or become resident as the others exit, and at most nq <= 16 CTAs ever wait, so the spin
    // cannot starve them. The last finisher to leave resets the counters for the next launch.
    if (threadIdx.x == 0) s_ticket = atomicAdd(a.ticket, 1u);
    __syncthreads();
    const unsigned n_fin = min(a.nq, gridDim.x), first_fin = gridDim.x - n_fin;
    if (s_ticket < first_fin) return;
    const unsigned fin = s_ticket - first_fin;
    if (threadIdx.x == 0) {
        const volatile unsigned *tk = a.ticket;
        while (*tk < gridDim.x) {}
    }
    __syncthreads();
    __threadfence();

    // Per query the scan kernel's tail (cta_topk_of_lists32: threshold = k-th smallest of the CTAs' minima, only the
    // survivors are sorted). Zero-norm rows (distance 0.0, ascending id; the first k allowed ones suffice) are the same for
    // every query. Scratch: the warps' selection state, no longer needed.
    const uint64_t total = (uint64_t)gridDim.x * a.k;
    uint64_t zrun = KEY_EMPTY;
    if (a.n_zero) {
        uint32_t found = 0;
        for (uint32_t o = 0; o < a.n_zero && found < a.k; o += SCAN_WARPS * 32) {   // CTA-uniform loop; warp w sorts its 32 ids of the batch (the lists are merged across warps below)
            uint64_t key = KEY_EMPTY;
            const uint32_t i = o + threadIdx.x;
            if (i < a.n_zero && zero_row_allowed(a.bitmap, a.n_bits, a.tags, a.lang_mask, a.file_lo, a.file_hi, a.zero_ids, a.n_zero, i))
                key = make_key(0.f, a.zero_ids[i]);
            found += __syncthreads_count(key != KEY_EMPTY);
            zrun = warp_merge_low32(zrun, warp_sort32(key, lane), lane);
        }
    }
    static_assert(SCAN_WARPS * multi_warp_keys<E>(MQT) >= 2 * SCAN_WARPS * 32, "scratch of cta_topk_of_lists32");
    for (unsigned b = fin; b < a.nq; b += n_fin) {
        cta_topk_of_lists32(a.cand + (size_t)b * total, gridDim.x, a.k, zrun, keys0, a.out_keys + (size_t)b * a.k, warp, lane);
        __syncthreads();
    }
    if (threadIdx.x == 0 && atomicAdd(a.ticket + 2, 1u) == n_fin - 1) {   // the last finisher to leave
        a.ticket[0] = 0; a.ticket[1] = 0; a.ticket[2] = 0;
        if (a.elapsed_out != nullptr) *a.elapsed_out = global_timer_ns() - *reinterpret_cast<volatile unsigned long long *>(a.t0_slot);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// 32 < k <= 256: the same pass with ONE candidate buffer per (CTA, query) — the CtaBuf of topk.cuh, as the
// single-query kernel uses for big k — instead of a sorted list per (warp, query). A (row, query) result that beats
// its query's threshold is appended by the lane that owns it with one shared-memory atomic (no warp-serialised
// insertion, no per-warp merges: 8 queries x k = 100 cost 4.0 ms per pass with the per-warp lists, of which 1.3 ms
// was selection); every 8 iterations the CTA checks whether a buffer is within `slack` of full and, if so, sorts it,
// keeps the best k and tightens the threshold. a.kpad = capacity of one buffer (ctabuf_cap(k)).
// smem: queries [MQT][dim4] float4 | buffers [MQT][cap] keys.
template <int V, int R, int MQ, int NG>
__global__ void __launch_bounds__(SCAN_THREADS, 2) scan_multi_cta_topk_kernel(const MultiArgs a)
{
    constexpr bool EXACT = true;
    constexpr int MQT = MQ * NG;   // queries per pass (NG groups of MQ share the rows a warp has loaded, see above)
    constexpr int NV = R * MQ;
    constexpr int SH = (NV == 32) ? 0 : (NV == 16 ? 1 : 2);
    static_assert(NV == 32 || NV == 16 || NV == 8, "R*MQ must be 8, 16 or 32");
    constexpr uint32_t SYNC_IT = NG == 1 ? 8 : 4;            // iterations between two sync points (NG = 2: smaller buffers)
    constexpr uint32_t SLACK = SYNC_IT * R * SCAN_WARPS;     // most keys one query can receive between two sync points
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned s_ticket;
    if (a.t0_slot != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *a.t0_slot = global_timer_ns();
    __shared__ unsigned cnt_s[MQT];
    __shared__ uint64_t thr_s[MQT];
    __shared__ float qflag[MQT];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t dim4 = a.dim4, cap = a.kpad, k = a.k;
    float4 *qs = reinterpret_cast<float4 *>(smem_raw);
    uint64_t *bufs = reinterpret_cast<uint64_t *>(smem_raw + (size_t)MQT * dim4 * sizeof(float4));

    for (int b = warp; b < MQT; b += SCAN_WARPS) {   // queries -> smem, unit-normalised exactly as scan.cuh does it
        float4 t[V];
        float ss = 0.f;
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const uint32_t c = lane + 32 * j;
            if ((uint32_t)b < a.nq && (EXACT || c < dim4)) t[j] = reinterpret_cast<const float4 *>(a.q)[(size_t)b * dim4 + c];
            else t[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            ss = fmaf(t[j].x, t[j].x, ss); ss = fmaf(t[j].y, t[j].y, ss);
            ss = fmaf(t[j].z, t[j].z, ss); ss = fmaf(t[j].w, t[j].w, ss);
        }
        ss = warp_sum_tree(ss);
        const bool qzero = !(ss > 0.f);
        const float qinv = qzero ? 0.f : 1.0f / sqrtf(ss);
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const uint32_t c = lane + 32 * j;
            if (EXACT || c < dim4) qs[(size_t)b * dim4 + c] = make_float4(t[j].x * qinv, t[j].y * qinv, t[j].z * qinv, t[j].w * qinv);
        }
        if (lane == 0) qflag[b] = qzero ? 1.f : 0.f;
    }
    if (threadIdx.x < MQT) { cnt_s[threadIdx.x] = 0; thr_s[threadIdx.x] = KEY_EMPTY; }
    __syncthreads();

    const int own = lane >> SH;
    const bool owner = (lane & ((1 << SH) - 1)) == 0;
    const int my_b = own % MQ, my_r = own / MQ;
    bool my_active[NG], my_qzero[NG];
    uint64_t thr[NG];
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        my_active[g] = owner && (uint32_t)(my_b + g * MQ) < a.nq;
        my_qzero[g] = qflag[my_b + g * MQ] != 0.f;
        thr[g] = my_active[g] ? KEY_EMPTY : 0ull;
    }

    // CTA-uniform: compact every buffer that could overflow before the next sync point, refresh the thresholds
    auto sync_point = [&]() {
        bool maybe = false;
#pragma unroll
        for (int b = 0; b < MQT; ++b) maybe |= *reinterpret_cast<volatile unsigned *>(&cnt_s[b]) + SLACK > cap;
        if (__syncthreads_or(maybe)) {   // nobody appends while the CTA is in here, so the counts are stable
            for (int b = 0; b < MQT; ++b)
                if (*reinterpret_cast<volatile unsigned *>(&cnt_s[b]) + SLACK > cap)
                    cta_buf_compact(bufs + (size_t)b * cap, &cnt_s[b], &thr_s[b], cap, k);
        }
#pragma unroll
        for (int g = 0; g < NG; ++g)
            if (my_active[g]) thr[g] = *reinterpret_cast<volatile uint64_t *>(&thr_s[my_b + g * MQ]);
    };

    const uint64_t n = a.n_rows;
    const uint64_t gw = (uint64_t)blockIdx.x * SCAN_WARPS + warp;
    const uint64_t n_warps = (uint64_t)gridDim.x * SCAN_WARPS;
    const uint64_t n_groups = (n + R - 1) / R, g_first = (uint64_t)blockIdx.x * SCAN_WARPS;
    const uint64_t n_iters = g_first < n_groups ? (n_groups - g_first + n_warps - 1) / n_warps : 0;   // same for every warp of the CTA
    auto scan_group = [&](uint64_t base) {
        float4 x[R][V];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const uint64_t row = base + r;
            const float4 *p = a.rows + row * dim4 + lane;
#pragma unroll
            for (int j = 0; j < V; ++j) {
                if (row < n && (EXACT || lane + 32 * j < dim4)) x[r][j] = ldg_stream<0>(p + 32 * j);
                else x[r][j] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            if (g > 0 && (uint32_t)(g * MQ) >= a.nq) break;   // CTA-uniform: no query in this group
            float v[32];
#pragma unroll
            for (int i = 0; i < NV; ++i) v[i] = 0.f;
#pragma unroll
            for (int j = 0; j < V; ++j) {
#pragma unroll
                for (int b = 0; b < MQ; ++b) {
                    float4 qq = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (EXACT || lane + 32 * j < dim4) qq = qs[(size_t)(g * MQ + b) * dim4 + lane + 32 * j];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        float acc = v[r * MQ + b];
                        acc = fmaf(x[r][j].x, qq.x, acc); acc = fmaf(x[r][j].y, qq.y, acc);
                        acc = fmaf(x[r][j].z, qq.z, acc); acc = fmaf(x[r][j].w, qq.w, acc);
                        v[r * MQ + b] = acc;
                    }
                }
            }
            butterfly_level<NV, 16>(v, lane);
            butterfly_level<(NV / 2 > 0 ? NV / 2 : 1), 8>(v, lane);
            butterfly_level<(NV / 4 > 0 ? NV / 4 : 1), 4>(v, lane);
            butterfly_level<(NV / 8 > 0 ? NV / 8 : 1), 2>(v, lane);
            butterfly_level<(NV / 16 > 0 ? NV / 16 : 1), 1>(v, lane);
            const float dist = my_qzero[g] ? 0.f : fmaf(-0.5f, v[0], 0.5f);
            const uint64_t row = base + my_r;
            if (my_active[g] && row < n && okey(dist) <= (uint32_t)(thr[g] >> 32)) {
                const uint32_t id = a.ids[row];
                const uint64_t kk = make_key(dist, id);
                if (kk < thr[g] && multi_row_allowed(a, row, id))
                    bufs[(size_t)(my_b + g * MQ) * cap + atomicAdd(&cnt_s[my_b + g * MQ], 1u)] = kk;
            }
        }
    };
    if (a.static_split) {
        for (uint64_t it = 0; it < n_iters; ++it) {
            scan_group((gw + it * n_warps) * R);
            if ((it % SYNC_IT) == SYNC_IT - 1) sync_point();
        }
    } else {
        // dynamic row split: the CTA synchronises at its sync points anyway, so the whole CTA grabs up to SYNC_IT (the
        // interval SLACK is sized for) "CTA iterations" (one group of R rows per warp) at a time from a global counter;
        // thread 0 issues the next grab before the chunk is processed (see scan.cuh, DYN && BIG)
        __shared__ uint32_t s_start;
        const uint32_t n_cit = (uint32_t)((n_groups + SCAN_WARPS - 1) / SCAN_WARPS), g2 = 2u * gridDim.x;
        unsigned *work = a.ticket + 1;
        uint32_t c_next = max(1u, min(SYNC_IT, n_cit / g2)), nxt = 0;
        if (threadIdx.x == 0) nxt = atomicAdd(work, c_next);
        for (;;) {
            if (threadIdx.x == 0) s_start = nxt;
            __syncthreads();
            const uint32_t c_start = s_start;
            if (c_start >= n_cit) break;
            const uint32_t c_end = min(c_start + c_next, n_cit);
            c_next = max(1u, min(SYNC_IT, (n_cit - c_start) / g2));
            if (threadIdx.x == 0) nxt = atomicAdd(work, c_next);
            for (uint32_t it = c_start; it < c_end; ++it) scan_group(((uint64_t)it * SCAN_WARPS + warp) * R);
            sync_point();   // barrier inside: s_start is free to be rewritten
        }
    }

    // ---- per query: CTA top-k -> cand[b][cta][k] ----
    for (int b = 0; b < (int)a.nq; ++b) {
        cta_buf_compact(bufs + (size_t)b * cap, &cnt_s[b], &thr_s[b], cap, k);
        uint64_t *dst = a.cand + ((size_t)b * gridDim.x + blockIdx.x) * k;
        for (uint32_t j = threadIdx.x; j < k; j += blockDim.x) dst[j] = bufs[(size_t)b * cap + j];
    }
    __threadfence();
    __syncthreads();
    // the LAST nq CTAs to arrive finish one query each (see scan_multi_topk_kernel: a finisher waits for every CTA's lists;
    // at most nq <= 16 CTAs ever wait and all others exit, so the spin cannot starve them)
    if (threadIdx.x == 0) s_ticket = atomicAdd(a.ticket, 1u);
    __syncthreads();
    const unsigned n_fin = min(a.nq, gridDim.x), first_fin = gridDim.x - n_fin;
    if (s_ticket < first_fin) return;
    const unsigned fin = s_ticket - first_fin;
    if (threadIdx.x == 0) {
        const volatile unsigned *tk = a.ticket;
        while (*tk < gridDim.x) {}
    }
    __syncthreads();
    __threadfence();

    // ---- finisher: merge one query across CTAs, column-wise over the sorted per-CTA lists (see scan.cuh) ----
    const uint32_t L = gridDim.x;
    const uint64_t total = (uint64_t)L * k;
    for (unsigned b = fin; b < a.nq; b += n_fin) {
        uint64_t *buf = bufs + (size_t)b * cap;
        const volatile uint64_t *cand = a.cand + (size_t)b * total;
        if (threadIdx.x == 0) { cnt_s[b] = 0; thr_s[b] = KEY_EMPTY; }
        __syncthreads();
        volatile uint64_t *thr_b = &thr_s[b];
        for (uint32_t j = 0; j < k; ++j) {
            bool passed = false;
            for (uint32_t c = threadIdx.x; c < L; c += blockDim.x) {
                const uint64_t key = cand[(size_t)c * k + j];
                if (key < *thr_b) { buf[atomicAdd(&cnt_s[b], 1u)] = key; passed = true; }
            }
            if (!__syncthreads_or(passed)) break;
            const unsigned cnt = *reinterpret_cast<volatile unsigned *>(&cnt_s[b]);
            const bool need = cnt + L > cap || (*thr_b == KEY_EMPTY && cnt >= k);
            if (__syncthreads_or(need)) cta_buf_compact(buf, &cnt_s[b], thr_b, cap, k);
        }
        uint32_t found = 0;   // zero-norm rows: distance 0.0, ascending id; the first k allowed ones suffice
        for (uint32_t o = 0; o < a.n_zero && found < k; o += blockDim.x) {
            uint64_t key = KEY_EMPTY;
            if (o + threadIdx.x < a.n_zero &&
                zero_row_allowed(a.bitmap, a.n_bits, a.tags, a.lang_mask, a.file_lo, a.file_hi, a.zero_ids, a.n_zero, o + threadIdx.x))
                key = make_key(0.f, a.zero_ids[o + threadIdx.x]);
            found += __syncthreads_count(key != KEY_EMPTY);
            if (key < *thr_b) buf[atomicAdd(&cnt_s[b], 1u)] = key;
            if (__syncthreads_or(*reinterpret_cast<volatile unsigned *>(&cnt_s[b]) + blockDim.x > cap))
                cta_buf_compact(buf, &cnt_s[b], thr_b, cap, k);
        }
        cta_buf_compact(buf, &cnt_s[b], thr_b, cap, k);
        for (uint32_t j = threadIdx.x; j < k; j += blockDim.x) a.out_keys[(size_t)b * k + j] = buf[j];
        __syncthreads();
    }
    if (threadIdx.x == 0 && atomicAdd(a.ticket + 2, 1u) == n_fin - 1) {   // the last finisher to leave
        a.ticket[0] = 0; a.ticket[1] = 0; a.ticket[2] = 0;
        if (a.elapsed_out != nullptr) *a.elapsed_out = global_timer_ns() - *reinterpret_cast<volatile unsigned long long *>(a.t0_slot);
    }
}


}  // namespace csgpu
