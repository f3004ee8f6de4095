// rescore.cuh — exact fp32 results at tensor-core speed: the opt-in "tensor prefilter" of an fp32 index
// (csgpu_set_tensor_prefilter). BASELINE config C3 on the DEFAULT fp32 index without giving up the exact ranking.
//
// The fp32 index keeps a bf16 SHADOW of its unit rows (+50 % HBM). A batch is contracted on the tensor cores against
// the shadow (gemm_topk_kernel, tcgen05 / TMEM) — but only as a FILTER: a (query, row) pair survives the epilogue if
// its bf16 distance is within a proven error bound of the query's current exact threshold. Survivors are then
// RESCORED here from the fp32 rows with exactly the arithmetic of the single-query scan kernel (scan.cuh: the same
// query normalisation, the same per-lane FMA chain, the same xor-shuffle tree, distance = fma(-0.5, cos, 0.5)), and
// the top-k selection runs on those exact keys. Results are therefore bit-identical to csgpu_search — ids AND
// distances — while > 99.9 % of the 2*N*dim*B flops run on the tensor pipe.
//
// Error bound (TC_MARGIN). Rows r and the query q are unit vectors in fp32 (|.| <= 1 + 2^-22); r~, q~ are their
// round-to-nearest bf16 images. bf16 carries 8 significant bits (1 implicit + 7 stored), so its unit roundoff is 2^-8:
// |x~ - x| <= 2^-8 |x| per element, hence ||x~ - x|| <= 2^-8 ||x||.
//     |q~.r~ - q.r| <= |q~.(r~ - r)| + |(q~ - q).r| <= (1 + 2^-8) 2^-8 + 2^-8  <  2^-7 (1 + 2^-9)        (Cauchy-Schwarz)
// The tensor core multiplies bf16 pairs exactly and accumulates in fp32 (at worst truncating): <= dim * 2^-23 * sum|q~_i r~_i|
// <= dim * 2^-23; the fp32 FMA chain of the scan contributes <= dim * 2^-24. With dim <= 512 (bf16 kernel limit):
//     |cos_tc - cos_f32| < 7.828e-3 + 6.2e-5 + 3.1e-5 < 7.93e-3      =>      |d_tc - d_f32| < 3.97e-3
// TC_MARGIN = 4.0e-3 (distance units). (Round 1 shipped 2.1e-3, derived with bf16's roundoff taken as 2^-9 — one bit too
// optimistic; no input ever came near it — the largest error the rescoring has seen is ~2.5e-4, csgpu_stats_t.filter_max_err
// — but a bound has to be a bound. The wider margin lets ~1.4x more candidates through; stage A/B below read few of them.)
// Exactness of the filter:
//   * a row of the final top-k has d_f32 <= T_final <= T (the exact k-th best so far, never tighter than final), so
//     d_tc <= d_f32 + MARGIN <= T + MARGIN: it passes the epilogue (which tests d_tc <= thr with thr = T + MARGIN);
//   * the select kernel drops a candidate unread only if d_tc - MARGIN > T' for some exact T' >= T_final, i.e. d_f32 > T_final.
// tests/test_gpu_prefilter.py checks bit-equality with the single-query kernel, and adversarial inputs (many near-ties
// inside the margin) exercise the overflow path.
#pragma once
#include "scan.cuh"

namespace csgpu {

constexpr float TC_MARGIN = 4.0e-3f;

// ---------------------------------------------------------------------------------------------------------------
// select_sorted_kernel — the per-query reduction between the row phases of every GEMM-shaped batch kernel
// (gemm_topk.cu): candidates -> exact top-k (ascending (distance, id)), written to the front of the query's block,
// new threshold published. One CTA per query; everything happens in shared memory around CTA-wide bitonic sorts.
//
// Inputs per query (see GemmTopkArgs): part 1 = cand[q*stride + 0 .. count[q])  — entries below n_done[q] are
// survivors of the previous select (exact keys carrying chunk ids), entries above are fresh candidates carrying ROW
// indices (SIMT kernel, global-atomic layout); part 2 = the tensor-core kernel's segments (all fresh candidates).
//
// RESCORE = false (bf16 index, fp32 SIMT kernel): candidate distances are final; swap ids in, one sort, keep k.
// RESCORE = true  (tensor prefilter): candidate distances are bf16 estimates d_tc. Sort the candidates by d_tc, then
//   stage A  rescore the best-looking a = pow2 >= max(32, k) of them from the fp32 rows (exact keys, in place, then sorted),
//            T_A = k-th smallest exact key of survivors U stage A  (a valid upper bound of the final k-th best);
//   stage B  keep rescoring down the sorted list while d_tc - MARGIN <= T_A can still hold (a prefix: the list is
//            sorted), drop everything behind it unread;
//   final    sort survivors U rescored, keep k, publish thr = exact k-th distance + MARGIN.
//   About 1.5 k rows are read per query and phase instead of every candidate the margin let through.
struct SelectArgs {
    uint64_t *cand;            // [nq][stride]
    unsigned *count;           // [nq] in: entries of part 1; out: survivors
    const unsigned *n_done;    // [nq] part 1: entries below this are survivors (already exact, with ids)
    const unsigned *seg_count; // [nq][n_seg] or nullptr
    uint32_t stride, surv, seg_len, n_seg;
    unsigned *overflow;        // set to 1 if a query holds more candidates than the sort buffer (host splits the range)
    float *thr;                // [nq] out
    const uint32_t *ids;       // [n_rows]
    const uint8_t *flags;      // [nq] zero-norm query flags (such queries are answered by the scan kernel instead)
    uint32_t k, n_active;
    const uint32_t *zero_ids;  // final pass only
    uint32_t n_zero;
    uint64_t *final_out;       // [nq][k] or nullptr
    // RESCORE only
    const float4 *rows;        // [n_rows, dim4] fp32 unit rows
    uint32_t dim4;
    const float *q_raw;        // [nq][dim4*4] raw queries (normalised in the prologue, exactly like the scan kernel)
    unsigned long long *n_rescored;   // statistics: rows actually read
    float margin;              // proven bound of |d_filter - d_f32|: TC_MARGIN (bf16 shadow) or TF_MARGIN (tf32 off the fp32 rows, gemm_tf32.cuh)
    unsigned *max_err;         // optional statistics: float bits of the largest |d_filter - d_f32| seen (atomicMax on non-negative floats)
};

constexpr uint32_t SEL_BUF = 8192;                       // sort buffer (keys) = the largest bitonic sort a query needs
constexpr uint32_t SEL_SORT_CAP = SEL_BUF - 2 * 1024;    // fresh candidates per query; the rest holds survivors + zero-norm ids

// Exact distances of candidates C[lo, hi) (keys carrying row indices) from the fp32 rows, in place; a candidate whose
// lower bound d_tc - MARGIN already exceeds `bound32` (okey of an exact upper bound of the final k-th distance) is
// dropped unread. 8 warps x R rows in flight. Same arithmetic as scan_topk_kernel, operation for operation.
template <int V, bool EXACT>
__device__ __forceinline__ unsigned rescore_range(const SelectArgs &a, const float4 (&qv)[V], uint64_t *C, uint32_t lo, uint32_t hi,
                                                  uint32_t bound32, int warp, int lane)
{
    constexpr int R = V <= 4 ? 4 : 2;   // rows in flight per warp (registers: R x V float4)
    unsigned read_rows = 0;
    for (uint32_t i0 = lo + (uint32_t)warp * R; i0 < hi; i0 += SCAN_WARPS * R) {
        uint32_t row[R];
        bool live[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const uint32_t i = i0 + r;
            const uint64_t key = i < hi ? C[i] : KEY_EMPTY;
            live[r] = false;
            if (key != KEY_EMPTY) {
                const float d_tc = __uint_as_float(bits_from_okey((uint32_t)(key >> 32)));
                live[r] = okey(d_tc - a.margin) <= bound32;
            }
            row[r] = (uint32_t)key;
        }
        __syncwarp();      // every lane has read C[i0 .. i0+R) before lane 0 overwrites those slots below
        uint32_t idv[R];   // chunk ids ride along with the row loads (no dependent load after the reduction)
#pragma unroll
        for (int r = 0; r < R; ++r) idv[r] = live[r] ? __ldg(a.ids + row[r]) : 0u;
        float4 x[R][V];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float4 *p = a.rows + (size_t)row[r] * a.dim4 + lane;
#pragma unroll
            for (int j = 0; j < V; ++j) {
                if (live[r] && (EXACT || lane + 32 * j < a.dim4)) x[r][j] = __ldg(p + 32 * j);
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
            const float dist = fmaf(-0.5f, acc, 0.5f);
            if (i0 + r < hi && lane == 0) {
                if (a.max_err != nullptr && live[r]) {
                    const float d_tc = __uint_as_float(bits_from_okey((uint32_t)(C[i0 + r] >> 32)));
                    atomicMax(a.max_err, __float_as_uint(fabsf(d_tc - dist)));
                }
                C[i0 + r] = live[r] ? make_key(dist, idv[r]) : KEY_EMPTY;
            }
            read_rows += live[r] ? 1u : 0u;
        }
    }
    return read_rows;
}

// dynamic smem: C[SEL_BUF] | S[1024]   (u64 each; S only when RESCORE) — 72 KB, three CTAs per SM
template <int V, bool EXACT, bool RESCORE>
__global__ void __launch_bounds__(SCAN_THREADS, (V <= 4 ? 3 : 2)) select_sorted_kernel(const SelectArgs a)
{
    extern __shared__ __align__(16) uint64_t smem[];
    __shared__ unsigned s_n, s_pref[2 * 148 + 8], s_end, s_bound;
    uint64_t *C = smem, *S = smem + SEL_BUF;
    const uint32_t q = blockIdx.x, k = a.k;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (q >= a.n_active || a.flags[q]) {   // padding rows and zero-norm queries stay inactive
        if (threadIdx.x == 0) { a.count[q] = 0; a.thr[q] = -1.f; }
        return;
    }
    uint64_t *mine = a.cand + (size_t)q * a.stride;
    const uint32_t n1 = min(a.count[q], a.n_seg ? a.surv : a.stride);
    const uint32_t done = a.n_seg ? n1 : min(a.n_done[q], n1);

    // ---- gather: survivors -> S (RESCORE) or C; fresh candidates -> C ----
    // (the overflow checks that used to be a launch of their own — check_counts_kernel — ride along here: a segment or the
    //  SIMT buffer holding more than it can, or more fresh candidates than the sort buffer takes, raises the sticky flag)
    for (uint32_t s = threadIdx.x; s < a.n_seg; s += blockDim.x) {
        const unsigned c = a.seg_count[(size_t)q * a.n_seg + s];
        if (c > a.seg_len) atomicExch(a.overflow, 1u);
        s_pref[s] = min(c, a.seg_len);
    }
    if (threadIdx.x == 0 && !a.n_seg && a.count[q] > a.stride) atomicExch(a.overflow, 1u);
    __syncthreads();
    if (threadIdx.x == 0) {   // exclusive prefix of the segment counts (n_seg <= 2 * 148)
        unsigned run = n1 - done;
        for (uint32_t s = 0; s < a.n_seg; ++s) {
            const unsigned c = s_pref[s];
            s_pref[s] = run;
            run += c;
        }
        s_pref[a.n_seg] = run;
        if (run > SEL_SORT_CAP) { atomicExch(a.overflow, 1u); run = SEL_SORT_CAP; }
        s_n = run;
    }
    __syncthreads();
    const uint32_t n_c = s_n;                       // fresh candidates
    const uint32_t n_s = done;                      // survivors (<= k <= 1024), ascending
    for (uint32_t t = threadIdx.x; t < n1 - done; t += blockDim.x) {
        if (t < n_c) {
            uint64_t key = mine[done + t];
            if (!RESCORE) key = (key & 0xFFFFFFFF00000000ull) | a.ids[(uint32_t)key];
            C[t] = key;
        }
    }
    if (a.n_seg) {   // flat over all fresh candidates (every load in flight at once): entry t lives in the segment s with
                     // s_pref[s] <= t < s_pref[s + 1]
        for (uint32_t t = (n1 - done) + threadIdx.x; t < n_c; t += blockDim.x) {
            uint32_t lo = 0, hi = a.n_seg;   // last s with s_pref[s] <= t
            while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (s_pref[mid] <= t) lo = mid; else hi = mid; }
            uint64_t key = mine[a.surv + (size_t)lo * a.seg_len + (t - s_pref[lo])];
            if (!RESCORE) key = (key & 0xFFFFFFFF00000000ull) | a.ids[(uint32_t)key];
            C[t] = key;
        }
    }
    uint32_t n_all;   // entries of C that take part in the final sort
    if constexpr (!RESCORE) {
        for (uint32_t t = threadIdx.x; t < n_s; t += blockDim.x) C[n_c + t] = mine[t];
        n_all = n_c + n_s;
    } else {
        for (uint32_t t = threadIdx.x; t < 1024; t += blockDim.x) S[t] = t < n_s ? mine[t] : KEY_EMPTY;
        // ---- query -> registers, scaled to unit length: the scan kernel's prologue, operation for operation ----
        float4 qv[V];
        float ss = 0.f;
        const float4 *qp = reinterpret_cast<const float4 *>(a.q_raw) + (size_t)q * a.dim4;
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const uint32_t c = lane + 32 * j;
            if (EXACT || c < a.dim4) qv[j] = qp[c];
            else qv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            ss = fmaf(qv[j].x, qv[j].x, ss); ss = fmaf(qv[j].y, qv[j].y, ss);
            ss = fmaf(qv[j].z, qv[j].z, ss); ss = fmaf(qv[j].w, qv[j].w, ss);
        }
        ss = warp_sum_tree(ss);
        const float qinv = 1.0f / sqrtf(ss);   // ss > 0: zero-norm queries were filtered out above
#pragma unroll
        for (int j = 0; j < V; ++j) { qv[j].x *= qinv; qv[j].y *= qinv; qv[j].z *= qinv; qv[j].w *= qinv; }

        // candidates ascending by d_tc
        const uint32_t npad = pow2_at_least(n_c, 32);
        for (uint32_t t = n_c + threadIdx.x; t < npad; t += blockDim.x) C[t] = KEY_EMPTY;
        cta_sort(C, npad);
        // stage A: a power of two (so the rescored keys can be sorted in place without touching stage B's part of C)
        const uint32_t a_pow = pow2_at_least(k, 32);
        const uint32_t a_end = min(n_c, a_pow);
        const uint32_t a_pad = n_c < a_pow ? npad : a_pow;      // n_c < a_pow: everything is stage A, C[n_c, npad) is EMPTY
        const uint32_t bound0 = n_s >= k ? (uint32_t)(S[k - 1] >> 32) : 0xFFFFFFFFu;
        unsigned read_rows = rescore_range<V, EXACT>(a, qv, C, 0, a_end, bound0, warp, lane);
        cta_sort(C, a_pad);   // exact keys ascending, dropped ones (EMPTY) last
        // T_A = k-th smallest exact key of survivors U stage A: both lists are sorted, so it is the k-th element of their
        // merge — found by one thread with a binary search over how many of the k come from S
        if (threadIdx.x == 0) {
            uint32_t n_a = a_end;
            while (n_a > 0 && C[n_a - 1] == KEY_EMPTY) --n_a;
            uint32_t bound = 0xFFFFFFFFu;   // fewer than k exact keys known: no pruning
            if (n_s + n_a >= k) {
                uint32_t lo = k > n_a ? k - n_a : 0, hi = min(k, n_s);   // i = number taken from S
                while (lo < hi) {
                    const uint32_t i = (lo + hi) >> 1;                    // take i from S, k - i from A
                    if (S[i] < C[k - i - 1]) lo = i + 1; else hi = i;     // S[i] would also belong to the k smallest
                }
                const uint64_t from_s = lo > 0 ? S[lo - 1] : 0ull, from_a = k - lo > 0 ? C[k - lo - 1] : 0ull;
                bound = (uint32_t)(umax64(from_s, from_a) >> 32);
            }
            s_bound = bound;
            s_end = n_c;
        }
        __syncthreads();
        const uint32_t bound_a = s_bound;
        // stage B: the sorted prefix that can still beat T_A
        for (uint32_t t = a_end + threadIdx.x; t < n_c; t += blockDim.x) {
            const float d_tc = __uint_as_float(bits_from_okey((uint32_t)(C[t] >> 32)));
            if (okey(d_tc - a.margin) > bound_a) atomicMin(&s_end, t);
        }
        __syncthreads();
        const uint32_t b_end = s_end;
        read_rows += rescore_range<V, EXACT>(a, qv, C, a_end, b_end, bound_a, warp, lane);
        if (a.n_rescored != nullptr && lane == 0 && read_rows) atomicAdd(a.n_rescored, (unsigned long long)read_rows);
        __syncthreads();
        for (uint32_t t = threadIdx.x; t < n_s; t += blockDim.x) C[b_end + t] = S[t];
        n_all = b_end + n_s;
    }
    __syncthreads();
    if (a.final_out != nullptr && a.n_zero) {   // zero-norm rows: distance 0.0 (arroy pn*qn == 0); the first k ids suffice
        const uint32_t nz = min(a.n_zero, k);
        for (uint32_t t = threadIdx.x; t < nz; t += blockDim.x) C[n_all + t] = make_key(0.f, a.zero_ids[t]);
        n_all += nz;
    }
    const uint32_t fpad = pow2_at_least(n_all, 32);
    for (uint32_t t = n_all + threadIdx.x; t < fpad; t += blockDim.x) C[t] = KEY_EMPTY;
    cta_sort(C, fpad);
    // ---- survivors to the front of the block, threshold ----
    for (uint32_t j = threadIdx.x; j < k; j += blockDim.x) {
        const uint64_t key = j < fpad ? C[j] : KEY_EMPTY;
        mine[j] = key;
        if (a.final_out != nullptr) a.final_out[(size_t)q * k + j] = key;
    }
    if (threadIdx.x == 0) {
        uint32_t m = min(n_all, k);
        while (m > 0 && C[m - 1] == KEY_EMPTY) --m;   // dropped candidates sort to the end
        a.count[q] = m;
        float t = __int_as_float(0x7f800000);  // +inf: everything passes until k rows are known
        if (m >= k) t = __uint_as_float(bits_from_okey((uint32_t)(C[k - 1] >> 32))) + (RESCORE ? a.margin : 0.f);
        a.thr[q] = t;
    }
}

// fp32 unit rows -> bf16 shadow (round to nearest even), one thread per 4 elements
static __global__ void shadow_from_rows_kernel(const float4 *__restrict__ rows, uint2 *__restrict__ out, uint64_t n4)
{
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n4; t += (uint64_t)gridDim.x * blockDim.x) {
        const float4 v = rows[t];
        const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
        out[t] = make_uint2(*reinterpret_cast<const uint32_t *>(&lo), *reinterpret_cast<const uint32_t *>(&hi));
    }
}

}  // namespace csgpu
