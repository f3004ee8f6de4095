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
// round-to-nearest bf16 images: |x~ - x| <= 2^-9 |x| per element, hence ||x~ - x|| <= 2^-9 ||x||.
//     |q~.r~ - q.r| <= |q~.(r~ - r)| + |(q~ - q).r| <= (1 + 2^-9) 2^-9 + 2^-9  <  2^-8 (1 + 2^-9)        (Cauchy-Schwarz)
// The tensor core multiplies bf16 pairs exactly and accumulates in fp32 (at worst truncating): <= dim * 2^-23 * sum|q~_i r~_i|
// <= dim * 2^-23; the fp32 FMA chain of the scan contributes <= dim * 2^-24. With dim <= 512 (bf16 kernel limit):
//     |cos_tc - cos_f32| < 3.914e-3 + 6.2e-5 + 3.1e-5 < 4.02e-3      =>      |d_tc - d_f32| < 2.01e-3
// TC_MARGIN = 2.1e-3 (distance units) leaves slack for the two final roundings. Exactness of the filter:
//   * a row of the final top-k has d_f32 <= T_final <= T (the exact k-th best so far, never tighter than final), so
//     d_tc <= d_f32 + MARGIN <= T + MARGIN: it passes the epilogue (which tests d_tc <= thr with thr = T + MARGIN);
//   * here a candidate is skipped unread only if d_tc - MARGIN > T_warp >= T_final, i.e. d_f32 > T_final.
// tests/test_gpu_prefilter.py checks bit-equality with the single-query kernel, and adversarial inputs (many near-ties
// inside the margin) exercise the overflow path.
#pragma once
#include "scan.cuh"

namespace csgpu {

constexpr float TC_MARGIN = 2.1e-3f;

struct RescoreArgs {
    const float4 *rows;        // [n_rows, dim4] fp32 unit rows
    const uint32_t *ids;       // [n_rows]
    uint32_t dim4;
    const float *q_raw;        // [nq][dim4*4] raw queries (normalised in the prologue, exactly like the scan kernel)
    const uint8_t *flags;      // [nq] zero-norm query flags (such queries are answered by the scan kernel instead)
    uint64_t *cand;            // [nq][cap]; entries [0, n_done[q]) = exact keys with chunk ids (survivors of earlier
                               // selects), entries beyond = (okey(d_tc) << 32 | ROW index) from the GEMM epilogue
    unsigned *count;           // [nq]
    const unsigned *n_done;    // [nq]
    float *thr;                // [nq] out: exact k-th best distance + TC_MARGIN (+inf until k rows are known)
    uint32_t cap, k, kpad, n_active;
    const uint32_t *zero_ids;  // final pass only
    uint32_t n_zero;
    uint64_t *final_out;       // [nq][k] or nullptr
    unsigned long long *n_rescored;   // optional statistics: rows actually read
};

// One CTA per query. V/EXACT as in scan_topk_kernel; R candidate rows in flight per warp.
template <int V, bool EXACT, bool BIG>
__global__ void __launch_bounds__(SCAN_THREADS, 2) rescore_select_kernel(const RescoreArgs a)
{
    extern __shared__ __align__(16) uint64_t smem[];
    constexpr int R = 4;
    const uint32_t q = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (q >= a.n_active || a.flags[q]) {
        if (threadIdx.x == 0) { a.count[q] = 0; a.thr[q] = -1.f; }
        return;
    }
    using Sel = typename SelOf<BIG>::type;
    Sel sel;
    if constexpr (BIG) sel.init(smem + (size_t)warp * a.kpad, smem + (size_t)(SCAN_WARPS + warp) * a.kpad, a.k, a.kpad, lane);
    else sel.init(a.k);

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

    uint64_t *mine = a.cand + (size_t)q * a.cap;
    const uint32_t n = min(a.count[q], a.cap);
    const uint32_t done = min(a.n_done[q], n);
    unsigned read_rows = 0;
    for (uint32_t b = warp * 32; b < n; b += SCAN_WARPS * 32) {
        uint64_t key = (b + lane < n) ? mine[b + lane] : KEY_EMPTY;
        bool todo = false;
        if (b + lane >= done && key != KEY_EMPTY) {
            // candidate from the tensor-core filter: can it still make the warp's top-k?  d_f32 >= d_tc - MARGIN
            const float d_tc = __uint_as_float(bits_from_okey((uint32_t)(key >> 32)));
            todo = okey(d_tc - TC_MARGIN) <= (uint32_t)(sel.thr >> 32);
            if (!todo) key = KEY_EMPTY;
        }
        unsigned m = __ballot_sync(FULL, todo);
        const uint32_t my_row = (uint32_t)key;
        while (m) {
            int src[R];
            uint32_t row[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                src[r] = m ? (__ffs(m) - 1) : -1;
                m &= m - 1;
                row[r] = __shfl_sync(FULL, my_row, src[r] < 0 ? 0 : src[r]);
            }
            float4 x[R][V];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float4 *p = a.rows + (size_t)row[r] * a.dim4 + lane;
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    if (src[r] >= 0 && (EXACT || lane + 32 * j < a.dim4)) x[r][j] = __ldg(p + 32 * j);
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
                if (src[r] >= 0) {
                    ++read_rows;
                    if (lane == src[r]) key = make_key(dist, __ldg(a.ids + row[r]));
                }
            }
        }
        offer_lane_keys(sel, key, lane);
    }
    if (a.final_out != nullptr && warp == 0 && a.n_zero) {   // zero-norm rows: distance 0.0 (arroy pn*qn == 0)
        uint32_t found = 0;
        for (uint32_t b = 0; b < a.n_zero && found < a.k; b += 32) {
            const uint64_t key = (b + lane < a.n_zero) ? make_key(0.f, a.zero_ids[b + lane]) : KEY_EMPTY;
            found += __popc(__ballot_sync(FULL, key != KEY_EMPTY));
            offer_lane_keys(sel, key, lane);
        }
    }
    if (a.n_rescored != nullptr && lane == 0 && read_rows) atomicAdd(a.n_rescored, (unsigned long long)read_rows);
    __syncthreads();   // every warp has finished reading cand[q] before it is overwritten
    cta_reduce<BIG>(sel, smem, a.k, a.kpad, mine, warp, lane);
    if (a.final_out != nullptr)
        for (uint32_t j = threadIdx.x; j < a.k; j += blockDim.x) a.final_out[(size_t)q * a.k + j] = mine[j];
    if (threadIdx.x == 0) {
        uint32_t mcount = 0;
        while (mcount < a.k && mine[mcount] != KEY_EMPTY) ++mcount;
        a.count[q] = mcount;
        float t = __int_as_float(0x7f800000);  // +inf: everything passes until k rows are known
        if (mcount >= a.k) t = __uint_as_float(bits_from_okey((uint32_t)(mine[a.k - 1] >> 32))) + TC_MARGIN;
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
