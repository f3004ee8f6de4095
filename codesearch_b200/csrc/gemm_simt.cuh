// gemm_simt.cuh — batched search on the default fp32 index: register-tiled fp32 SIMT Q x C^T with the
// top-k threshold filter fused into the epilogue (BASELINE config C3, "fp32 SIMT"; SURVEY.md §7 kernel 4).
// Serves csgpu_search_batch for large batches; small batches (the <= 9 query variants of
// /root/reference/src/search/mod.rs:508-511) stay on the HBM-bound multi-query scan (scan_multi.cuh).
//
// Bound: the FP32 FMA pipe, not HBM and not the tensor cores (the index is fp32 and results must equal
// the exact fp32 ranking, so the contraction is not reshaped for tcgen05). Algorithmic work per
// (query, row) = dim FMAs = 2*dim flop; a batch of B queries over N rows = 2*N*dim*B flop
// (7.864 TFLOP at 10M x 384 x 1024). Measured ceiling of this loop shape on B200 (tools/ubench_ffma.cu,
// operands from shared memory, 8x8 per thread): 51-54 TFLOP/s.
//
// Shapes. One CTA = (16 TM) queries x (16 TN) corpus rows per work item, K in 32-float (128 B) chunks; 256 compute threads
//   in a 16 x 16 grid, TM x TN outputs each. Round 2 retiled from 8 x 8 (128 x 128) to 8 x 12 (128 queries x 192 rows): the
//   round-2 microbenchmark (tools/ubench_ffma2.cu, profiles/r02_ubench_ffma2.txt) shows that the FMA pipe ALONE tops out
//   at 60.9 TFLOP/s on this part (registers only, no memory: 82 % of 148 x 128 x 2 x 1.965 GHz), that the 8 x 8 loop with
//   operands from shared memory reaches 53.2 whatever the lane layout or the number of warps, and that what moves it is
//   the LDS.128 : FFMA ratio — 8 x 12 (0.052 instead of 0.0625) reaches 57.2-57.9, 8 x 16 spills. A second instantiation,
//   4 x 16 (64 queries x 256 rows), serves batches of <= 64 queries, which used to pay for a half-empty 128-query tile.
//   smem ring (STAGES deep): per stage [16 TM q][32] + [16 TN rows][32] fp32 = 40 KB, written by TMA in the
//   SWIZZLE_128B layout (16-byte chunk index XOR (row & 7)), so the float4 reads below are conflict-free
//   without padding.
//   Outputs of a thread: queries i*16 + ty, rows j*16 + tx (interleaved, so that the 8 lanes of a quarter-warp hit 8
//   different swizzle phases); a warp covers 4 ty x 8 tx -> an A fragment read touches 4 distinct 16 B words
//   (broadcast), a B fragment read 8.
//   1 producer warp: one elected lane issues the TMA loads (cp.async.bulk.tensor, SASS UTMALDG).
// Every (query, row) dot product is one fixed FMA chain over k = 0..dim-1 in one thread, so a score does
// not depend on where the row sits: duplicate rows tie bit-exactly and the id tie-break is well defined.
//
// Work items (row tile t, query block qb) are dealt round-robin with qb fastest, so the ~148/QB row tiles
// in flight are read from HBM once and served to the other query blocks from L2.
//
// Epilogue = the progressive-threshold filter of gemm_topk.cuh: a score passes if its distance is <= the
// query's current threshold (k-th best of the rows scanned in earlier phases); passing keys are appended
// to the query's candidate buffer (one atomicAdd per (thread, query) per item). Scores never touch HBM.
#pragma once
#include <cuda.h>

#include "gemm_topk.cuh"

namespace csgpu {

constexpr int GS_BK = 32;         // floats per K chunk (128 B = one swizzle row)
constexpr int GS_COMPUTE_WARPS = 8;
constexpr int GS_THREADS = (GS_COMPUTE_WARPS + 4) * 32;   // 384: two compute warpgroups + one producer warpgroup
// the two tile shapes the library instantiates (gemm_topk.cu picks by batch size)
constexpr int GS_TM = 8, GS_TN = 12;            // large batches: 128 queries x 192 rows
constexpr int GS_TM_SMALL = 4, GS_TN_SMALL = 16;   // <= 64 queries: 64 queries x 256 rows
constexpr int GS_BM = GS_TM * 16, GS_BN = GS_TN * 16;
constexpr int GS_BM_SMALL = GS_TM_SMALL * 16, GS_BN_SMALL = GS_TN_SMALL * 16;
__host__ __device__ constexpr uint32_t gs_stage_bytes(int tm, int tn) { return (uint32_t)(tm + tn) * 16u * GS_BK * 4u; }

// Same argument block as the bf16 kernel (GemmTopkArgs): n_kchunks = ceil(dim_pad / 32); n_qblocks counts BM-query blocks.
//
// Registers. The 8 x 12 tile needs ~200 registers per compute thread. A ninth (producer) warp next to eight compute warps
// puts three warps on one SM sub-partition (16 K registers each) and caps every thread at 170 (round 1's 288-thread 8 x 8
// kernel sat at 160); issuing the TMA loads from a compute thread instead was built and measured 8 % slower (the small
// tile: 10.8 -> 11.65 ms). So the CTA is three warpgroups and the register file is re-split with setmaxnreg, the
// warp-specialised pattern of the tensor-core GEMMs: the producer warpgroup drops to 40 registers per thread, the two
// compute warpgroups rise to 232 (128 x 40 + 256 x 232 = 64512 <= 65536).
template <int TM, int TN, int STAGES>
__global__ void __launch_bounds__(GS_THREADS, 1)
gemm_simt_topk_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_c, const GemmTopkArgs a)
{
    constexpr int BM = TM * 16, BN = TN * 16;
    constexpr uint32_t A_BYTES = BM * GS_BK * 4, STAGE_BYTES = gs_stage_bytes(TM, TN);
    extern __shared__ __align__(1024) unsigned char gs_smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // 1024-B alignment for the 128-byte swizzle atoms; offset arithmetic keeps the pointer in the shared window (LDS, not LD)
    unsigned char *ring = gs_smem_raw + ((1024u - (smem_u32(gs_smem_raw) & 1023u)) & 1023u);

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], GS_COMPUTE_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const uint64_t n_items = (a.tile_end - a.tile_begin) * a.n_qblocks;

    if (warp >= GS_COMPUTE_WARPS) {
        // ================= producer warpgroup: give the registers away, one lane drives the TMA engine =================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == GS_COMPUTE_WARPS && lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_c) : "memory");
            uint32_t it = 0;
            for (uint64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
                const uint32_t qb = (uint32_t)(item % a.n_qblocks);
                const uint64_t t = a.tile_begin + item / a.n_qblocks;
                for (uint32_t kc = 0; kc < a.n_kchunks; ++kc, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait_backoff(&empty_bar[s], ph ^ 1, 256);   // a stage lasts ~3 us: do not spin
                    mbar_expect_tx(&full_bar[s], STAGE_BYTES);
                    unsigned char *dst = ring + (size_t)s * STAGE_BYTES;
                    tma_load_2d(dst, &map_q, &full_bar[s], (int32_t)(kc * GS_BK), (int32_t)(qb * BM));
                    tma_load_2d(dst + A_BYTES, &map_c, &full_bar[s], (int32_t)(kc * GS_BK), (int32_t)(t * BN));
                }
            }
        }
        return;
    }
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");

    // ================= compute warps =================
    const int tx = (warp & 1) * 8 + (lane & 7);     // row slot   0..15
    const int ty = (warp >> 1) * 4 + (lane >> 3);   // query slot 0..15
    const uint32_t sa = (uint32_t)ty & 7u, sb = (uint32_t)tx & 7u;   // swizzle phases of this thread's rows
    const uint32_t a_base = (uint32_t)ty * 8u;      // float4 index of query slot ty, chunk 0 (row pitch = 8 float4)
    const uint32_t b_base = (uint32_t)tx * 8u;

    uint32_t it = 0;
    for (uint64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const uint32_t qb = (uint32_t)(item % a.n_qblocks);
        const uint64_t t = a.tile_begin + item / a.n_qblocks;

        float acc[TM][TN];
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

        for (uint32_t kc = 0; kc < a.n_kchunks; ++kc, ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            mbar_wait(&full_bar[s], ph);
            const float4 *As = reinterpret_cast<const float4 *>(ring + (size_t)s * STAGE_BYTES);
            const float4 *Bs = As + A_BYTES / 16;
#pragma unroll 2
            for (uint32_t c = 0; c < GS_BK / 4; ++c) {
                float4 av[TM], bv[TN];
                const float4 *ap = As + a_base + (c ^ sa);
                const float4 *bp = Bs + b_base + (c ^ sb);
#pragma unroll
                for (int i = 0; i < TM; ++i) av[i] = ap[i * 16 * 8];
#pragma unroll
                for (int j = 0; j < TN; ++j) bv[j] = bp[j * 16 * 8];
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) {
                        acc[i][j] = fmaf(av[i].x, bv[j].x, acc[i][j]);
                        acc[i][j] = fmaf(av[i].y, bv[j].y, acc[i][j]);
                        acc[i][j] = fmaf(av[i].z, bv[j].z, acc[i][j]);
                        acc[i][j] = fmaf(av[i].w, bv[j].w, acc[i][j]);
                    }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);   // this warp is done reading the stage
        }

        // ---- epilogue: threshold filter, candidates -> HBM (thresholds are read here, not held across the k loop:
        //      the wider accumulator tile needs the registers) ----
        const uint64_t row0 = t * BN + tx;
        uint32_t any = 0;
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            const float thr = __ldg(a.thr + qb * BM + i * 16 + ty);
            float best = acc[i][0];
#pragma unroll
            for (int j = 1; j < TN; ++j) best = fmaxf(best, acc[i][j]);
            any |= (fmaf(-0.5f, best, 0.5f) <= thr) ? (1u << i) : 0u;
        }
        if (any) {   // rare after the first phases
#pragma unroll
            for (int i = 0; i < TM; ++i) {
                if (!((any >> i) & 1u)) continue;
                const uint32_t q = qb * BM + i * 16 + ty;
                const float thr = __ldg(a.thr + q);
                uint32_t cnt = 0;
#pragma unroll
                for (int j = 0; j < TN; ++j)
                    cnt += (fmaf(-0.5f, acc[i][j], 0.5f) <= thr && row0 + j * 16 < a.n_rows) ? 1u : 0u;
                if (!cnt) continue;
                unsigned pos = atomicAdd(a.count + q, cnt);
                uint64_t *my_cand = a.cand + (size_t)q * a.cap;
#pragma unroll
                for (int j = 0; j < TN; ++j) {
                    const float dist = fmaf(-0.5f, acc[i][j], 0.5f);
                    const uint64_t row = row0 + j * 16;
                    if (dist <= thr && row < a.n_rows) {
                        if (pos < a.cap) my_cand[pos] = make_key(dist, (uint32_t)row);   // row index; id swapped in by the select kernel
                        ++pos;
                    }
                }
            }
        }
    }
}

// ---- query prep: fp32 [b][dim] -> unit length (f64 norm, as csgpu_build does for rows) -> [b_pad][dim_pad]
// (rows >= b and columns >= dim are zero). flags[q] = 1 if the query has zero norm.
static __global__ void prep_queries_f32_kernel(const float *__restrict__ q, uint32_t q_pitch, uint32_t b, uint32_t dim, uint32_t dim_pad,
                                               float *__restrict__ out, uint32_t b_pad, uint8_t *__restrict__ flags,
                                               float *__restrict__ thr, unsigned *__restrict__ count)
{
    const int lane = threadIdx.x & 31;
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= b_pad) return;
    if (w >= b) {
        for (uint32_t c = lane; c < dim_pad; c += 32) out[(size_t)w * dim_pad + c] = 0.f;
        if (lane == 0) { flags[w] = 0; thr[w] = -1.f; count[w] = 0; }
        return;
    }
    double ss = 0.0;
    for (uint32_t c = lane; c < dim; c += 32) { const float x = q[(size_t)w * q_pitch + c]; ss += (double)x * x; }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) ss += __shfl_xor_sync(FULL, ss, m);
    const bool zero = !(ss > 0.0);
    const double inv = zero ? 0.0 : 1.0 / sqrt(ss);
    for (uint32_t c = lane; c < dim_pad; c += 32)
        out[(size_t)w * dim_pad + c] = c < dim ? (float)(q[(size_t)w * q_pitch + c] * inv) : 0.f;
    if (lane == 0) { flags[w] = zero ? 1 : 0; thr[w] = zero ? -1.f : __int_as_float(0x7f800000); count[w] = 0; }
}

}  // namespace csgpu
