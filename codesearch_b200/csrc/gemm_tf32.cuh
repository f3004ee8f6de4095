// gemm_tf32.cuh — batches on the DEFAULT fp32 index, contracted on the 5th-gen tensor cores straight off the fp32 rows
// (tcgen05.mma kind::tf32, accumulators in TMEM) — as a FILTER; the survivors are rescored in exact fp32 by
// select_sorted_kernel<.., RESCORE = true> (rescore.cuh), so ids AND distances stay bit-identical to csgpu_search.
// BASELINE configs[2] (10M x 384, 1024 queries x top-100) and every multi-query route above a handful of queries.
// No reference counterpart: the reference answers query variants one arroy search at a time
// (/root/reference/src/search/mod.rs:508-511, /root/reference/src/vectordb/store.rs:446-459).
//
// Why this shape. kind::tf32 reads the 32-bit containers as they sit in HBM and uses sign + 8 exponent + 10 mantissa bits:
// no shadow copy of the corpus (the bf16 tensor prefilter of rescore.cuh needs +50 % HBM), no conversion pass. The fp32
// SIMT kernel (gemm_simt.cuh) is bound by the FP32 FMA pipe at ~52 TFLOP/s; the tensor pipe runs the same contraction
// ~15x faster: up to 128 queries the pass is bound by HBM (one pass over the rows, ~2.3 ms at 10M x 384; ncu: 85 % of the
// DRAM peak), above by the tf32 tensor pipe itself (half the bf16 rate; ncu at 1024 queries: pipe 80 % active, its memory
// side 88 %, L2 -> SM ~10 TB/s because every 128-query block re-reads the row tiles from L2).
//
// One CTA = 128 queries (MMA M = 128, one TMEM lane per query) x a stream of 256-row tiles (MMA N = 256) x K = dim in
// 32-float chunks (128 B = one swizzle row; UMMA_K = 8 -> 4 MMAs per chunk). 128 queries x 384 floats = 192 KB cannot
// stay resident next to a ring, so BOTH operands stream: stage = [q_rows x 128 B query chunk | 256 x 128 B row chunk],
// 48 KB, 4 stages; the query chunk is an L2 hit (1.5 KB per query, re-read once per row tile). Batches of fewer than 128
// queries load only round_up(nq, 8) query rows per chunk (the TMA box is that small; the rest of the 16 KB stays zero from
// the kernel prologue, so the unused TMEM lanes hold 0.0 = distance 0.5 and never pass an inactive query's threshold -1).
// Warp roles, TMEM double buffering, segmented candidate layout and epilogue: exactly gemm_topk_kernel's (gemm_topk.cuh).
//
// Error bound of the filter (TF_MARGIN, the tf32 counterpart of rescore.cuh's TC_MARGIN). Rows r and the query q are fp32
// unit vectors; the tensor core sees r~, q~ = their images with the low 13 mantissa bits dropped (truncation; rounding would
// only be tighter): |x~ - x| < 2^-10 |x| per element, hence ||x~ - x|| < 2^-10 ||x||.
//     |q~.r~ - q.r| <= |q~.(r~ - r)| + |(q~ - q).r| < 2^-10 + 2^-10 = 2^-9                                  (Cauchy-Schwarz)
// tf32 x tf32 products are exact in fp32 (11 x 11 significant bits); accumulation in fp32, at worst truncating:
// <= dim * 2^-23; the scan kernel's fp32 FMA chain: <= dim * 2^-24. For dim <= 1024:
//     |cos_tf - cos_f32| < 1.954e-3 + 1.23e-4 + 6.2e-5 < 2.14e-3      =>      |d_tf - d_f32| < 1.07e-3
// TF_MARGIN = 1.1e-3 (distance units). select_sorted_kernel records the largest |d_tf - d_f32| it has seen when asked to
// (SelectArgs::max_err), and tests/test_gpu_tf32_batch.py asserts it stays below the margin.
#pragma once
#include "gemm_topk.cuh"

namespace csgpu {

constexpr int TF_BLOCK_K = 32;   // fp32 elements per K chunk (= 128 B = one swizzle row)
constexpr int TF_UMMA_K = 8;
constexpr uint32_t TF_A_BYTES = GT_BLOCK_M * 128;   // 16 KB: 128 queries x 32 floats
constexpr uint32_t TF_B_BYTES = GT_BLOCK_N * 128;   // 32 KB: 256 rows x 32 floats
constexpr uint32_t TF_STAGE_BYTES = TF_A_BYTES + TF_B_BYTES;
constexpr uint32_t TF_MAX_DIM = 1024;               // the margin's derivation
constexpr float TF_MARGIN = 1.1e-3f;

// D[tmem] (+)= A[smem desc] * B[smem desc], tf32 x tf32 -> f32, cta_group::1
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// cute::UMMA::InstrDescriptor: c_format=F32(1) [4,6) | a_format=TF32(2) [7,10) | b_format=TF32(2) [10,13)
// | a_major=K(0) [15] | b_major=K(0) [16] | N>>3 [17,23) | M>>4 [24,29)
__host__ __device__ constexpr uint32_t make_idesc_tf32_f32(uint32_t M, uint32_t N)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// dynamic smem: [STAGES][16 KB query chunk | 32 KB row chunk]   (1024-B aligned)
// a.n_kchunks = ceil(dim_pad / 32) (columns past the matrix read as zero: TMA out-of-bounds fill), a.q_rows = query rows
// per chunk load (multiple of 8, <= 128; < 128 only with one query block).
template <int STAGES, int HALVES>
__global__ void __launch_bounds__(64 + 128 * HALVES, 1)
gemm_tf32_topk_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_c, const GemmTopkArgs a)
{
    extern __shared__ __align__(1024) unsigned char tf_smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *ring = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(tf_smem_raw) + 1023) & ~(uintptr_t)1023);

    const uint32_t qb = blockIdx.x % a.n_qblocks;
    const uint32_t group = blockIdx.x / a.n_qblocks;
    const uint32_t n_groups = gridDim.x / a.n_qblocks;
    const bool cta_active = group < n_groups;   // leftover CTAs (gridDim % QB) idle
    const uint32_t q_bytes = a.q_rows * 128u;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4 * HALVES); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (q_bytes < TF_A_BYTES) {   // query rows the TMA box never writes: zero once, for every stage
        const uint32_t n16 = (TF_A_BYTES - q_bytes) / 16;
        for (int s = 0; s < STAGES; ++s) {
            uint4 *z = reinterpret_cast<uint4 *>(ring + (size_t)s * TF_STAGE_BYTES + q_bytes);
            for (uint32_t i = threadIdx.x; i < n16; i += blockDim.x) z[i] = make_uint4(0u, 0u, 0u, 0u);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core's reads
    }
    if (warp == 1) tmem_alloc(&tmem_base_s, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (cta_active) {
        if (warp == 0) {
            // ================= TMA producer =================
            if (lane == 0) {
                asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
                asm volatile("prefetch.tensormap [%0];" ::"l"(&map_c) : "memory");
                uint32_t it = 0;
                for (uint64_t t = a.tile_begin + group; t < a.tile_end; t += n_groups) {
                    // (A contiguous cp.async.bulk.prefetch.L2 of the upcoming tile — DRAM streams it page by page, the strided
                    //  128-byte boxes then hit L2 — was built and measured: 16 queries 2.34 -> 4.11 ms, 1024 queries 12.0 -> 15.3;
                    //  profiles/r02_tf32_prefetch_ab.txt. Not in the code.)
                    for (uint32_t kc = 0; kc < a.n_kchunks; ++kc, ++it) {
                        const int s = it % STAGES;
                        const uint32_t ph = (it / STAGES) & 1;
                        unsigned char *stage = ring + (size_t)s * TF_STAGE_BYTES;
                        mbar_wait_backoff(&empty_bar[s], ph ^ 1, 64);
                        mbar_expect_tx(&full_bar[s], q_bytes + TF_B_BYTES);
                        tma_load_2d(stage + TF_A_BYTES, &map_c, &full_bar[s], (int32_t)(kc * TF_BLOCK_K), (int32_t)(t * GT_BLOCK_N));
                        tma_load_2d(stage, &map_q, &full_bar[s], (int32_t)(kc * TF_BLOCK_K), (int32_t)(qb * GT_BLOCK_M));
                    }
                }
            }
            __syncwarp();
        } else if (warp == 1) {
            // ================= MMA issuer =================
            if (lane == 0) {
                constexpr uint32_t idesc = make_idesc_tf32_f32(GT_BLOCK_M, GT_BLOCK_N);
                uint32_t it = 0, tile_it = 0;
                for (uint64_t t = a.tile_begin + group; t < a.tile_end; t += n_groups, ++tile_it) {
                    const uint32_t acc = tile_it & 1;
                    mbar_wait(&tempty_bar[acc], ((tile_it >> 1) & 1) ^ 1);   // epilogue drained this buffer
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * GT_BLOCK_N;
                    for (uint32_t kc = 0; kc < a.n_kchunks; ++kc, ++it) {
                        const int s = it % STAGES;
                        const uint32_t ph = (it / STAGES) & 1;
                        mbar_wait(&full_bar[s], ph);
                        tc_fence_after();
                        const uint32_t sbase = smem_u32(ring + (size_t)s * TF_STAGE_BYTES);
                        const uint64_t adesc = make_sw128_kmajor_desc(sbase);
                        const uint64_t bdesc = make_sw128_kmajor_desc(sbase + TF_A_BYTES);
#pragma unroll
                        for (uint32_t k = 0; k < TF_BLOCK_K / TF_UMMA_K; ++k) {
                            // advance 8 elements = 32 B along K inside the swizzle atom: +2 in (addr >> 4) units
                            umma_tf32(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kc | k) ? 1u : 0u);
                        }
                        umma_commit(&empty_bar[s]);          // smem slot reusable once these MMAs retire
                    }
                    umma_commit(&tfull_bar[acc]);            // accumulator ready for the epilogue
                }
            }
            __syncwarp();
        } else {
            // ================= epilogue: HALVES threads (column halves) per query =================
            const uint32_t quarter = warp & 3;                       // TMEM lane quarter this warp may access
            const uint32_t half = (uint32_t)(warp - 2) >> 2;         // column block of the accumulator this thread reads
            const uint32_t q_glob = qb * GT_BLOCK_M + quarter * 32 + lane;
            const float thr = a.thr[q_glob];
            const uint32_t seg = group * HALVES + half;
            uint64_t *my_seg = a.cand + (size_t)q_glob * a.stride + GT_SURV + (size_t)seg * a.seg_len;
            uint32_t pos = 0;
            uint32_t tile_it = 0;
            constexpr uint32_t NCH = GT_BLOCK_N / HALVES / 32;       // chunks of 32 columns per thread per tile
            for (uint64_t t = a.tile_begin + group; t < a.tile_end; t += n_groups, ++tile_it) {
                const uint32_t acc = tile_it & 1;
                mbar_wait(&tfull_bar[acc], (tile_it >> 1) & 1);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + acc * GT_BLOCK_N + half * (GT_BLOCK_N / HALVES);
                const uint64_t row0 = t * GT_BLOCK_N + half * (GT_BLOCK_N / HALVES);
#pragma unroll 1
                for (uint32_t c = 0; c < NCH; ++c) {
                    uint32_t r[32];
                    tmem_ld32(taddr + c * 32, r);
                    gt_epilogue_chunk(r, row0 + c * 32, a.n_rows, thr, my_seg, a.seg_len, pos);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            }
            a.seg_count[(size_t)q_glob * a.n_seg + seg] = pos;
            if (pos > a.seg_len) atomicExch(a.overflow, 1u);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace csgpu
