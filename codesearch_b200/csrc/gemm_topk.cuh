// gemm_topk.cuh — batched search on the opt-in bf16 index: dense Q x C^T on the 5th-gen tensor
// cores (tcgen05.mma, accumulators in TMEM) with the top-k filter fused into the epilogue.
// SURVEY.md §8a row A9 / BASELINE config C3 (10M x 384, B = 1024, k = 100). No reference
// counterpart (the reference has one f32 ANN path, /root/reference/src/vectordb/store.rs:446-459);
// semantics are the same ranking on bf16-rounded unit vectors, fp32 accumulate.
//
// Shapes. One CTA = 128 queries (MMA M = 128, one TMEM lane per query) x a stream of 256-row
// corpus tiles (MMA N = 256) x K = dim in 64-element chunks (UMMA_K = 16 -> 4 MMAs per chunk).
//   smem: the CTA's 128 queries, all of K, resident   [dim/64][128 x 64] bf16  (96 KB at D = 384)
//         ring of STAGES corpus chunks                 [256 x 64] bf16 = 32 KB each
//         both in the canonical K-major SWIZZLE_128B layout that TMA writes and UMMA reads.
//   TMEM: 2 accumulator buffers x 256 fp32 columns = all 512 columns (epilogue of tile i overlaps
//         the MMAs of tile i+1).
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM alloc + MMA issuer (one elected
// lane), warps 2-9 = epilogue: warp w reads TMEM lanes 32*(w%4)..+31 (32 queries) and the column half
// (w-2)/4 of the accumulator (128 rows of the tile), so two threads share a query (measured 3 % faster at k = 100
// than one thread per query, which CSGPU_TC_EPI=1 still selects).
//
// Epilogue = progressive-threshold filter. Each epilogue thread holds its query's current threshold
// distance in a register, reads its lane's 128 scores with tcgen05.ld, 32 columns at a time, and
// appends the (distance, row) keys that pass to ITS OWN segment of
// the query's candidate buffer in HBM: segment = (CTA group, column half), position counter in a
// register — no atomics, no second pass over TMEM; the counts are written once when the kernel ends.
// Between row phases (sizes growing geometrically) a select kernel reduces each query's segments to its
// exact top-k and publishes the new threshold, so after the first phases only ~k * phase_growth rows per
// query pass. Exactness: the threshold is always the k-th best of a SUBSET of the rows, hence never
// tighter than the final k-th best; ties pass (<=) and are ordered by the full (distance, id) key in the
// select kernel. If a segment overflows (adversarially ordered data) the host halves the range and
// retries, which terminates because a range of one tile (128 rows per segment) cannot overflow.
//
// CTA -> work: query block qb = cta % QB, group = cta / QB; the QB CTAs of a group walk the same
// tiles at the same time, so a corpus tile is read from HBM once and served to the other QB-1
// CTAs from L2.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include "scan_tma.cuh"  // mbarrier helpers
#include "topk.cuh"

namespace csgpu {

constexpr int GT_BLOCK_M = 128;   // queries per CTA
constexpr int GT_BLOCK_N = 256;   // corpus rows per tile
constexpr int GT_BLOCK_K = 64;    // bf16 elements per K chunk (= 128 B = one swizzle row)
constexpr int GT_UMMA_K = 16;


constexpr uint32_t GT_SURV = 1024;   // slots at the front of a query's candidate block reserved for the survivors of the last select (= CSGPU_MAX_K)
constexpr uint32_t GT_STAGE_BYTES = GT_BLOCK_N * GT_BLOCK_K * 2;   // 32 KB
constexpr uint32_t GT_QCHUNK_BYTES = GT_BLOCK_M * GT_BLOCK_K * 2;  // 16 KB

struct GemmTopkArgs {
    uint64_t n_rows;           // rows in the matrix (tensor map extent)
    uint64_t tile_begin, tile_end;  // this phase scans tiles [tile_begin, tile_end)
    uint32_t n_kchunks;        // dim / 64
    uint32_t n_qblocks;        // QB
    const float *thr;          // [QB*128] threshold distance per query (< 0: inactive, +inf: pass all)
    uint64_t *cand;            // [QB*128][cap] candidate keys (okey(distance) << 32 | ROW index)
    unsigned *count;           // [QB*128] entries in cand (may exceed cap => overflow)
    uint32_t cap;
    uint32_t prefetch_tiles;   // bf16 kernel: L2 prefetch distance in tiles of a CTA's walk (0 = off)
    // tensor-core kernel only — segmented candidate layout: query q owns cand[q * stride, +stride):
    //   [0, GT_SURV)                     survivors of the previous select (exact keys with chunk ids; count[q] of them)
    //   [GT_SURV + s * seg_len, +seg_len) segment s = group * 2 + column half, seg_count[q * n_seg + s] entries
    unsigned *seg_count = nullptr;   // [QB*128][n_seg], WRITTEN (not accumulated) by every launch
    unsigned *overflow = nullptr;    // set to 1 when a segment would overflow
    uint32_t stride = 0, seg_len = 0, n_seg = 0;
    uint32_t q_rows = GT_BLOCK_M;    // tf32 kernel: query rows per chunk load (gemm_tf32.cuh)
};

// ---- tcgen05 wrappers ---------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_dst, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 x bf16 -> f32, cta_group::1
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 2-D TMA tile load (inner coordinate = element column, outer = row), completes on an mbarrier
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int32_t c_inner, int32_t c_outer)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer)
        : "memory");
}

// predicated 8-byte global store without a branch
__device__ __forceinline__ void st_global_pred(uint64_t *p, uint64_t v, bool pred)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.global.u64 [%0], %1;\n\t}" ::"l"(p), "l"(v), "r"((uint32_t)pred) : "memory");
}

// L2 prefetch of a 2-D TMA tile (no shared-memory destination, no barrier)
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap *map, int32_t c_inner, int32_t c_outer)
{
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c_inner), "r"(c_outer) : "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// start>>4 [0,14) | LBO>>4 [16,30) (unused for swizzled K-major) | SBO>>4 [32,46) = 1024 B (8 rows x 128 B)
// | version=1 [46,48) | layout_type=2 (SWIZZLE_128B) [61,64)
__device__ __forceinline__ uint64_t make_sw128_kmajor_desc(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// cute::UMMA::InstrDescriptor: c_format=F32(1) [4,6) | a_format=BF16(1) [7,10) | b_format=BF16(1) [10,13)
// | a_major=K(0) [15] | b_major=K(0) [16] | N>>3 [17,23) | M>>4 [24,29)
__host__ __device__ constexpr uint32_t make_idesc_bf16_f32(uint32_t M, uint32_t N)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// One 32-column chunk of one thread's query: max -> one compare in the common case; a hit (lane-divergent, rare
// after the first phases) appends the passing keys to the thread's own segment with branch-free predicated
// stores. The key carries the ROW index; the select kernel swaps in the chunk id (no dependent load in here).
__device__ __forceinline__ void gt_epilogue_chunk(const uint32_t (&r)[32], uint64_t row0, uint64_t n_rows, float thr,
                                                  uint64_t *my_seg, uint32_t seg_len, uint32_t &pos)
{
    float m0 = __uint_as_float(r[0]), m1 = __uint_as_float(r[1]), m2 = __uint_as_float(r[2]), m3 = __uint_as_float(r[3]);
#pragma unroll
    for (int j = 4; j < 32; j += 4) {   // four independent chains
        m0 = fmaxf(m0, __uint_as_float(r[j])); m1 = fmaxf(m1, __uint_as_float(r[j + 1]));
        m2 = fmaxf(m2, __uint_as_float(r[j + 2])); m3 = fmaxf(m3, __uint_as_float(r[j + 3]));
    }
    const float best = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
    if (fmaf(-0.5f, best, 0.5f) <= thr) {
        const uint32_t rbase = (uint32_t)row0;
        const uint32_t lim = (uint32_t)min((uint64_t)32, n_rows - min(n_rows, row0));
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float dist = fmaf(-0.5f, __uint_as_float(r[j]), 0.5f);
            const bool pass = (dist <= thr) & ((uint32_t)j < lim);
            st_global_pred(my_seg + min(pos, seg_len - 1), make_key(dist, rbase + j), pass & (pos < seg_len));
            pos += pass ? 1u : 0u;
        }
    }
}

// dynamic smem: [n_kchunks][16 KB] queries | [STAGES][32 KB] corpus ring   (1024-B aligned)
// HALVES = threads per query in the epilogue (1: warps 2-5, 192 threads; 2: warps 2-9, 320 threads, column halves).
template <int STAGES, int HALVES>
__global__ void __launch_bounds__(64 + 128 * HALVES, 1)
gemm_topk_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_c, const GemmTopkArgs a)
{
    extern __shared__ __align__(1024) unsigned char gt_smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], q_bar, tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // 1024-B alignment of the dynamic region (swizzle atoms are 1024 B)
    unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(gt_smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *q_smem = base;
    unsigned char *ring = base + (size_t)a.n_kchunks * GT_QCHUNK_BYTES;

    const uint32_t qb = blockIdx.x % a.n_qblocks;
    const uint32_t group = blockIdx.x / a.n_qblocks;
    const uint32_t n_groups = gridDim.x / a.n_qblocks;
    const bool cta_active = group < n_groups;   // leftover CTAs (gridDim % QB) idle

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&q_bar, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4 * HALVES); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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
                mbar_expect_tx(&q_bar, a.n_kchunks * GT_QCHUNK_BYTES);
                for (uint32_t kc = 0; kc < a.n_kchunks; ++kc)
                    tma_load_2d(q_smem + (size_t)kc * GT_QCHUNK_BYTES, &map_q, &q_bar, (int32_t)(kc * GT_BLOCK_K), (int32_t)(qb * GT_BLOCK_M));
                uint32_t it = 0;
                for (uint64_t t = a.tile_begin + group; t < a.tile_end; t += n_groups) {
                    // The QB CTAs of a group read the same tile at about the same time; the first one pulls the
                    // tile it will need prefetch_tiles iterations from now into L2, so the demand loads of all
                    // QB CTAs hit L2 (DRAM latency off the MMA's critical path, one DRAM read per tile).
                    if (qb == 0 && a.prefetch_tiles) {
                        const uint64_t tp = t + (uint64_t)a.prefetch_tiles * n_groups;
                        if (tp < a.tile_end)
                            for (uint32_t kc = 0; kc < a.n_kchunks; ++kc)
                                tma_prefetch_2d(&map_c, (int32_t)(kc * GT_BLOCK_K), (int32_t)(tp * GT_BLOCK_N));
                    }
                    for (uint32_t kc = 0; kc < a.n_kchunks; ++kc, ++it) {
                        const int s = it % STAGES;
                        const uint32_t ph = (it / STAGES) & 1;
                        mbar_wait_backoff(&empty_bar[s], ph ^ 1, 64);
                        mbar_expect_tx(&full_bar[s], GT_STAGE_BYTES);
                        tma_load_2d(ring + (size_t)s * GT_STAGE_BYTES, &map_c, &full_bar[s], (int32_t)(kc * GT_BLOCK_K), (int32_t)(t * GT_BLOCK_N));
                    }
                }
            }
            __syncwarp();
        } else if (warp == 1) {
            // ================= MMA issuer =================
            if (lane == 0) {
                constexpr uint32_t idesc = make_idesc_bf16_f32(GT_BLOCK_M, GT_BLOCK_N);
                mbar_wait(&q_bar, 0);
                tc_fence_after();
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
                        const uint64_t adesc = make_sw128_kmajor_desc(smem_u32(q_smem + (size_t)kc * GT_QCHUNK_BYTES));
                        const uint64_t bdesc = make_sw128_kmajor_desc(smem_u32(ring + (size_t)s * GT_STAGE_BYTES));
#pragma unroll
                        for (uint32_t k = 0; k < GT_BLOCK_K / GT_UMMA_K; ++k) {
                            // advance 16 elements = 32 B along K inside the swizzle atom: +2 in (addr >> 4) units
                            umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kc | k) ? 1u : 0u);
                        }
                        umma_commit(&empty_bar[s]);          // smem slot reusable once these MMAs retire
                    }
                    umma_commit(&tfull_bar[acc]);            // accumulator ready for the epilogue
                }
            }
            __syncwarp();
        } else {
            // ================= epilogue: two threads (column halves) per query =================
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
                // One chunk at a time: keeping the tcgen05.ld of chunk c+1 in flight while chunk c is tested was
                // measured 40-120 % SLOWER (profiles/r01_bf16_epilogue_modes.txt), so the load is waited for at once.
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


// ---- query prep: fp32 [b][dim] -> unit length -> bf16 [QB*128][dim] (rows >= b zero) -------------
// flags[q] = 1 if the query has zero norm.
static __global__ void prep_queries_bf16_kernel(const float *__restrict__ q, uint32_t b, uint32_t dim,
                                         __nv_bfloat16 *__restrict__ out, uint32_t b_pad, uint8_t *__restrict__ flags,
                                         float *__restrict__ thr, unsigned *__restrict__ count)
{
    const int lane = threadIdx.x & 31;
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= b_pad) return;
    if (w >= b) {
        for (uint32_t c = lane; c < dim; c += 32) out[(size_t)w * dim + c] = __float2bfloat16(0.f);
        if (lane == 0) { flags[w] = 0; thr[w] = -1.f; count[w] = 0; }
        return;
    }
    double ss = 0.0;
    for (uint32_t c = lane; c < dim; c += 32) { const float x = q[(size_t)w * dim + c]; ss += (double)x * x; }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) ss += __shfl_xor_sync(FULL, ss, m);
    const bool zero = !(ss > 0.0);
    const double inv = zero ? 0.0 : 1.0 / sqrt(ss);
    for (uint32_t c = lane; c < dim; c += 32) out[(size_t)w * dim + c] = __float2bfloat16((float)(q[(size_t)w * dim + c] * inv));
    if (lane == 0) { flags[w] = zero ? 1 : 0; thr[w] = zero ? -1.f : __int_as_float(0x7f800000); count[w] = 0; }
}

}  // namespace csgpu
