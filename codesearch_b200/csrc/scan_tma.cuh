// scan_tma.cuh — the single-query scan with rows staged through shared memory by the TMA engine.
//
// Same arithmetic, same reduction tree and same top-k selection as scan.cuh (results are
// bit-identical); only the data movement differs: one producer thread streams contiguous row
// tiles with 1-D bulk copies (cp.async.bulk ... mbarrier::complete_tx, SASS UBLKCP) into a
// STAGES-deep ring, consumer warps read rows back with conflict-free LDS.128. This keeps
// STAGES x TILE bytes in flight per SM with a handful of instructions instead of one LDG per
// 512 B, and frees the register file of the R x V float4 staging registers.
#pragma once
#include "scan.cuh"

namespace csgpu {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// Producer-side wait: a warp that only feeds the ring must not burn its scheduler's issue slots while the
// ring is full, so back off between probes (the compute warps on the same SM sub-partition get the slots).
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_backoff(uint64_t *bar, uint32_t parity, unsigned ns)
{
    while (!mbar_try_wait(bar, parity)) __nanosleep(ns);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// dynamic smem: [STAGES][TILE_ROWS * dim4] float4 | top-k scratch (cta_reduce layout) | barriers
template <int V, int TILE_ROWS, int STAGES, bool BIG>
__global__ void __launch_bounds__(SCAN_THREADS + 32, 1) scan_topk_tma_kernel(const ScanArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ bool is_last;
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t dim4 = a.dim4;  // == 32 * V
    const size_t tile_f4 = (size_t)TILE_ROWS * dim4;
    float4 *ring = reinterpret_cast<float4 *>(smem_raw);
    uint64_t *ksm = reinterpret_cast<uint64_t *>(smem_raw + (size_t)STAGES * tile_f4 * sizeof(float4));

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], SCAN_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const uint64_t n = a.n_rows;
    const uint64_t n_tiles = (n + TILE_ROWS - 1) / TILE_ROWS;

    if (warp == SCAN_WARPS) {
        // ===== producer: one elected lane feeds the ring =====
        if (lane == 0) {
            uint32_t it = 0;
            for (uint64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                const uint64_t row0 = t * TILE_ROWS;
                const uint32_t rows = (uint32_t)((n - row0 < (uint64_t)TILE_ROWS) ? (n - row0) : TILE_ROWS);
                const uint32_t bytes = rows * dim4 * (uint32_t)sizeof(float4);
                mbar_expect_tx(&full_bar[s], bytes);
                bulk_g2s(ring + (size_t)s * tile_f4, a.rows + row0 * dim4, bytes, &full_bar[s]);
            }
        }
        __syncwarp();
        // the producer warp takes no part in selection; it still joins the CTA-wide barriers below
    }

    using Sel = typename SelOf<BIG>::type;
    Sel sel;
    float4 qv[V];
    bool qzero = false;
    if (warp < SCAN_WARPS) {
        if constexpr (BIG) sel.init(ksm + (size_t)warp * a.kpad, ksm + (size_t)(SCAN_WARPS + warp) * a.kpad, a.k, a.kpad, lane);
        else sel.init(a.k);
        float ss = 0.f;
#pragma unroll
        for (int j = 0; j < V; ++j) {
            qv[j] = reinterpret_cast<const float4 *>(a.q)[lane + 32 * j];
            ss = fmaf(qv[j].x, qv[j].x, ss); ss = fmaf(qv[j].y, qv[j].y, ss);
            ss = fmaf(qv[j].z, qv[j].z, ss); ss = fmaf(qv[j].w, qv[j].w, ss);
        }
        ss = warp_sum_tree(ss);
        qzero = !(ss > 0.f);
        const float qinv = qzero ? 0.f : 1.0f / sqrtf(ss);
#pragma unroll
        for (int j = 0; j < V; ++j) { qv[j].x *= qinv; qv[j].y *= qinv; qv[j].z *= qinv; qv[j].w *= qinv; }

        // ===== consumers =====
        constexpr int RW = TILE_ROWS / SCAN_WARPS;  // rows per warp per tile
        uint32_t it = 0;
        for (uint64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            mbar_wait(&full_bar[s], ph);
            const float4 *tile = ring + (size_t)s * tile_f4;
            const uint64_t row0 = t * TILE_ROWS;
            float4 x[RW][V];
#pragma unroll
            for (int r = 0; r < RW; ++r)
#pragma unroll
                for (int j = 0; j < V; ++j) x[r][j] = tile[(size_t)(warp * RW + r) * dim4 + lane + 32 * j];
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);  // rows are in registers: hand the slot back early
#pragma unroll
            for (int r = 0; r < RW; ++r) {
                float acc = 0.f;
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    acc = fmaf(x[r][j].x, qv[j].x, acc); acc = fmaf(x[r][j].y, qv[j].y, acc);
                    acc = fmaf(x[r][j].z, qv[j].z, acc); acc = fmaf(x[r][j].w, qv[j].w, acc);
                }
                acc = warp_sum_tree(acc);
                const float dist = qzero ? 0.f : fmaf(-0.5f, acc, 0.5f);
                const uint64_t row = row0 + warp * RW + r;
                if (row < n && okey(dist) <= (uint32_t)(sel.thr >> 32)) {
                    const uint32_t id = a.ids[row];
                    const uint64_t key = make_key(dist, id);
                    if (key < sel.thr && id_allowed(a.bitmap, a.n_bits, id)) sel.insert(key, lane);
                }
            }
        }
    }

    // ---- CTA top-k, then last-CTA merge: identical to scan.cuh, with the producer warp idle ----
    auto reduce_to = [&](uint64_t *dst) {
        if (warp < SCAN_WARPS) {
            sel.flush(lane);
            if constexpr (BIG) {
                uint64_t *mine = ksm + (size_t)warp * a.kpad;
                if (sel.cur != mine)
                    for (uint32_t j = lane; j < a.kpad; j += 32) mine[j] = sel.cur[j];
            } else {
                ksm[warp * 32 + lane] = sel.v;
            }
        }
        cta_sort(ksm, BIG ? SCAN_WARPS * a.kpad : SCAN_WARPS * 32);
        for (uint32_t j = threadIdx.x; j < a.k; j += blockDim.x) dst[j] = ksm[j];
        __syncthreads();
    };
    reduce_to(a.cand + (size_t)blockIdx.x * a.k);

    __threadfence();
    if (threadIdx.x == 0) {
        unsigned tk = atomicAdd(a.ticket, 1u);
        is_last = (tk == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (warp < SCAN_WARPS) {
        if constexpr (BIG) sel.init(ksm + (size_t)warp * a.kpad, ksm + (size_t)(SCAN_WARPS + warp) * a.kpad, a.k, a.kpad, lane);
        else sel.init(a.k);
        const uint64_t total = (uint64_t)gridDim.x * a.k;
        const volatile uint64_t *cand = a.cand;
        for (uint64_t b = (uint64_t)warp * 32; b < total; b += SCAN_WARPS * 32) {
            uint64_t key = (b + lane < total) ? cand[b + lane] : KEY_EMPTY;
            offer_lane_keys(sel, key, lane);
        }
        if (warp == 0 && a.n_zero) {
            uint32_t found = 0;
            for (uint32_t b = 0; b < a.n_zero && found < a.k; b += 32) {
                uint64_t key = KEY_EMPTY;
                if (b + lane < a.n_zero) {
                    uint32_t id = a.zero_ids[b + lane];
                    if (id_allowed(a.bitmap, a.n_bits, id)) key = make_key(0.f, id);
                }
                found += __popc(__ballot_sync(FULL, key != KEY_EMPTY));
                offer_lane_keys(sel, key, lane);
            }
        }
    }
    reduce_to(a.out_keys);
    if (threadIdx.x == 0) *a.ticket = 0;
}

}  // namespace csgpu
