// scan_multi.cu — instantiations + launcher of the multi-query scan (own translation unit so the
// library builds in parallel).
#include <algorithm>

#include "index.h"
#include "scan_multi.cuh"

namespace csgpu {

// Two instantiations per dim: one group of MQ = 8 queries (R = 4 rows per warp iteration, R = 2 from 768-d up), and —
// round 2 — NG = 2 groups for 9..16 queries: the rows a warp has loaded are scored against both groups, so the reference's
// default hybrid search (<= 9 query variants x limit 200, /root/reference/src/search/mod.rs:498-511) is ONE pass over HBM
// instead of an 8-query pass plus a single scan.
constexpr int MQ = 8;
constexpr int MQ_MAX = 16;

template <int V, int R, int NG, int E>
static cudaError_t launch_one(const MultiArgs &a, uint32_t grid, cudaStream_t st)
{
    auto kern = scan_multi_topk_kernel<V, R, MQ, NG, E>;
    constexpr int MQT = MQ * NG;
    const size_t smem = (size_t)MQT * a.dim4 * sizeof(float4) + (size_t)SCAN_WARPS * multi_warp_keys<E>(MQT) * sizeof(uint64_t) +
                        (size_t)SCAN_WARPS * 32 * E * sizeof(uint64_t) + (size_t)SCAN_WARPS * MQT * sizeof(uint32_t);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    kern<<<grid, SCAN_THREADS, smem, st>>>(a);
    g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

// k > 32: one CTA-shared candidate buffer per query (a.kpad = its capacity)
template <int V, int R, int NG>
static cudaError_t launch_cta(const MultiArgs &a, uint32_t grid, cudaStream_t st)
{
    auto kern = scan_multi_cta_topk_kernel<V, R, MQ, NG>;
    constexpr int MQT = MQ * NG;
    const size_t smem = (size_t)MQT * a.dim4 * sizeof(float4) + (size_t)MQT * a.kpad * sizeof(uint64_t);
    // the ceiling of this instantiation (k = 256 -> kpad = 1024), not this launch's own size: the attribute is per
    // function, and concurrent searches with different k must not undercut each other's launches
    const size_t smem_max = (size_t)MQT * 32 * V * sizeof(float4) + (size_t)MQT * 1024 * sizeof(uint64_t);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(smem, smem_max));
    if (e != cudaSuccess) return e;
    kern<<<grid, SCAN_THREADS, smem, st>>>(a);
    g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

template <int V, int R>
static cudaError_t launch_e(const MultiArgs &a, uint32_t grid, cudaStream_t st)
{
    if (a.nq > (uint32_t)MQ) {
        if (a.k > 32) return launch_cta<V, R, 2>(a, grid, st);
        return launch_one<V, R, 2, 1>(a, grid, st);
    }
    if (a.k > 32) return launch_cta<V, R, 1>(a, grid, st);
    return launch_one<V, R, 1, 1>(a, grid, st);
}

bool multi_scan_supported(uint32_t dim4, uint32_t k)
{
    if (dim4 % 32 != 0 || k > 256) return false;
    const uint32_t V = dim4 / 32;
    return V == 1 || V == 2 || V == 3 || V == 4 || V == 6 || V == 8;
}

uint32_t multi_scan_max_queries() { return MQ_MAX; }

uint32_t multi_scan_rows_per_iter(uint32_t dim4, uint32_t) { return (dim4 / 32 <= 4) ? 4u : 2u; }

// capacity of one (CTA, query) candidate buffer for k > 32. Up to 8 queries: the single-query kernel's ctabuf_cap. 9..16
// queries: as small as the kernel's invariants allow (k + 2 sync intervals while streaming, k + one key per CTA in the last
// CTA's column walk), so that sixteen buffers still leave room for two CTAs per SM at the reference's k = 200.
uint32_t multi_scan_cap(uint32_t k, uint32_t nq, uint32_t grid)
{
    if (k <= 32) return 32;
    if (nq <= 8) return ctabuf_cap(k);
    const uint32_t slack = 4u * 4u * SCAN_WARPS;   // SYNC_IT x R x warps of the NG = 2 kernel (R = 4 is the larger case)
    return pow2_at_least(std::max(k + 2 * slack, k + grid), 512);
}

// resident CTAs per SM the launch should be sized for (dynamic shared memory is what limits it)
uint32_t multi_scan_ctas_per_sm(uint32_t dim4, uint32_t k, uint32_t nq, uint32_t kpad)
{
    const uint32_t mq = nq > 8 ? 16u : 8u;
    size_t smem = (size_t)mq * dim4 * sizeof(float4);
    if (k > 32) smem += (size_t)mq * kpad * sizeof(uint64_t);
    else smem += (size_t)SCAN_WARPS * multi_warp_keys<1>(mq) * sizeof(uint64_t) + (size_t)SCAN_WARPS * 32 * sizeof(uint64_t) + (size_t)SCAN_WARPS * mq * sizeof(uint32_t);
    return smem + 1024 <= (227u * 1024u) / 2 ? 2u : 1u;
}

cudaError_t launch_scan_multi(const MultiArgs &a, uint32_t grid, cudaStream_t st)
{
    switch (a.dim4 / 32) {
        case 1: return launch_e<1, 4>(a, grid, st);
        case 2: return launch_e<2, 4>(a, grid, st);
        case 3: return launch_e<3, 4>(a, grid, st);
        case 4: return launch_e<4, 4>(a, grid, st);
        case 6: return launch_e<6, 2>(a, grid, st);
        case 8: return launch_e<8, 2>(a, grid, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace csgpu
