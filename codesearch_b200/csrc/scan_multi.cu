// scan_multi.cu — instantiations + launcher of the multi-query scan (own translation unit so the
// library builds in parallel).
#include <algorithm>

#include "index.h"
#include "scan_multi.cuh"

namespace csgpu {

constexpr int MQ = 8;

template <int V, int R, int E>
static cudaError_t launch_one(const MultiArgs &a, uint32_t grid, cudaStream_t st)
{
    auto kern = scan_multi_topk_kernel<V, R, MQ, E>;
    const size_t smem = (size_t)MQ * a.dim4 * sizeof(float4) + (size_t)SCAN_WARPS * multi_warp_keys<E>(MQ) * sizeof(uint64_t) +
                        (size_t)SCAN_WARPS * 32 * E * sizeof(uint64_t) + (size_t)SCAN_WARPS * MQ * sizeof(uint32_t);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    kern<<<grid, SCAN_THREADS, smem, st>>>(a);
    g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

// k > 32: one CTA-shared candidate buffer per query (a.kpad = its capacity)
template <int V, int R>
static cudaError_t launch_cta(const MultiArgs &a, uint32_t grid, cudaStream_t st)
{
    auto kern = scan_multi_cta_topk_kernel<V, R, MQ>;
    const size_t smem = (size_t)MQ * a.dim4 * sizeof(float4) + (size_t)MQ * a.kpad * sizeof(uint64_t);
    // the ceiling of this instantiation (k = 256 -> kpad = 1024), not this launch's own size: the attribute is per
    // function, and concurrent searches with different k must not undercut each other's launches
    const size_t smem_max = (size_t)MQ * 32 * V * sizeof(float4) + (size_t)MQ * 1024 * sizeof(uint64_t);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(smem, smem_max));
    if (e != cudaSuccess) return e;
    kern<<<grid, SCAN_THREADS, smem, st>>>(a);
    g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

template <int V, int R>
static cudaError_t launch_e(const MultiArgs &a, uint32_t grid, cudaStream_t st)
{
    if (a.k > 32) return launch_cta<V, R>(a, grid, st);
    return launch_one<V, R, 1>(a, grid, st);
}

bool multi_scan_supported(uint32_t dim4, uint32_t k)
{
    if (dim4 % 32 != 0 || k > 256) return false;
    const uint32_t V = dim4 / 32;
    return V == 1 || V == 2 || V == 3 || V == 4 || V == 6 || V == 8;
}

uint32_t multi_scan_max_queries() { return MQ; }

uint32_t multi_scan_ctas_per_sm(uint32_t) { return 2; }

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
