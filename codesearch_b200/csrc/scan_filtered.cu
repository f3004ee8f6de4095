// scan_filtered.cu — instantiations + launcher of the pre-filtering single-query scan (csgpu_search_filtered):
// scan_topk_kernel<..., FILT = true> (scan.cuh, scan_rows_filtered). Own translation unit so the library builds
// in parallel.
#include "index.h"
#include "scan.cuh"

namespace csgpu {

template <int V, bool EXACT, bool BIG, int OCC>
static cudaError_t launch_f_v(const ScanArgs &a, uint32_t grid, size_t smem, cudaStream_t st)
{
    constexpr int R = (V <= 2) ? 8 : (V <= 4 ? 4 : 2);
    auto kern = scan_topk_kernel<V, EXACT, R, BIG, OCC, 0, true>;
    if (grid == 0) { cudaFuncAttributes fa; return cudaFuncGetAttributes(&fa, kern); }   // preload only
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    kern<<<grid, SCAN_THREADS, smem, st>>>(a);
    count_launch();
    return cudaGetLastError();
}

template <bool BIG, int OCC>
static cudaError_t launch_f_b(const ScanArgs &a, uint32_t grid, size_t smem, cudaStream_t st)
{
    const uint32_t V = (a.dim4 + 31) / 32;
    const bool exact = (a.dim4 % 32) == 0;
#define CS_CASE(v)                                                                 \
    case v: return exact ? launch_f_v<v, true, BIG, OCC>(a, grid, smem, st)         \
                         : launch_f_v<v, false, BIG, OCC>(a, grid, smem, st);
    switch (V) {
        CS_CASE(1) CS_CASE(2) CS_CASE(3) CS_CASE(4) CS_CASE(5) CS_CASE(6) CS_CASE(7) CS_CASE(8)
        default: return cudaErrorInvalidValue;
    }
#undef CS_CASE
}

cudaError_t launch_scan_filtered(const ScanArgs &a, uint32_t grid, size_t smem, cudaStream_t st)
{
    return a.k > 32 ? launch_f_b<true, 2>(a, grid, smem, st) : launch_f_b<false, 2>(a, grid, smem, st);
}

}  // namespace csgpu
