// tune_scan.cu — standalone tuning harness for the single-query scan (not part of libcsgpu.so).
// Sweeps rows-in-flight R, CTAs/SM, load flavour and the TMA-bulk pipeline on a 10M x 384 corpus,
// checks every variant returns the same keys as the baseline, prints GB/s (CUDA events).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo tune_scan.cu -o tune_scan
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <string>
#include "scan.cuh"
#include "scan_tma.cuh"
#include "synth.cuh"
using namespace csgpu;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

struct Variant { std::string name; void (*launch)(const ScanArgs &, int sms, cudaStream_t); };

template <int R, int OCC, int LD>
static void launch_ldg(const ScanArgs &a, int sms, cudaStream_t st)
{
    scan_topk_kernel<3, true, R, false, OCC, LD><<<sms * OCC, SCAN_THREADS, SCAN_SMALL_SMEM, st>>>(a);
}
template <int R, int OCC, int LD>
static void launch_ldg_static(const ScanArgs &a, int sms, cudaStream_t st)   // fixed-stride row split (DYN = false)
{
    scan_topk_kernel<3, true, R, false, OCC, LD, false, false><<<sms * OCC, SCAN_THREADS, SCAN_SMALL_SMEM, st>>>(a);
}
template <int TILE, int STAGES>
static void launch_tma(const ScanArgs &a, int sms, cudaStream_t st)
{
    auto k = scan_topk_tma_kernel<3, TILE, STAGES, false>;
    size_t smem = (size_t)STAGES * TILE * 1536 + SCAN_WARPS * 32 * 8;
    static bool set = false;
    if (!set) { CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); set = true; }
    k<<<sms, SCAN_THREADS + 32, smem, st>>>(a);
}
template <int TILE, int STAGES>
static void launch_tma2(const ScanArgs &a, int sms, cudaStream_t st)   // 2 CTAs / SM
{
    auto k = scan_topk_tma_kernel<3, TILE, STAGES, false>;
    size_t smem = (size_t)STAGES * TILE * 1536 + SCAN_WARPS * 32 * 8;
    static bool set = false;
    if (!set) { CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); set = true; }
    k<<<sms * 2, SCAN_THREADS + 32, smem, st>>>(a);
}

int main(int argc, char **argv)
{
    const uint64_t n = argc > 1 ? strtoull(argv[1], 0, 10) : 10000000ull;
    const int iters = argc > 2 ? atoi(argv[2]) : 20;
    const uint32_t dim4 = 96, k = 10;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    float4 *rows; uint32_t *ids; uint8_t *status; float *q; uint64_t *cand, *out, *ref; unsigned *ticket;
    CK(cudaMalloc(&rows, n * dim4 * 16)); CK(cudaMalloc(&ids, n * 4)); CK(cudaMalloc(&status, n)); CK(cudaMemset(status, 0, n));
    CK(cudaMalloc(&q, 384 * 4)); CK(cudaMalloc(&cand, 148 * 8 * 1024 * 8)); CK(cudaMalloc(&out, 1024 * 8)); CK(cudaMalloc(&ref, 1024 * 8));
    CK(cudaMalloc(&ticket, 256)); CK(cudaMemset(ticket, 0, 256));
    synth_rows_kernel<<<sms * 16, 256>>>(rows, ids, 1234, 0, n, dim4, 0);
    normalise_rows_kernel<<<sms * 8, 256>>>(rows, status, 0, n, dim4);
    synth_rows_kernel<<<1, 96>>>((float4 *)q, nullptr, 4321, 0, 1, dim4, 0);
    CK(cudaDeviceSynchronize());
    ScanArgs a{};
    a.rows = rows; a.ids = ids; a.n_rows = n; a.dim4 = dim4; a.q = q; a.k = k; a.kpad = 32;
    a.bitmap = nullptr; a.n_bits = 0; a.zero_ids = nullptr; a.n_zero = 0; a.cand = cand; a.ticket = ticket; a.out_keys = out;

    std::vector<Variant> vs = {
        {"ldg R4 occ2 ld0 (baseline)", launch_ldg<4, 2, 0>},
        {"static R4 occ2 ld0", launch_ldg_static<4, 2, 0>},
        {"static R6 occ2 ld0", launch_ldg_static<6, 2, 0>},
        {"ldg R5 occ2 ld0", launch_ldg<5, 2, 0>},
        {"ldg R4 occ2 ld1", launch_ldg<4, 2, 1>},
        {"ldg R4 occ2 ld2", launch_ldg<4, 2, 2>},
        {"ldg R4 occ2 ld3 (L2::256B)", launch_ldg<4, 2, 3>},
        {"ldg R4 occ2 ld4 (L2::evict_first)", launch_ldg<4, 2, 4>},
        {"ldg R4 occ2 ld5 (both)", launch_ldg<4, 2, 5>},
        {"ldg R6 occ2 ld3", launch_ldg<6, 2, 3>},
        {"ldg R2 occ4 ld3", launch_ldg<2, 4, 3>},
        {"ldg R2 occ2 ld0", launch_ldg<2, 2, 0>},
        {"ldg R2 occ3 ld0", launch_ldg<2, 3, 0>},
        {"ldg R2 occ4 ld0", launch_ldg<2, 4, 0>},
        {"ldg R4 occ1 ld0", launch_ldg<4, 1, 0>},
        {"ldg R4 occ3 ld0", launch_ldg<4, 3, 0>},
        {"ldg R6 occ2 ld0", launch_ldg<6, 2, 0>},
        {"ldg R8 occ1 ld0", launch_ldg<8, 1, 0>},
        {"ldg R8 occ2 ld0", launch_ldg<8, 2, 0>},
        {"ldg R1 occ4 ld0", launch_ldg<1, 4, 0>},
        {"ldg R1 occ6 ld0", launch_ldg<1, 6, 0>},
        {"ldg R2 occ6 ld0", launch_ldg<2, 6, 0>},
        {"tma tile16 x8 occ1", launch_tma<16, 8>},
        {"tma tile32 x4 occ1", launch_tma<32, 4>},
        {"tma tile32 x3 occ1", launch_tma<32, 3>},
        {"tma tile16 x4 occ2", launch_tma2<16, 4>},
        {"tma tile8 x8 occ2", launch_tma2<8, 8>},
        {"ldg R4 occ2 ld0 (baseline again)", launch_ldg<4, 2, 0>},
        {"ldg R6 occ2 ld0 (again)", launch_ldg<6, 2, 0>},
    };
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    std::vector<uint64_t> href(k), hout(k);
    for (size_t v = 0; v < vs.size(); ++v) {
        CK(cudaMemset(out, 0, k * 8));
        for (int i = 0; i < 3; ++i) vs[v].launch(a, sms, 0);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%-30s FAILED: %s\n", vs[v].name.c_str(), cudaGetErrorString(e)); return 1; }
        CK(cudaEventRecord(e0));
        for (int i = 0; i < iters; ++i) vs[v].launch(a, sms, 0);
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= iters;
        CK(cudaMemcpy(hout.data(), out, k * 8, cudaMemcpyDeviceToHost));
        if (v == 0) href = hout;
        bool same = hout == href;
        printf("%-36s %8.4f ms  %8.1f GB/s  %s\n", vs[v].name.c_str(), ms, n * 1536.0 / ms / 1e6, same ? "keys==baseline" : "KEYS DIFFER");
        fflush(stdout);
    }
    return 0;
}
