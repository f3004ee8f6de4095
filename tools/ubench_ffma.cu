// ubench_ffma.cu — what is the fp32 SIMT ceiling on B200 for a register-tiled outer-product loop?
// Measures the inner loop of a 128x128 CTA tile (8x8 per thread, operands from shared memory) with
//   mode 0: scalar FFMA                       acc[i][j]    += a[i][k]   * b[j][k]
//   mode 1: packed FFMA2 over k pairs         acc2[i][j]   += a2[i][kk] * b2[j][kk]   (lo/hi = even/odd k)
//   mode 2: packed FFMA2 over j pairs (a dup) acc2[i][j/2] += {a,a}     * {b[j],b[j+1]}
// Result feeds the design of csrc/gemm_simt.cuh (BASELINE config C3, fp32 batched search).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/ubench_ffma.cu -o build/ubench_ffma
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#ifndef UNR
#define UNR 1
#endif
constexpr int KUNROLL = UNR;
constexpr int BK = 32;          // floats per row in the smem tile
constexpr int LDS_STRIDE = 36;  // padded row pitch (floats): 144 B -> conflict-free float4 reads of 8 consecutive rows

__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c)
{
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(b));
    return c;
}
__device__ __forceinline__ unsigned long long pack(float lo, float hi)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float2 unpack(unsigned long long v)
{
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}

template <int MODE>
__global__ void __launch_bounds__(256, 1) k_tile(float *out, int iters)
{
    __shared__ __align__(16) float As[128 * LDS_STRIDE];
    __shared__ __align__(16) float Bs[128 * LDS_STRIDE];
    for (int i = threadIdx.x; i < 128 * LDS_STRIDE; i += 256) { As[i] = 1.0f + 1e-6f * i; Bs[i] = 1.0f - 1e-6f * i; }
    __syncthreads();
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // 16 x 16 threads, 8 x 8 each
    float sum = 0.f;
    if constexpr (MODE == 0) {
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
        for (int it = 0; it < iters; ++it) {
#pragma unroll KUNROLL
            for (int k4 = 0; k4 < BK / 4; ++k4) {
                float4 a[8], b[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const float4 *>(&As[(i * 16 + ty) * LDS_STRIDE + k4 * 4]);
#pragma unroll
                for (int j = 0; j < 8; ++j) b[j] = *reinterpret_cast<const float4 *>(&Bs[(j * 16 + tx) * LDS_STRIDE + k4 * 4]);
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]);
                        acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
                        acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]);
                        acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
                    }
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) sum += acc[i][j];
    } else if constexpr (MODE == 1) {
        unsigned long long acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0ull;
        for (int it = 0; it < iters; ++it) {
#pragma unroll KUNROLL
            for (int k4 = 0; k4 < BK / 4; ++k4) {
                ulonglong2 a[8], b[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const ulonglong2 *>(&As[(i * 16 + ty) * LDS_STRIDE + k4 * 4]);
#pragma unroll
                for (int j = 0; j < 8; ++j) b[j] = *reinterpret_cast<const ulonglong2 *>(&Bs[(j * 16 + tx) * LDS_STRIDE + k4 * 4]);
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        acc[i][j] = ffma2(a[i].x, b[j].x, acc[i][j]);
                        acc[i][j] = ffma2(a[i].y, b[j].y, acc[i][j]);
                    }
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) { float2 v = unpack(acc[i][j]); sum += v.x + v.y; }
    } else {
        unsigned long long acc[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0ull;
        for (int it = 0; it < iters; ++it) {
#pragma unroll KUNROLL
            for (int k4 = 0; k4 < BK / 4; ++k4) {
                float4 a[8], b[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const float4 *>(&As[(i * 16 + ty) * LDS_STRIDE + k4 * 4]);
#pragma unroll
                for (int j = 0; j < 8; ++j) b[j] = *reinterpret_cast<const float4 *>(&Bs[(j * 16 + tx) * LDS_STRIDE + k4 * 4]);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const unsigned long long ax = pack(a[i].x, a[i].x), ay = pack(a[i].y, a[i].y),
                                             az = pack(a[i].z, a[i].z), aw = pack(a[i].w, a[i].w);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        acc[i][j] = ffma2(ax, pack(b[2 * j].x, b[2 * j + 1].x), acc[i][j]);
                        acc[i][j] = ffma2(ay, pack(b[2 * j].y, b[2 * j + 1].y), acc[i][j]);
                        acc[i][j] = ffma2(az, pack(b[2 * j].z, b[2 * j + 1].z), acc[i][j]);
                        acc[i][j] = ffma2(aw, pack(b[2 * j].w, b[2 * j + 1].w), acc[i][j]);
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) { float2 v = unpack(acc[i][j]); sum += v.x + v.y; }
    }
    out[blockIdx.x * 256 + threadIdx.x] = sum;
}

template <int MODE>
static void run(const char *name, int ctas_per_sm)
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = sms * ctas_per_sm, iters = 4000;
    float *out;
    cudaMalloc(&out, (size_t)grid * 256 * sizeof(float));
    k_tile<MODE><<<grid, 256>>>(out, 10);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        k_tile<MODE><<<grid, 256>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double flop = 2.0 * grid * 128.0 * 128.0 * BK * iters;
    printf("%-34s grid %4d  %8.3f ms  %7.2f TFLOP/s  (%s)\n", name, grid, best, flop / best / 1e9, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main()
{
    run<0>("scalar FFMA 8x8", 1);
    run<0>("scalar FFMA 8x8", 2);
    run<1>("FFMA2 k-pairs (128 acc regs)", 1);
    run<2>("FFMA2 j-pairs, a duplicated", 1);
    run<2>("FFMA2 j-pairs, a duplicated", 2);
    return 0;
}
