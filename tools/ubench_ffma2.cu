// ubench_ffma2.cu — round 2: WHERE does the fp32 SIMT outer-product loop lose its FMA slots on B200?
// Round 1 (tools/ubench_ffma.cu) found 51-54 TFLOP/s for the 8x8-per-thread loop with operands from shared memory, at one or
// two CTAs per SM, scalar or packed — i.e. ~72 % of 148 SM x 128 lanes x 2 x 1.965 GHz = 74.4 TFLOP/s. This file separates
// the candidates: (P) the FMA pipe alone, registers only; (S) the same loop with varying LDS.128 : FFMA ratios (per-thread
// tile 8x8, 8x16, 16x8, 4 warps x 8 ...), lane layouts (how many lanes of a warp share an A / B address) and thread counts.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/ubench_ffma2.cu -o build/ubench_ffma2
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int BK = 32;
constexpr int STRIDE = 36;   // floats; 144 B pitch -> 8 consecutive rows hit 8 different 16-byte bank groups

// ---- P: registers only -----------------------------------------------------------------------------------------------
template <int ACC>
__global__ void __launch_bounds__(256, 1) k_pure(float *out, int iters, float a0, float b0)
{
    float acc[ACC];
    float a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = a0 + i + threadIdx.x; b[i] = b0 - i; }
#pragma unroll
    for (int i = 0; i < ACC; ++i) acc[i] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < ACC; ++i) acc[i] = fmaf(a[(i / 8 + r) & 7], b[(i + r) & 7], acc[i]);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- S: operands from shared memory ----------------------------------------------------------------------------------
// per-thread tile TM (queries) x TN (rows); a warp is WQ x (32/WQ) threads; THREADS per CTA; the CTA tile follows.
template <int TM, int TN, int WQ, int THREADS, int WARPS_Q>
__global__ void __launch_bounds__(THREADS, 1) k_smem(float *out, int iters)
{
    constexpr int WR = 32 / WQ, NW = THREADS / 32, WARPS_R = NW / WARPS_Q;
    constexpr int TYD = WARPS_Q * WQ, TXD = WARPS_R * WR, BM = TM * TYD, BN = TN * TXD;
    extern __shared__ __align__(16) float sm[];
    float *As = sm, *Bs = sm + BM * STRIDE;
    for (int i = threadIdx.x; i < (BM + BN) * STRIDE; i += THREADS) sm[i] = 1.0f + 1e-6f * i;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int ty = (w / WARPS_R) * WQ + lane / WR, tx = (w % WARPS_R) * WR + lane % WR;
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 2
        for (int k4 = 0; k4 < BK / 4; ++k4) {
            float4 a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = *reinterpret_cast<const float4 *>(&As[(i * TYD + ty) * STRIDE + k4 * 4]);
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = *reinterpret_cast<const float4 *>(&Bs[(j * TXD + tx) * STRIDE + k4 * 4]);
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) {
                    acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]);
                    acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
                    acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]);
                    acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
                }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) s += acc[i][j];
    out[blockIdx.x * THREADS + threadIdx.x] = s;
}

// Same, but the B operand is streamed one float4 at a time (fewer live registers: lets 512 threads fit 128 regs)
template <int TM, int TN, int WQ, int THREADS, int WARPS_Q>
__global__ void __launch_bounds__(THREADS, 1) k_smem_stream(float *out, int iters)
{
    constexpr int WR = 32 / WQ, NW = THREADS / 32, WARPS_R = NW / WARPS_Q;
    constexpr int TYD = WARPS_Q * WQ, TXD = WARPS_R * WR, BM = TM * TYD, BN = TN * TXD;
    extern __shared__ __align__(16) float sm[];
    float *As = sm, *Bs = sm + BM * STRIDE;
    for (int i = threadIdx.x; i < (BM + BN) * STRIDE; i += THREADS) sm[i] = 1.0f + 1e-6f * i;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int ty = (w / WARPS_R) * WQ + lane / WR, tx = (w % WARPS_R) * WR + lane % WR;
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 1
        for (int k4 = 0; k4 < BK / 4; ++k4) {
            float4 a[TM];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = *reinterpret_cast<const float4 *>(&As[(i * TYD + ty) * STRIDE + k4 * 4]);
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                const float4 b = *reinterpret_cast<const float4 *>(&Bs[(j * TXD + tx) * STRIDE + k4 * 4]);
#pragma unroll
                for (int i = 0; i < TM; ++i) {
                    acc[i][j] = fmaf(a[i].x, b.x, acc[i][j]);
                    acc[i][j] = fmaf(a[i].y, b.y, acc[i][j]);
                    acc[i][j] = fmaf(a[i].z, b.z, acc[i][j]);
                    acc[i][j] = fmaf(a[i].w, b.w, acc[i][j]);
                }
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) s += acc[i][j];
    out[blockIdx.x * THREADS + threadIdx.x] = s;
}

static int g_sms = 148;

template <class F>
static float time_best(F launch)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(10);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        launch(4000);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    return best;
}

template <int ACC>
static void run_pure(int ctas)
{
    float *out;
    cudaMalloc(&out, (size_t)g_sms * ctas * 256 * sizeof(float));
    const float ms = time_best([&](int it) { k_pure<ACC><<<g_sms * ctas, 256>>>(out, it, 1.f, 2.f); });
    const double flop = 2.0 * g_sms * ctas * 256.0 * ACC * 4 * 4000;
    printf("P  registers only, %3d acc, %d CTA/SM x 256 thr          %8.3f ms  %7.2f TFLOP/s  (%s)\n", ACC, ctas, ms, flop / ms / 1e9,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

template <int TM, int TN, int WQ, int THREADS, int WARPS_Q, bool STREAM>
static void run_smem(int ctas)
{
    constexpr int WR = 32 / WQ, NW = THREADS / 32, WARPS_R = NW / WARPS_Q;
    constexpr int BM = TM * WARPS_Q * WQ, BN = TN * WARPS_R * WR;
    const size_t smem = (size_t)(BM + BN) * STRIDE * sizeof(float);
    float *out;
    cudaMalloc(&out, (size_t)g_sms * ctas * THREADS * sizeof(float));
    auto kern = STREAM ? k_smem_stream<TM, TN, WQ, THREADS, WARPS_Q> : k_smem<TM, TN, WQ, THREADS, WARPS_Q>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, kern);
    const float ms = time_best([&](int it) { kern<<<g_sms * ctas, THREADS, smem>>>(out, it); });
    const double flop = 2.0 * g_sms * ctas * (double)BM * BN * BK * 4000;
    printf("S%s tile %2dx%-2d warp %2dx%-2d thr %3d CTA %3dx%-3d x%d  regs %3d spill %4zu  LDS128/FFMA %.4f  %8.3f ms  %7.2f TFLOP/s  (%s)\n",
           STREAM ? "s" : " ", TM, TN, WQ, WR, THREADS, BM, BN, ctas, fa.numRegs, (size_t)fa.localSizeBytes, (double)(TM + TN) / (TM * TN * 4.0), ms,
           flop / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main()
{
    cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, 0);
    run_pure<64>(1);
    run_pure<64>(2);
    run_pure<128>(1);
    //        TM  TN  WQ  THR  WARPS_Q
    run_smem<8, 8, 4, 256, 4, false>(1);    // round-1 shape
    run_smem<8, 8, 8, 256, 2, false>(1);    // 8 x 4 lanes
    run_smem<8, 8, 2, 256, 8, false>(1);    // 2 x 16 lanes
    run_smem<8, 8, 1, 256, 8, false>(1);    // 1 x 32 lanes (A pure broadcast)
    run_smem<8, 16, 4, 256, 4, false>(1);   // 128 accumulators
    run_smem<16, 8, 4, 256, 4, false>(1);
    run_smem<8, 16, 4, 256, 4, true>(1);
    run_smem<16, 8, 4, 256, 4, true>(1);
    run_smem<8, 12, 4, 256, 4, false>(1);
    run_smem<12, 8, 4, 256, 4, false>(1);
    run_smem<8, 8, 4, 512, 4, true>(1);     // 16 warps, <= 128 regs
    run_smem<8, 8, 4, 512, 4, false>(1);
    run_smem<4, 16, 4, 256, 4, false>(1);
    run_smem<16, 4, 4, 256, 4, false>(1);
    return 0;
}
