// synth.cuh — counter-based synthetic corpus generator + build-time kernels (normalise, compact).
//
// Generator spec (bit-identical to oracle/oracle.c cs_synth_rows, pinned by the Random123
// Philox4x32-10 known-answer vectors in tests/test_oracle.py):
//   (x0..x3) = Philox4x32-10(ctr = {row_lo, row_hi, col/4, 0}, key = {seed_lo, seed_hi})
//   value[col+j] = float(byte0(xj)+byte1(xj)+byte2(xj)+byte3(xj) - 510)      (Irwin-Hall(4), exact)
// Raw rows are not unit length; csgpu_build normalises (the reference's embeddings are only
// approximately unit too: /root/reference/src/embed/embedder.rs:461-463).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace csgpu {

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

__device__ __forceinline__ float synth_val(uint32_t x)
{
    // sum of the four bytes: __vsadu4(x, 0) = |b0-0|+|b1-0|+|b2-0|+|b3-0|
    return (float)((int)__vsadu4(x, 0u) - 510);
}

// One thread per float4. rows: [n, dim4] float4, row-major.
static __global__ void synth_rows_kernel(float4 *__restrict__ rows, uint32_t *__restrict__ ids, uint64_t seed,
                                  uint64_t first_row, uint64_t n, uint32_t dim4, uint32_t id_base)
{
    const uint64_t total = n * dim4;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = t / dim4;
        const uint32_t c4 = (uint32_t)(t - i * dim4);
        const uint64_t row = first_row + i;
        const uint4 x = philox4x32_10(make_uint4((uint32_t)row, (uint32_t)(row >> 32), c4, 0u),
                                      make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
        rows[t] = make_float4(synth_val(x.x), synth_val(x.y), synth_val(x.z), synth_val(x.w));
        if (ids != nullptr && c4 == 0) ids[i] = (uint32_t)row + id_base;
    }
}

// Synthetic row tags (bit-identical to oracle/oracle.py synth_tags): file_id = row / 37,
// lang_id = fmix32(file_id) % 23, tag = (lang_id << 27) | file_id.
__host__ __device__ __forceinline__ uint32_t fmix32(uint32_t h)
{
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}
static __global__ void synth_tags_kernel(uint32_t *__restrict__ tags, uint64_t first_row, uint64_t n)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t file = (uint32_t)((first_row + i) / 37) & 0x07FFFFFFu;
        tags[i] = ((fmix32(file) % 23u) << 27) | file;
    }
}

// Row status written by normalise_rows_kernel.
enum : uint8_t { ROW_OK = 0, ROW_DEAD = 1, ROW_ZERO = 2, ROW_NONFINITE = 3 };

// One warp per row: unit-normalise rows [first, first+n) in place. The norm is accumulated in
// f64 (build is a one-off pass; B200 has the f64 rate to spare) so stored unit rows are the
// correctly-rounded v/|v| to within 1 ulp. Zero-norm and non-finite rows are flagged, not scaled.
static __global__ void normalise_rows_kernel(float4 *__restrict__ rows, uint8_t *__restrict__ status,
                                      uint64_t first, uint64_t n, uint32_t dim4)
{
    const int lane = threadIdx.x & 31;
    const uint64_t w0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t i = w0; i < n; i += nw) {
        const uint64_t row = first + i;
        if (status[row] == ROW_DEAD) continue;
        float4 *p = rows + row * dim4;
        double ss = 0.0;
        bool finite = true;
        for (uint32_t c = lane; c < dim4; c += 32) {
            const float4 v = p[c];
            ss += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
            finite = finite && isfinite(v.x) && isfinite(v.y) && isfinite(v.z) && isfinite(v.w);
        }
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) ss += __shfl_xor_sync(0xFFFFFFFFu, ss, m);
        finite = __all_sync(0xFFFFFFFFu, finite);
        if (!finite) { if (lane == 0) status[row] = ROW_NONFINITE; continue; }
        if (!(ss > 0.0)) { if (lane == 0) status[row] = ROW_ZERO; continue; }
        const double inv = 1.0 / sqrt(ss);
        for (uint32_t c = lane; c < dim4; c += 32) {
            float4 v = p[c];
            v.x = (float)(v.x * inv); v.y = (float)(v.y * inv);
            v.z = (float)(v.z * inv); v.w = (float)(v.w * inv);
            p[c] = v;
        }
    }
}

// Marks rows [0, n) whose id is in the sorted list `kill` (ascending, n_kill entries) as ROW_DEAD.
// counts[0] += rows newly killed.
static __global__ void mark_dead_kernel(const uint32_t *__restrict__ ids, uint8_t *__restrict__ status, uint64_t n,
                                 const uint32_t *__restrict__ kill, uint32_t n_kill,
                                 unsigned long long *__restrict__ counts)
{
    unsigned long long local = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x) {
        if (status[i] == ROW_DEAD) continue;
        const uint32_t id = ids[i];
        uint32_t lo = 0, hi = n_kill;
        while (lo < hi) {
            uint32_t mid = (lo + hi) >> 1;
            if (kill[mid] < id) lo = mid + 1; else hi = mid;
        }
        if (lo < n_kill && kill[lo] == id) { status[i] = ROW_DEAD; ++local; }
    }
    if (local) atomicAdd(counts, local);
}

// Stable compaction, one chunk at a time through a bounce buffer (dst <= src always, so moving
// chunk c never overwrites rows of chunks > c). dst_index[i] = destination row of row (first+i)
// or ~0ull if dropped; one warp per row.
static __global__ void gather_rows_kernel(const float4 *__restrict__ src, float4 *__restrict__ dst,
                                   const uint64_t *__restrict__ dst_index, uint64_t n, uint32_t dim4,
                                   uint64_t src_first, uint64_t dst_base_sub)
{
    const int lane = threadIdx.x & 31;
    const uint64_t w0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t i = w0; i < n; i += nw) {
        const uint64_t d = dst_index[i];
        if (d == ~0ull) continue;
        const float4 *s = src + (src_first + i) * dim4;
        float4 *o = dst + (d - dst_base_sub) * dim4;
        for (uint32_t c = lane; c < dim4; c += 32) o[c] = s[c];
    }
}

}  // namespace csgpu
