// scan_i8.cu — host side of the byte prefilter (scan_i8.cuh): the int8 shadow of a shard's built rows, the
// per-context scratch, and the launcher. Own translation unit so the library builds in parallel.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "index.h"
#include "scan_i8.cuh"

namespace csgpu {

// below this many rows per shard the fp32 scan is < 60 us and the threshold has too few warps to feed on
// (CSGPU_I8_MIN_ROWS overrides it: the tests exercise the kernel on small corpora)
static uint64_t i8_min_rows()
{
    static const uint64_t v = [] {
        const char *e = getenv("CSGPU_I8_MIN_ROWS");
        return e && *e ? (uint64_t)strtoull(e, nullptr, 10) : (uint64_t)524288;
    }();
    return v;
}
constexpr uint32_t I8_MAX_GRID = I8_MAX_WARPS / I8_WARPS;

static inline uint32_t i8_lines(uint32_t dim4) { return (dim4 + 31) / 32; }   // 128-byte lines per shadow row (= V)

void i8_free_shard(Shard *sh)
{
    DeviceGuard g(sh->device);
    cudaFree(sh->shadow_i8); cudaFree(sh->meta_i8);
    sh->shadow_i8 = nullptr; sh->meta_i8 = nullptr; sh->i8_valid = false; sh->i8_rows = 0;
}

void i8_free_ctx(SearchCtx *c)
{
    cudaFree(c->i8_scratch);
    if (c->i8_status) cudaFreeHost(c->i8_status);
    c->i8_scratch = nullptr; c->i8_status = nullptr;
}

// (re)build or drop the int8 shadow of an fp32 shard's built rows (csgpu_set_byte_prefilter)
int i8_refresh(const csgpu_index *ix, Shard *sh)
{
    i8_free_shard(sh);
    if (!ix->byte_prefilter || ix->dtype != CSGPU_DTYPE_F32 || sh->n_built == 0 || sh->rows == nullptr) return CSGPU_OK;
    const uint32_t V = i8_lines(ix->dim4);
    if (V > 8) return fail(CSGPU_ERR_ARG, "byte prefilter needs dim <= 1024");
    DeviceGuard g(sh->device);
    const size_t d8 = (size_t)V * 128;
    cudaError_t e = cudaMalloc(&sh->shadow_i8, sh->n_built * d8);
    if (e == cudaSuccess) e = cudaMalloc(&sh->meta_i8, sh->n_built * sizeof(uint32_t));
    if (e != cudaSuccess) {
        cudaGetLastError();
        i8_free_shard(sh);
        return fail(CSGPU_ERR_OOM, "no room in HBM for the int8 shadow of the fp32 rows (byte prefilter needs +25 %)");
    }
    const uint32_t grid = (uint32_t)std::min<uint64_t>((sh->n_built + 7) / 8, (uint64_t)sh->sm_count * 8);
    shadow_i8_from_rows_kernel<<<grid, 256, 0, sh->stream>>>(reinterpret_cast<const float4 *>(sh->rows),
                                                            reinterpret_cast<uint32_t *>(sh->shadow_i8), sh->meta_i8, 0, sh->n_built,
                                                            ix->dim4, (uint32_t)d8);
    count_launch();
    CS_CUDA(cudaGetLastError());
    CS_CUDA(cudaStreamSynchronize(sh->stream));
    sh->i8_rows = sh->n_built;
    sh->i8_valid = true;
    return CSGPU_OK;
}

bool i8_eligible(const csgpu_index *ix, uint32_t k)
{
    if (!ix->byte_prefilter || ix->dtype != CSGPU_DTYPE_F32 || k == 0 || k > I8_MAX_K) return false;
    for (const Shard *sh : ix->shards)
        if (!sh->i8_valid || sh->i8_rows != sh->n_built || sh->n_built < i8_min_rows()) return false;
    return true;
}

struct I8Scratch {   // one cudaMalloc block per SearchCtx
    unsigned long long timing[(I8_MAX_GRID + 1) * 4];
    uint32_t warp_min[I8_MAX_WARPS];
    unsigned counters[8];
    uint64_t final_list[I8_FINAL_CAP];
    uint64_t region[(size_t)I8_MAX_GRID * I8_REGION];
};

// rows per 4-row group count of the FILT instantiation: 4 R must divide a 32-row block (R = 6 at dim 384 does not: R = 4)
template <int R>
constexpr int i8_filt_r() { return R == 6 ? 4 : R; }

template <int V, bool EXACT, int R>
static cudaError_t launch_i8_vr(const I8Args &a, uint32_t grid, cudaStream_t st, bool filtered)
{
    auto kern = scan_i8_kernel<V, EXACT, R, false>;
    auto kern_f = scan_i8_kernel<V, EXACT, i8_filt_r<R>(), true>;
    if (grid == 0) {   // preload only
        cudaFuncAttributes fa;
        cudaError_t e = cudaFuncGetAttributes(&fa, kern);
        return e == cudaSuccess ? cudaFuncGetAttributes(&fa, kern_f) : e;
    }
    const size_t smem = (size_t)I8_TAIL_CAP * sizeof(uint64_t);
    if (filtered) kern_f<<<grid, I8_THREADS, smem, st>>>(a);
    else kern<<<grid, I8_THREADS, smem, st>>>(a);
    count_launch();
    return cudaGetLastError();
}

// R per V. V = 3 (dim 384): 24 rows = 9 KB per warp iteration; measured on one box, A/B/A/B: R = 4 / 5 / 6 ->
// 558 / 553-558 / 549-553 us per 10M-row query (more bytes in flight across the longer compute phase of this kernel).
template <int V>
constexpr int i8_default_r() { return (V <= 2) ? 8 : (V == 3 ? 6 : (V == 4 ? 4 : 2)); }

static uint32_t i8_rows_per_iter(uint32_t V)
{
    return 4u * ((V <= 2) ? 8 : (V == 3 ? 6 : (V == 4 ? 4 : 2)));
}

template <int V, bool EXACT>
static cudaError_t launch_i8_v(const I8Args &a, uint32_t grid, cudaStream_t st, bool filtered)
{
    return launch_i8_vr<V, EXACT, i8_default_r<V>()>(a, grid, st, filtered);
}

const unsigned *i8_status_dev(const SearchCtx *c) { return reinterpret_cast<const I8Scratch *>(c->i8_scratch)->counters + 4; }

// force-load this index's instantiation (see preload_exchange_kernels in csgpu.cu)
void i8_preload(const csgpu_index *ix)
{
    const uint32_t V = i8_lines(ix->dim4);
    const bool exact = (ix->dim4 % 32) == 0;
    I8Args a{};
#define CS_CASE(v) case v: if (exact) launch_i8_v<v, true>(a, 0, nullptr, false); else launch_i8_v<v, false>(a, 0, nullptr, false); break;
    switch (V) {
        CS_CASE(1) CS_CASE(2) CS_CASE(3) CS_CASE(4) CS_CASE(5) CS_CASE(6) CS_CASE(7) CS_CASE(8)
        default: break;
    }
#undef CS_CASE
    cudaGetLastError();
}

// The context's scratch: both blocks or neither. Called when a context is created on an index with the prefilter on,
// and for the warm contexts of csgpu_set_byte_prefilter / csgpu_build — not from inside a search (round 1 paid a
// multi-millisecond cudaMalloc in the first query).
int i8_prepare_ctx(SearchCtx *c)
{
    if (c->i8_scratch != nullptr) return CSGPU_OK;
    void *scratch = nullptr;
    uint64_t *status = nullptr;
    cudaError_t ea = cudaMalloc(&scratch, sizeof(I8Scratch));
    if (ea == cudaSuccess) ea = cudaHostAlloc(&status, 8 * sizeof(uint64_t), cudaHostAllocMapped | cudaHostAllocPortable);
    if (ea == cudaSuccess) ea = cudaMemset(reinterpret_cast<I8Scratch *>(scratch)->warp_min, 0xFF, sizeof(uint32_t) * I8_MAX_WARPS);
    if (ea == cudaSuccess) ea = cudaMemset(reinterpret_cast<I8Scratch *>(scratch)->counters, 0, sizeof(unsigned) * 8);
    if (ea != cudaSuccess) {
        cudaFree(scratch);
        if (status) cudaFreeHost(status);
        return fail_cuda(ea, "byte prefilter scratch", __FILE__, __LINE__);
    }
    memset(status, 0, 8 * sizeof(uint64_t));
    c->i8_scratch = scratch;
    c->i8_status = status;
    return CSGPU_OK;
}

int enqueue_scan_i8(const csgpu_index *ix, const Shard *sh, SearchCtx *c, const float *q_dev, uint32_t k,
                    bool with_zero_ids, uint64_t *out_keys, cudaStream_t st, bool host_status,
                    const uint64_t *bitmap_dev, uint64_t n_bits, const csgpu_predicate_t *pred)
{
    if (int rc = i8_prepare_ctx(c)) return rc;   // contexts that predate csgpu_set_byte_prefilter
    I8Scratch *s = reinterpret_cast<I8Scratch *>(c->i8_scratch);
    const uint32_t V = i8_lines(ix->dim4);
    I8Args a;
    a.shadow = sh->shadow_i8;
    a.meta = sh->meta_i8;
    a.rows = reinterpret_cast<const float4 *>(sh->rows);
    a.ids = sh->ids;
    a.n_rows = sh->n_built;
    a.dim4 = ix->dim4;
    a.d8 = V * 128;
    a.q = q_dev;
    a.k = k;
    a.zero_ids = with_zero_ids ? ix->zero_ids_dev : nullptr;
    a.n_zero = with_zero_ids ? (uint32_t)ix->zero_ids.size() : 0;
    a.warp_min = s->warp_min;
    a.region = s->region;
    a.final_list = s->final_list;
    a.counters = s->counters;
    a.out_keys = out_keys;
    a.status = c->i8_status;
    const bool filtered = bitmap_dev != nullptr || pred != nullptr;
    if (filtered) {   // same conventions as enqueue_scan: with pred, bitmap_dev is its per-FILE bitmap (or nullptr)
        a.bitmap = bitmap_dev;
        a.n_bits = n_bits;
        if (pred) { a.tags = sh->tags; a.lang_mask = pred->lang_mask; a.file_lo = pred->file_lo; a.file_hi = pred->file_hi; }
    }
    static const bool timing = getenv("CSGPU_I8_TIMING") != nullptr;
    a.timing = timing ? s->timing : nullptr;
    if (host_status) c->i8_status[0] = 1;   // a launch that never runs must not look like a success
    const uint64_t per_cta = (uint64_t)I8_WARPS * i8_rows_per_iter(V);
    const uint64_t want = (sh->n_built + per_cta - 1) / per_cta;
    const uint32_t grid = (uint32_t)std::min<uint64_t>(std::min<uint64_t>((uint64_t)sh->sm_count * 2, I8_MAX_GRID), std::max<uint64_t>(want, 1));
    const bool exact = (ix->dim4 % 32) == 0;
    cudaError_t e;
#define CS_CASE(v) case v: e = exact ? launch_i8_v<v, true>(a, grid, st, filtered) : launch_i8_v<v, false>(a, grid, st, filtered); break;
    switch (V) {
        CS_CASE(1) CS_CASE(2) CS_CASE(3) CS_CASE(4) CS_CASE(5) CS_CASE(6) CS_CASE(7) CS_CASE(8)
        default: e = cudaErrorInvalidValue;
    }
#undef CS_CASE
    if (e != cudaSuccess) return fail_cuda(e, "scan_i8_kernel launch", __FILE__, __LINE__);
    if (timing && host_status) {   // diagnostic: per-CTA globaltimer stamps -> where a query's time goes (synchronises!)
        std::vector<unsigned long long> t((grid + 1) * 4);
        CS_CUDA(cudaStreamSynchronize(st));
        CS_CUDA(cudaMemcpy(t.data(), s->timing, t.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        unsigned long long t0 = ~0ull, s_lo = ~0ull, s_hi = 0, e_hi = 0, st_hi = 0;
        for (uint32_t b = 0; b < grid; ++b) {
            t0 = std::min(t0, t[b * 4]); st_hi = std::max(st_hi, t[b * 4]);
            s_lo = std::min(s_lo, t[b * 4 + 1]); s_hi = std::max(s_hi, t[b * 4 + 1]); e_hi = std::max(e_hi, t[b * 4 + 2]);
        }
        unsigned n_inf = 0; unsigned long long surv = 0, surv_max = 0;
        for (uint32_t b = 0; b < grid; ++b) { n_inf += (uint32_t)t[b * 4 + 3] == 0xFFFFFFFFu; surv += t[b * 4 + 3] >> 32; surv_max = std::max(surv_max, t[b * 4 + 3] >> 32); }
        fprintf(stderr, "[i8 timing] CTAs whose G was still +inf at their end: %u of %u; survivors rescored: %llu (max %llu in one CTA)\n", n_inf, grid, surv, surv_max);
        fprintf(stderr, "[i8 timing] grid %u: last CTA start +%.1f us | streaming ends first +%.1f last +%.1f | last ticket +%.1f | tail done +%.1f us\n",
                grid, (st_hi - t0) / 1e3, (s_lo - t0) / 1e3, (s_hi - t0) / 1e3, (e_hi - t0) / 1e3, (t[grid * 4] - t0) / 1e3);
    }
    return CSGPU_OK;
}

}  // namespace csgpu
