// gemm_topk.cu — the batched search path (BASELINE config C3): TMA tensor maps, the phase loop around the
// GEMM-shaped kernels and the per-query exact select kernel. Three contraction kernels share it:
//   fp32 index (default) : gemm_tf32_topk_kernel (gemm_tf32.cuh, tcgen05 kind::tf32 straight off the fp32 rows, as a FILTER;
//                          survivors rescored exactly by select_sorted_kernel<.., RESCORE>, rescore.cuh)
//   fp32 index, dim > 1024 or CSGPU_BATCH_SIMT=1 : gemm_simt_topk_kernel (gemm_simt.cuh, register-tiled FP32 SIMT)
//   bf16 index, or the opt-in bf16 shadow of an fp32 index : gemm_topk_kernel (gemm_topk.cuh, tcgen05 kind::f16 / TMEM)
// plus the bf16 index's storage hooks. See gemm_topk.cuh for the progressive-threshold design.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <thread>

#include "index.h"
#include "gemm_simt.cuh"
#include "gemm_topk.cuh"
#include "gemm_tf32.cuh"
#include "rescore.cuh"
#include "scan.cuh"
#include "synth.cuh"

namespace csgpu {

constexpr uint32_t BF_CAP = 8192;          // candidate slots per query
constexpr uint32_t BF_MAX_QBLOCKS = 8;     // 1024 queries per pass
constexpr uint32_t BF_PHASE_GROWTH = 8;

// Which contraction answers a batch on this shard:
//   TC_BF16  tcgen05 kind::f16 over bf16 rows — the bf16 index itself, or the bf16 shadow of an fp32 index (opt-in tensor
//            prefilter, rescore.cuh);
//   TC_TF32  tcgen05 kind::tf32 straight off the fp32 rows (gemm_tf32.cuh) — the DEFAULT for an fp32 index (round 2): a
//            filter, rescored exactly, so results stay bit-identical to csgpu_search;
//   SIMT     the register-tiled FP32 kernel (gemm_simt.cuh): dim > 1024, or CSGPU_BATCH_SIMT=1 (A/B runs, tests).
enum class Contraction { SIMT, TC_BF16, TC_TF32 };
static bool simt_forced()
{
    const char *e = getenv("CSGPU_BATCH_SIMT");   // read per batch (not cached): tests flip it inside one process
    return e != nullptr && e[0] == '1';
}
static Contraction contraction_of(const csgpu_index *ix, const Shard *sh)
{
    if (ix->dtype == CSGPU_DTYPE_BF16 || (ix->tensor_prefilter && sh->shadow_valid)) return Contraction::TC_BF16;
    if (ix->dim_pad <= TF_MAX_DIM && !simt_forced()) return Contraction::TC_TF32;
    return Contraction::SIMT;
}
static bool tc_path(const csgpu_index *ix, const Shard *sh) { return contraction_of(ix, sh) != Contraction::SIMT; }
static bool rescore_path(const csgpu_index *ix, const Shard *sh) { return ix->dtype == CSGPU_DTYPE_F32 && tc_path(ix, sh); }
// query rows the tf32 kernel loads per K chunk: all 128 of a block, or only the (8-row groups of) queries there are
static uint32_t tf32_q_rows(uint32_t nq) { return nq >= (uint32_t)GT_BLOCK_M ? (uint32_t)GT_BLOCK_M : std::max<uint32_t>(8, (nq + 7) & ~7u); }

// ---------------------------------------------------------------------------------------------
// kernels local to this file
// ---------------------------------------------------------------------------------------------
__global__ void init_thresholds_kernel(float *thr, unsigned *count, const uint8_t *flags, uint32_t n_active, uint32_t n_total)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_total) { thr[i] = (i < n_active && !flags[i]) ? __int_as_float(0x7f800000) : -1.f; count[i] = 0; }
}

// One thread per query: would this range's candidates overflow a segment, the SIMT buffer or the select kernel's
// sort buffer? Sets *overflow (sticky). The single source of truth for the host's split-and-retry.
__global__ void check_counts_kernel(const unsigned *__restrict__ count, const unsigned *__restrict__ count_saved,
                                    const unsigned *__restrict__ seg_count, uint32_t n_seg, uint32_t seg_len, uint32_t cap,
                                    uint32_t nq, unsigned *__restrict__ overflow)
{
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    bool over = false;
    if (n_seg) {
        unsigned total = 0;
        for (uint32_t s = 0; s < n_seg; ++s) {
            const unsigned c = seg_count[(size_t)q * n_seg + s];
            over |= c > seg_len;
            total += c;
        }
        over |= total > SEL_SORT_CAP;
    } else {
        over = count[q] > cap || count[q] - min(count[q], count_saved[q]) > SEL_SORT_CAP;
    }
    if (over) atomicExch(overflow, 1u);
}

// Zero-norm query on the bf16 index (contract, include/csgpu.h: "a zero-norm query or row has distance 0.0"; arroy's
// pn*qn == 0 branch): every live row ties at distance 0.0, so the answer is the k smallest chunk ids — of this shard's
// rows plus (shard 0) the zero-norm side list. One pass over ids[] (4 B per row); per-CTA selection in a CtaBuf, the last
// CTA re-selects over the CTAs' lists. The fp32 index gets the same answer from scan_topk_kernel's qzero branch.
__global__ void __launch_bounds__(SCAN_THREADS, 1)
zero_query_ids_kernel(const uint32_t *__restrict__ ids, uint64_t n, const uint32_t *__restrict__ zero_ids, uint32_t n_zero,
                      uint32_t k, uint32_t cap, uint64_t *__restrict__ cand, unsigned *__restrict__ ticket, uint64_t *__restrict__ out)
{
    extern __shared__ __align__(16) uint64_t smem[];
    __shared__ unsigned cb_cnt;
    __shared__ uint64_t cb_thr;
    __shared__ bool is_last;
    CtaSel sel;
    sel.cb.buf = smem; sel.cb.cnt = &cb_cnt; sel.cb.thr = &cb_thr; sel.cap = cap; sel.k = k;
    sel.reset();
    const uint64_t per = (n + gridDim.x - 1) / gridDim.x;
    const uint64_t lo = min(n, per * blockIdx.x), hi = min(n, lo + per);
    cta_buf_stream(sel.cb, cap, k, hi - lo, [&](uint64_t t) { return make_key(0.f, ids[lo + t]); });
    sel.finish();
    for (uint32_t j = threadIdx.x; j < k; j += blockDim.x) cand[(size_t)blockIdx.x * k + j] = smem[j];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    sel.reset();
    const volatile uint64_t *cv = cand;
    cta_buf_stream(sel.cb, cap, k, (uint64_t)gridDim.x * k, [&](uint64_t t) { return cv[t]; });
    cta_buf_stream(sel.cb, cap, k, (uint64_t)min(n_zero, k), [&](uint64_t t) { return make_key(0.f, zero_ids[t]); });   // ascending: the first k suffice
    sel.finish();
    for (uint32_t j = threadIdx.x; j < k; j += blockDim.x) out[j] = smem[j];
    if (threadIdx.x == 0) *ticket = 0;
}

// pending fp32 rows (stage) -> unit length (f64 norm) -> bf16 rows at [dst_first + i]; flags zero / non-finite
__global__ void normalise_to_bf16_kernel(const float *__restrict__ stage, __nv_bfloat16 *__restrict__ rows,
                                         uint8_t *__restrict__ status, uint64_t dst_first, uint64_t n, uint32_t dim)
{
    const int lane = threadIdx.x & 31;
    const uint64_t w0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t i = w0; i < n; i += nw) {
        const uint64_t row = dst_first + i;
        const float *p = stage + i * dim;
        __nv_bfloat16 *o = rows + row * dim;
        double ss = 0.0;
        bool finite = true;
        for (uint32_t c = lane; c < dim; c += 32) { const float x = p[c]; ss += (double)x * x; finite = finite && isfinite(x); }
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) ss += __shfl_xor_sync(0xFFFFFFFFu, ss, m);
        finite = __all_sync(0xFFFFFFFFu, finite);
        const bool dead = status[row] == ROW_DEAD;
        const bool zero = !(ss > 0.0);
        const double inv = (finite && !zero) ? 1.0 / sqrt(ss) : 0.0;
        for (uint32_t c = lane; c < dim; c += 32) o[c] = __float2bfloat16((float)(p[c] * inv));
        if (lane == 0 && !dead) {
            if (!finite) status[row] = ROW_NONFINITE;
            else if (zero) status[row] = ROW_ZERO;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// tensor maps
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// [rows][cols] row-major matrix of 2-byte (bf16) or 4-byte (fp32) elements with a row pitch of `cols` elements;
// box = [box_rows][128 bytes], 128-byte swizzle; out-of-bounds elements read as zero.
static int make_map(CUtensorMap *map, const void *base, uint64_t rows, uint32_t cols, uint32_t box_rows, bool f32)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(CSGPU_ERR_CUDA, "cuTensorMapEncodeTiled is unavailable in this driver");
    const uint32_t esz = f32 ? 4 : 2;
    const cuuint64_t gdim[2] = {cols, std::max<uint64_t>(rows, 1)};
    const cuuint64_t gstride[1] = {(cuuint64_t)cols * esz};
    const cuuint32_t box[2] = {128 / esz, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    static const int env_promo = getenv("CSGPU_TMA_L2_PROMO") ? atoi(getenv("CSGPU_TMA_L2_PROMO")) : 3;   // experiments: 0 none, 1 64 B, 2 128 B, 3 256 B
    const CUtensorMapL2promotion promo = env_promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : env_promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                         : env_promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    CUresult r = fn(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base),
                    gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(CSGPU_ERR_CUDA, "cuTensorMapEncodeTiled failed (code " + std::to_string((int)r) + ")");
    return CSGPU_OK;
}

// ---------------------------------------------------------------------------------------------
// storage hooks (called from csgpu.cu)
// ---------------------------------------------------------------------------------------------
bool bf16_dim_supported(uint32_t dim) { return dim % 64 == 0 && dim >= 64 && dim <= 512; }

int bf16_reserve_rows(const csgpu_index *ix, Shard *sh, uint64_t rows)
{
    if (rows <= sh->cap) return CSGPU_OK;
    DeviceGuard g(sh->device);
    uint64_t ncap = std::max<uint64_t>(std::max<uint64_t>(rows, sh->cap + sh->cap / 2), 1024);
    void *nrows = nullptr; uint32_t *nids = nullptr, *ntags = nullptr; uint8_t *nst = nullptr;
    const size_t row_bytes = (size_t)ix->dim * 2;
    cudaError_t e = cudaMalloc(&nrows, ncap * row_bytes);
    if (e != cudaSuccess) return fail_cuda(e, "cudaMalloc(bf16 rows)", __FILE__, __LINE__);
    if ((e = cudaMalloc(&nids, ncap * sizeof(uint32_t))) != cudaSuccess) { cudaFree(nrows); return fail_cuda(e, "cudaMalloc(ids)", __FILE__, __LINE__); }
    if ((e = cudaMalloc(&nst, ncap)) != cudaSuccess) { cudaFree(nrows); cudaFree(nids); return fail_cuda(e, "cudaMalloc(status)", __FILE__, __LINE__); }
    if ((e = cudaMalloc(&ntags, ncap * sizeof(uint32_t))) != cudaSuccess) { cudaFree(nrows); cudaFree(nids); cudaFree(nst); return fail_cuda(e, "cudaMalloc(tags)", __FILE__, __LINE__); }
    CS_CUDA(cudaMemsetAsync(nst, 0, ncap, sh->stream));
    CS_CUDA(cudaMemsetAsync(ntags, 0xFF, ncap * sizeof(uint32_t), sh->stream));
    if (sh->n_built) CS_CUDA(cudaMemcpyAsync(nrows, sh->rows_bf16, sh->n_built * row_bytes, cudaMemcpyDeviceToDevice, sh->stream));
    if (sh->n_total) {
        CS_CUDA(cudaMemcpyAsync(nids, sh->ids, sh->n_total * sizeof(uint32_t), cudaMemcpyDeviceToDevice, sh->stream));
        CS_CUDA(cudaMemcpyAsync(nst, sh->status, sh->n_total, cudaMemcpyDeviceToDevice, sh->stream));
        CS_CUDA(cudaMemcpyAsync(ntags, sh->tags, sh->n_total * sizeof(uint32_t), cudaMemcpyDeviceToDevice, sh->stream));
    }
    CS_CUDA(cudaStreamSynchronize(sh->stream));
    cudaFree(sh->rows_bf16); cudaFree(sh->ids); cudaFree(sh->status); cudaFree(sh->tags);
    sh->rows_bf16 = nrows; sh->ids = nids; sh->status = nst; sh->tags = ntags; sh->cap = ncap;
    return CSGPU_OK;
}

// room for `pending` fp32 rows in the staging area
int bf16_reserve_stage(const csgpu_index *ix, Shard *sh, uint64_t pending)
{
    if (pending <= sh->stage_cap) return CSGPU_OK;
    DeviceGuard g(sh->device);
    const uint64_t have = sh->n_total - sh->n_built;
    uint64_t ncap = std::max<uint64_t>(std::max<uint64_t>(pending, sh->stage_cap * 2), 1024);
    float *ns = nullptr;
    cudaError_t e = cudaMalloc(&ns, ncap * (size_t)ix->dim * sizeof(float));
    if (e != cudaSuccess && ncap > pending) { cudaGetLastError(); ncap = pending; e = cudaMalloc(&ns, ncap * (size_t)ix->dim * sizeof(float)); }
    if (e != cudaSuccess) return fail_cuda(e, "cudaMalloc(stage)", __FILE__, __LINE__);
    if (have) CS_CUDA(cudaMemcpyAsync(ns, sh->stage, have * (size_t)ix->dim * sizeof(float), cudaMemcpyDeviceToDevice, sh->stream));
    CS_CUDA(cudaStreamSynchronize(sh->stream));
    cudaFree(sh->stage);
    sh->stage = ns; sh->stage_cap = ncap;
    return CSGPU_OK;
}

// normalise + convert pending rows; leaves status flags for the generic compaction in csgpu.cu
int bf16_convert_pending(const csgpu_index *ix, Shard *sh)
{
    const uint64_t pending = sh->n_total - sh->n_built;
    if (!pending) return CSGPU_OK;
    DeviceGuard g(sh->device);
    const uint32_t grid = (uint32_t)std::min<uint64_t>((pending + 7) / 8, (uint64_t)sh->sm_count * 8);
    normalise_to_bf16_kernel<<<grid, 256, 0, sh->stream>>>(sh->stage, reinterpret_cast<__nv_bfloat16 *>(sh->rows_bf16), sh->status,
                                                           sh->n_built, pending, ix->dim);
    count_launch();
    CS_CUDA(cudaGetLastError());
    CS_CUDA(cudaStreamSynchronize(sh->stream));
    cudaFree(sh->stage);   // staging is only needed between append and build
    sh->stage = nullptr; sh->stage_cap = 0;
    return CSGPU_OK;
}

// (re)encode the TMA map over the shard's built rows; the base pointer can move at every reserve
int batch_after_build(const csgpu_index *ix, Shard *sh)
{
    sh->map_valid = false;
    if (ix->dtype == CSGPU_DTYPE_BF16) {
        int rc = make_map(&sh->map_c, sh->rows_bf16 ? sh->rows_bf16 : (void *)sh->ids, sh->n_built, ix->dim, GT_BLOCK_N, false);
        if (!rc) rc = make_map(&sh->map_c2, sh->rows_bf16 ? sh->rows_bf16 : (void *)sh->ids, sh->n_built, ix->dim, GT_BLOCK_N / 2, false);
        if (rc) return rc;
    } else {
        int rc = shadow_refresh(ix, sh);
        if (rc) return rc;
        if (!batch_f32_dim_supported(ix->dim_pad) || sh->rows == nullptr) return CSGPU_OK;
        rc = make_map(&sh->map_c, sh->rows, sh->n_built, ix->dim_pad, GS_BN, true);                 // 128-query tile: 192 rows
        if (!rc) rc = make_map(&sh->map_c2, sh->rows, sh->n_built, ix->dim_pad, GS_BN_SMALL, true);  // 64-query tile: 256 rows
        if (rc) return rc;
    }
    sh->map_valid = true;
    return CSGPU_OK;
}

// (re)build or drop the bf16 shadow of an fp32 shard's built rows (csgpu_set_tensor_prefilter; rescore.cuh)
int shadow_refresh(const csgpu_index *ix, Shard *sh)
{
    DeviceGuard g(sh->device);
    sh->shadow_valid = false;
    cudaFree(sh->shadow_bf16);
    sh->shadow_bf16 = nullptr;
    sh->shadow_rows = 0;
    if (!ix->tensor_prefilter || ix->dtype != CSGPU_DTYPE_F32 || sh->n_built == 0 || sh->rows == nullptr) return CSGPU_OK;
    if (!bf16_dim_supported(ix->dim)) return fail(CSGPU_ERR_ARG, "tensor prefilter needs dim % 64 == 0 and 64 <= dim <= 512");
    cudaError_t e = cudaMalloc(&sh->shadow_bf16, sh->n_built * (size_t)ix->dim * 2);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(CSGPU_ERR_OOM, "no room in HBM for the bf16 shadow of the fp32 rows (tensor prefilter needs +50 %)"); }
    const uint64_t n4 = sh->n_built * (uint64_t)ix->dim4;
    const uint32_t grid = (uint32_t)std::min<uint64_t>((n4 + 255) / 256, (uint64_t)sh->sm_count * 16);
    shadow_from_rows_kernel<<<grid, 256, 0, sh->stream>>>(reinterpret_cast<const float4 *>(sh->rows), reinterpret_cast<uint2 *>(sh->shadow_bf16), n4);
    count_launch();
    CS_CUDA(cudaGetLastError());
    CS_CUDA(cudaStreamSynchronize(sh->stream));
    int rc = make_map(&sh->map_shadow, sh->shadow_bf16, sh->n_built, ix->dim, GT_BLOCK_N, false);
    if (rc) return rc;
    sh->shadow_rows = sh->n_built;
    sh->shadow_valid = true;
    return CSGPU_OK;
}

bool batch_f32_dim_supported(uint32_t dim_pad) { return dim_pad >= GS_BK; }

void batch_free_ctx(Shard *sh)
{
    if (!sh->batch) return;
    DeviceGuard g(sh->device);
    BatchCtx *c = sh->batch;
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->e0) cudaEventDestroy(c->e0);
    if (c->e1) cudaEventDestroy(c->e1);
    cudaFree(c->q_f32); cudaFree(c->q_prep); cudaFree(c->flags); cudaFree(c->thr); cudaFree(c->count); cudaFree(c->count_saved);
    cudaFree(c->cand); cudaFree(c->out); cudaFree(c->scalar); cudaFree(c->seg_count);
    cudaFreeHost(c->q_pin); cudaFreeHost(c->out_pin);
    delete c;
    sh->batch = nullptr;
}

static int batch_ctx(const csgpu_index *ix, Shard *sh, BatchCtx **out)
{
    if (sh->batch) { *out = sh->batch; return CSGPU_OK; }
    DeviceGuard g(sh->device);
    BatchCtx *c = new BatchCtx();
    sh->batch = c;
    const size_t nq = (size_t)BF_MAX_QBLOCKS * GT_BLOCK_M;
    CS_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CS_CUDA(cudaEventCreate(&c->e0));
    CS_CUDA(cudaEventCreate(&c->e1));
    CS_CUDA(cudaMalloc(&c->q_f32, nq * ix->dim_pad * sizeof(float)));   // raw queries, row pitch dim_pad (zero-padded like the scan kernel's query)
    CS_CUDA(cudaMalloc(&c->q_prep, nq * ix->dim_pad * sizeof(float)));   // bf16 uses half of it
    CS_CUDA(cudaMalloc(&c->flags, nq));
    CS_CUDA(cudaMalloc(&c->thr, nq * sizeof(float)));
    CS_CUDA(cudaMalloc(&c->count, nq * sizeof(unsigned)));
    CS_CUDA(cudaMalloc(&c->count_saved, nq * sizeof(unsigned)));
    CS_CUDA(cudaMalloc(&c->cand, nq * BF_CAP * sizeof(uint64_t)));
    CS_CUDA(cudaMalloc(&c->out, nq * CSGPU_MAX_K * sizeof(uint64_t)));
    CS_CUDA(cudaMalloc(&c->scalar, 64));
    CS_CUDA(cudaMalloc(&c->seg_count, nq * 2 * 148 * sizeof(unsigned)));
    CS_CUDA(cudaHostAlloc(&c->q_pin, nq * ix->dim_pad * sizeof(float), cudaHostAllocDefault));
    CS_CUDA(cudaHostAlloc(&c->out_pin, nq * CSGPU_MAX_K * sizeof(uint64_t) + nq, cudaHostAllocDefault));
    *out = c;
    return CSGPU_OK;
}

// ---------------------------------------------------------------------------------------------
// search
// ---------------------------------------------------------------------------------------------
// the SIMT kernel has two tile shapes: <= 64 queries take the 64-query x 256-row tile instead of a half-empty 128-query one
static bool simt_small(uint32_t nq) { return nq <= (uint32_t)GS_BM_SMALL; }
static uint32_t tile_rows_of(const csgpu_index *ix, const Shard *sh, uint32_t nq)
{
    return tc_path(ix, sh) ? GT_BLOCK_N : (simt_small(nq) ? GS_BN_SMALL : GS_BN);
}

// epilogue shape of the tensor-core kernel: threads per query. CSGPU_TC_EPI=1 selects the 4-warp epilogue for experiments.
static uint32_t tc_epi_halves()
{
    static const uint32_t h = (getenv("CSGPU_TC_EPI") && getenv("CSGPU_TC_EPI")[0] == '1') ? 1u : 2u;
    return h;
}

// Candidate layout of one launch (tensor-core kernel: segments; SIMT kernel: one global-atomic buffer per query).
struct CandLayout {
    uint32_t stride, seg_len, n_seg, groups;
};
static CandLayout cand_layout(const csgpu_index *ix, const Shard *sh, uint32_t n_qblocks, uint64_t n_tiles)
{
    CandLayout l;
    if (!tc_path(ix, sh)) { l.stride = BF_CAP; l.seg_len = 0; l.n_seg = 0; l.groups = 0; return l; }
    const uint32_t nq_pad = n_qblocks * GT_BLOCK_M;
    l.stride = (uint32_t)((uint64_t)BF_MAX_QBLOCKS * GT_BLOCK_M * BF_CAP / nq_pad);   // fewer queries -> longer blocks
    l.groups = (uint32_t)std::min<uint64_t>(std::max<uint32_t>(std::min<uint32_t>(sh->sm_count, 148) / n_qblocks, 1), n_tiles);
    l.n_seg = tc_epi_halves() * l.groups;
    l.seg_len = std::min<uint32_t>((l.stride - GT_SURV) / l.n_seg, 4096);
    return l;
}

static int launch_gemm(const csgpu_index *ix, Shard *sh, BatchCtx *c, const CUtensorMap &map_q, uint32_t n_qblocks, uint32_t nq,
                       uint64_t t0, uint64_t t1)
{
    GemmTopkArgs a;
    a.n_rows = sh->n_built;
    a.tile_begin = t0;
    a.tile_end = t1;
    a.n_qblocks = n_qblocks;
    a.thr = c->thr;
    a.cand = c->cand;
    a.count = c->count;
    a.cap = BF_CAP;
    // L2 prefetch of upcoming tiles: measured slower than no prefetch at every distance (profiles/r01_bf16_prefetch.txt:
    // 8.5 ms off vs 8.9-9.2 ms at distance 1-4), so it is off; CSGPU_BF16_PREFETCH=<tiles> re-enables it for experiments
    static const int env_pf = getenv("CSGPU_BF16_PREFETCH") ? atoi(getenv("CSGPU_BF16_PREFETCH")) : 0;
    a.prefetch_tiles = env_pf > 0 ? (uint32_t)env_pf : 0u;
    const uint64_t n_tiles = t1 - t0;
    const CandLayout lay = cand_layout(ix, sh, n_qblocks, n_tiles);
    a.stride = lay.stride; a.seg_len = lay.seg_len; a.n_seg = lay.n_seg;
    a.seg_count = c->seg_count;
    a.overflow = c->scalar;
    cudaError_t e = cudaSuccess;
    const Contraction con = contraction_of(ix, sh);
    if (con == Contraction::TC_TF32) {
        a.n_kchunks = (ix->dim_pad + TF_BLOCK_K - 1) / TF_BLOCK_K;
        a.q_rows = n_qblocks == 1 ? tf32_q_rows(nq) : (uint32_t)GT_BLOCK_M;
        constexpr int STAGES = 4;   // 4 x 48 KB
        const size_t smem = (size_t)STAGES * TF_STAGE_BYTES + 1024;
        const uint32_t grid = lay.groups * n_qblocks;
        const uint32_t halves = tc_epi_halves();
        const uint32_t threads = 64 + 128 * halves;
        if (halves == 2) {
            e = cudaFuncSetAttribute(gemm_tf32_topk_kernel<STAGES, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) gemm_tf32_topk_kernel<STAGES, 2><<<grid, threads, smem, c->stream>>>(map_q, sh->map_c2, a);
        } else {
            e = cudaFuncSetAttribute(gemm_tf32_topk_kernel<STAGES, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) gemm_tf32_topk_kernel<STAGES, 1><<<grid, threads, smem, c->stream>>>(map_q, sh->map_c2, a);
        }
    } else if (con == Contraction::TC_BF16) {
        const CUtensorMap &map_rows = ix->dtype == CSGPU_DTYPE_BF16 ? sh->map_c : sh->map_shadow;
        a.n_kchunks = ix->dim / GT_BLOCK_K;
        const size_t q_bytes = (size_t)a.n_kchunks * GT_QCHUNK_BYTES;
        const size_t avail = 227 * 1024 - 1024 /*alignment slack*/ - 256 /*static*/ - q_bytes;
        const int stages = (int)std::min<size_t>(4, avail / GT_STAGE_BYTES);
        if (stages < 2) return fail(CSGPU_ERR_ARG, "dim too large for the bf16 kernel's shared-memory plan");
        const size_t smem = q_bytes + (size_t)stages * GT_STAGE_BYTES + 1024;
        // cta_group::2 CTA pairs (M = 256) were built and measured 4-7 % slower than this one-CTA kernel
        // (profiles/r01_bf16_2cta.txt; the variant lives in git history, commit a9ec50f)
        const uint32_t grid = lay.groups * n_qblocks;
        const uint32_t halves = tc_epi_halves();
        const uint32_t threads = 64 + 128 * halves;
#define CS_GTK(S, H)                                                                                                  \
    do {                                                                                                              \
        e = cudaFuncSetAttribute(gemm_topk_kernel<S, H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
        if (e == cudaSuccess) gemm_topk_kernel<S, H><<<grid, threads, smem, c->stream>>>(map_q, map_rows, a);         \
    } while (0)
#define CS_GT(S) case S: if (halves == 2) CS_GTK(S, 2); else CS_GTK(S, 1); break;
        switch (stages) { CS_GT(2) CS_GT(3) CS_GT(4) }
#undef CS_GT
#undef CS_GTK
    } else {
        a.n_kchunks = (ix->dim_pad + GS_BK - 1) / GS_BK;
        constexpr int STAGES = 5;   // 5 x 40 KB
        if (simt_small(nq)) {
            a.n_qblocks = 1;        // one 64-query block (the q map's box is 64 rows)
            const size_t smem = (size_t)STAGES * gs_stage_bytes(GS_TM_SMALL, GS_TN_SMALL) + 1024;
            const uint32_t grid = (uint32_t)std::min<uint64_t>((uint64_t)sh->sm_count, n_tiles);
            auto kern = gemm_simt_topk_kernel<GS_TM_SMALL, GS_TN_SMALL, STAGES>;
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) kern<<<grid, GS_THREADS, smem, c->stream>>>(map_q, sh->map_c2, a);
        } else {
            const size_t smem = (size_t)STAGES * gs_stage_bytes(GS_TM, GS_TN) + 1024;
            const uint32_t grid = (uint32_t)std::min<uint64_t>((uint64_t)sh->sm_count, n_tiles * n_qblocks);
            auto kern = gemm_simt_topk_kernel<GS_TM, GS_TN, STAGES>;
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) kern<<<grid, GS_THREADS, smem, c->stream>>>(map_q, sh->map_c, a);
        }
    }
    count_launch();
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return fail_cuda(e, "gemm top-k kernel launch", __FILE__, __LINE__);
    return CSGPU_OK;
}

static int launch_select(const csgpu_index *ix, Shard *sh, BatchCtx *c, uint32_t nq_pad, uint32_t nq, uint32_t k, bool final,
                         const CandLayout &lay)
{
    SelectArgs sa;
    sa.cand = c->cand; sa.count = c->count; sa.n_done = c->count_saved;
    sa.seg_count = lay.n_seg ? c->seg_count : nullptr;
    sa.stride = lay.stride; sa.surv = GT_SURV; sa.seg_len = lay.seg_len; sa.n_seg = lay.n_seg;
    sa.overflow = c->scalar; sa.thr = c->thr; sa.ids = sh->ids; sa.flags = c->flags; sa.k = k; sa.n_active = nq;
    const bool inject_zero = final && sh == ix->shards[0];   // the zero-norm id list lives on (and is injected by) shard 0 only
    sa.zero_ids = inject_zero ? ix->zero_ids_dev : nullptr;
    sa.n_zero = inject_zero ? (uint32_t)ix->zero_ids.size() : 0;
    sa.final_out = final ? c->out : nullptr;
    sa.rows = reinterpret_cast<const float4 *>(sh->rows); sa.dim4 = ix->dim4; sa.q_raw = c->q_f32;
    sa.n_rescored = reinterpret_cast<unsigned long long *>(c->scalar + 2);
    sa.margin = contraction_of(ix, sh) == Contraction::TC_TF32 ? TF_MARGIN : TC_MARGIN;
    sa.max_err = c->scalar + 4;
    cudaError_t e = cudaSuccess;
    if (rescore_path(ix, sh)) {   // exact fp32 rescoring of the tensor-core filter's survivors (rescore.cuh)
        const size_t smem = (size_t)(SEL_BUF + 1024) * sizeof(uint64_t);
        const uint32_t V = (ix->dim4 + 31) / 32;
        const bool exact = ix->dim4 % 32 == 0;
#define CS_RS(v, ex)                                                                                                         \
    do {                                                                                                                     \
        e = cudaFuncSetAttribute(select_sorted_kernel<v, ex, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
        if (e == cudaSuccess) select_sorted_kernel<v, ex, true><<<nq_pad, SCAN_THREADS, smem, c->stream>>>(sa);               \
    } while (0)
#define CS_RSV(v) case v: if (exact) CS_RS(v, true); else CS_RS(v, false); break;
        switch (V) {
            CS_RSV(1) CS_RSV(2) CS_RSV(3) CS_RSV(4) CS_RSV(5) CS_RSV(6) CS_RSV(7) CS_RSV(8)
            default: return fail(CSGPU_ERR_ARG, "tensor-core filter: unsupported dim");
        }
#undef CS_RSV
#undef CS_RS
    } else {
        const size_t smem = (size_t)SEL_BUF * sizeof(uint64_t);
        e = cudaFuncSetAttribute(select_sorted_kernel<1, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) select_sorted_kernel<1, true, false><<<nq_pad, SCAN_THREADS, smem, c->stream>>>(sa);
    }
    count_launch();
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return fail_cuda(e, "select_sorted_kernel launch", __FILE__, __LINE__);
    return CSGPU_OK;
}

// scan tiles [t0, t1) with the current thresholds, then reduce every query to its exact top-k so far.
// careful = false (the normal run): nothing is synchronised; an overflow only raises the sticky flag c->scalar[0],
// which the caller reads once with the results and, if set, repeats the batch carefully.
// careful = true: the flag is read after the contraction; on overflow the range is rolled back and split.
// count_saved = entries per query before this range = the survivors of the previous select (they carry ids).
static int run_range(const csgpu_index *ix, Shard *sh, BatchCtx *c, const CUtensorMap &map_q, uint32_t n_qblocks,
                     uint32_t nq, uint32_t k, uint64_t t0, uint64_t t1, uint64_t t_last, bool careful, int depth)
{
    const uint32_t nq_pad = n_qblocks * GT_BLOCK_M;
    const CandLayout lay = cand_layout(ix, sh, n_qblocks, t1 - t0);
    CS_CUDA(cudaMemcpyAsync(c->count_saved, c->count, nq_pad * sizeof(unsigned), cudaMemcpyDeviceToDevice, c->stream));
    if (careful) CS_CUDA(cudaMemsetAsync(c->scalar, 0, sizeof(unsigned), c->stream));
    int rc = launch_gemm(ix, sh, c, map_q, n_qblocks, nq, t0, t1);
    if (rc) return rc;
    if (careful) {   // the optimistic run leaves these checks to the select kernel (same conditions, same sticky flag)
        check_counts_kernel<<<(nq_pad + 255) / 256, 256, 0, c->stream>>>(c->count, c->count_saved, lay.n_seg ? c->seg_count : nullptr,
                                                                         lay.n_seg, lay.seg_len, BF_CAP, nq_pad, c->scalar);
        count_launch();
        unsigned over = 0;
        CS_CUDA(cudaMemcpyAsync(&over, c->scalar, sizeof over, cudaMemcpyDeviceToHost, c->stream));
        CS_CUDA(cudaStreamSynchronize(c->stream));
        if (over) {
            if (t1 - t0 <= 1 || depth > 40) return fail(CSGPU_ERR_CUDA, "candidate buffer overflow on a single tile (internal error)");
            CS_CUDA(cudaMemcpyAsync(c->count, c->count_saved, nq_pad * sizeof(unsigned), cudaMemcpyDeviceToDevice, c->stream));
            const uint64_t mid = t0 + (t1 - t0) / 2;
            rc = run_range(ix, sh, c, map_q, n_qblocks, nq, k, t0, mid, t_last, true, depth + 1);
            if (rc) return rc;
            return run_range(ix, sh, c, map_q, n_qblocks, nq, k, mid, t1, t_last, true, depth + 1);
        }
    }
    return launch_select(ix, sh, c, nq_pad, nq, k, /*final=*/t1 == t_last, lay);   // the last range's select also writes the output
}

// Up to 1024 queries against one shard; final keys land in c->out [nq][k] (device); zero-norm query flags
// in flags_host[0..nq).
static int batch_search_shard(const csgpu_index *ix, Shard *sh, BatchCtx *c, const float *q_host, uint32_t nq, uint32_t k,
                              const uint8_t **flags_out)
{
    DeviceGuard g(sh->device);
    const Contraction con = contraction_of(ix, sh);
    const bool bf16 = con == Contraction::TC_BF16;   // queries go to the tensor cores as bf16
    uint32_t n_qblocks = (nq + GT_BLOCK_M - 1) / GT_BLOCK_M;
    static const bool pad_even = getenv("CSGPU_BF16_2CTA") && atoi(getenv("CSGPU_BF16_2CTA")) != 0;
    if (pad_even && ix->dtype == CSGPU_DTYPE_BF16 && n_qblocks >= 2 && (n_qblocks & 1)) ++n_qblocks;   // CTA pairs take two query blocks each
    const uint32_t nq_pad = n_qblocks * GT_BLOCK_M;
    CS_CUDA(cudaMemsetAsync(c->scalar, 0, 64, c->stream));
    // raw queries on the device with a row pitch of dim_pad floats, zero-padded: what scan_topk_kernel is given, so the
    // rescoring prologue (select_sorted_kernel) normalises exactly the same numbers
    if (ix->dim_pad == ix->dim) {
        memcpy(c->q_pin, q_host, (size_t)nq * ix->dim * sizeof(float));
    } else {
        memset(c->q_pin, 0, (size_t)nq * ix->dim_pad * sizeof(float));
        for (uint32_t j = 0; j < nq; ++j) memcpy(c->q_pin + (size_t)j * ix->dim_pad, q_host + (size_t)j * ix->dim, (size_t)ix->dim * sizeof(float));
    }
    CS_CUDA(cudaMemcpyAsync(c->q_f32, c->q_pin, (size_t)nq * ix->dim_pad * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    // query prep also initialises the per-query threshold (+inf, or -1 for padding / zero-norm queries) and candidate count
    if (bf16)
        prep_queries_bf16_kernel<<<(nq_pad * 32 + 255) / 256, 256, 0, c->stream>>>(c->q_f32, nq, ix->dim /* == dim_pad: bf16 kernels need dim % 64 == 0 */, reinterpret_cast<__nv_bfloat16 *>(c->q_prep), nq_pad, c->flags, c->thr, c->count);
    else
        prep_queries_f32_kernel<<<(nq_pad * 32 + 255) / 256, 256, 0, c->stream>>>(c->q_f32, ix->dim_pad, nq, ix->dim, ix->dim_pad, reinterpret_cast<float *>(c->q_prep), nq_pad, c->flags, c->thr, c->count);
    count_launch();
    uint8_t *flags_host = reinterpret_cast<uint8_t *>(c->out_pin) + (size_t)BF_MAX_QBLOCKS * GT_BLOCK_M * CSGPU_MAX_K * sizeof(uint64_t);
    CS_CUDA(cudaMemcpyAsync(flags_host, c->flags, nq, cudaMemcpyDeviceToHost, c->stream));   // read after the batch's final synchronisation
    *flags_out = flags_host;
    CUtensorMap map_q;
    const uint32_t q_box = con == Contraction::TC_TF32 ? (n_qblocks == 1 ? tf32_q_rows(nq) : (uint32_t)GT_BLOCK_M)
                           : (!bf16 && simt_small(nq)) ? (uint32_t)GS_BM_SMALL : (uint32_t)GT_BLOCK_M;
    int rc = make_map(&map_q, c->q_prep, nq_pad, bf16 ? ix->dim : ix->dim_pad, q_box, !bf16);
    if (rc) return rc;

    const uint32_t tile_rows = tile_rows_of(ix, sh, nq);
    const uint64_t n_tiles = (sh->n_built + tile_rows - 1) / tile_rows;
    // Each phase scans (growth-1) x everything seen so far, so ~ (growth-1) * k rows per query pass (x ~1.4 through the
    // prefilter's margin); that has to stay well inside the select kernel's sort buffer.
    static const int env_growth = getenv("CSGPU_BATCH_GROWTH") ? std::max(2, atoi(getenv("CSGPU_BATCH_GROWTH"))) : 0;
    const double per_k = rescore_path(ix, sh) ? 2.8 : 1.5;
    uint64_t growth = env_growth ? (uint64_t)env_growth
                                 : 1 + std::min<uint64_t>(BF_PHASE_GROWTH - 1, std::max<uint64_t>(1, (uint64_t)(SEL_SORT_CAP / (per_k * k))));
    // Small corpora (<= 512k rows) are bound by the fixed cost of a phase (~40 us: launch, TMEM, pipeline fill, select), not by
    // its rows: grow faster where the candidates still fit a 2048-key sort, so 100k rows are two phases instead of three
    // (profiles/r02_growth_ab.txt: 16 queries x top-10 over 100k rows 0.25 -> 0.23 ms; at 10M rows the larger factor is no
    // faster — 1024 queries: +2 % — so it stays 8 there)
    if (!env_growth && sh->n_built <= 512 * 1024)
        growth = std::max<uint64_t>(growth, 1 + std::min<uint64_t>(63, (uint64_t)(2048 / (per_k * k))));
    // Balanced phases: `growth` is the LARGEST factor the candidate buffers allow; the phase count it implies is kept, but the
    // factor is lowered to the one that reaches the end of the corpus in exactly that many phases (10M rows: 5 phases need
    // 7.03, not 8). Fewer candidates per phase then — for k = 100 they fit a 1024-key sort instead of a 2048-key one.
    double growth_f = (double)growth;
    {
        const uint64_t first = std::max<uint64_t>(1, (BF_CAP / 2) / tile_rows);
        if (!env_growth && n_tiles > first) {
            const double ratio = (double)n_tiles / (double)first;
            const double steps = std::ceil(std::log(ratio) / std::log((double)growth) - 1e-9);   // growth steps after phase 0
            growth_f = std::min((double)growth, std::max(2.0, std::pow(ratio, 1.0 / std::max(1.0, steps)) * 1.0005));
        }
    }
    unsigned scalar_host[6] = {0, 0, 0, 0, 0, 0};
    for (int attempt = 0; attempt < 2; ++attempt) {
        const bool careful = attempt == 1;
        if (careful) {   // an overflow spoiled the optimistic run: start over, checking every range
            init_thresholds_kernel<<<(nq_pad + 255) / 256, 256, 0, c->stream>>>(c->thr, c->count, c->flags, nq, nq_pad);
            count_launch();
        }
        // phase 0 lets everything through, so it must fit a segment (128 rows per tile and column half) and the
        // sort buffer on its own: <= CAP/2 rows
        uint64_t done = 0;
        uint64_t next = std::max<uint64_t>(1, (BF_CAP / 2) / tile_rows);
        while (done < n_tiles) {
            const uint64_t t1 = std::min(n_tiles, done + next);
            rc = run_range(ix, sh, c, map_q, n_qblocks, nq, k, done, t1, n_tiles, careful, 0);
            if (rc) return rc;
            done = t1;
            next = std::max<uint64_t>(1, (uint64_t)std::ceil((double)done * (growth_f - 1.0)));
        }
        if (n_tiles == 0) {   // no rows in the matrix: the select still injects the zero-norm ids and writes the output
            CS_CUDA(cudaMemcpyAsync(c->count_saved, c->count, nq_pad * sizeof(unsigned), cudaMemcpyDeviceToDevice, c->stream));
            CandLayout fin = cand_layout(ix, sh, n_qblocks, 1);
            fin.n_seg = 0;
            rc = launch_select(ix, sh, c, nq_pad, nq, k, true, fin);
            if (rc) return rc;
        }
        // one read-back per attempt: [0] overflow flag, [2..3] rows rescored, [4] largest filter error
        CS_CUDA(cudaMemcpyAsync(scalar_host, c->scalar, sizeof scalar_host, cudaMemcpyDeviceToHost, c->stream));
        CS_CUDA(cudaStreamSynchronize(c->stream));
        if (!scalar_host[0]) break;
        if (careful) return fail(CSGPU_ERR_CUDA, "candidate overflow in the careful pass (internal error)");
    }
    if (ix->dtype == CSGPU_DTYPE_BF16) {   // zero-norm queries: distance 0.0 everywhere -> the k smallest ids (after the final select wrote its empty lists)
        bool any = false;
        for (uint32_t j = 0; j < nq; ++j) {
            if (!flags_host[j]) continue;
            const uint32_t cap = ctabuf_cap(k);
            const uint32_t grid = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)sh->sm_count, (sh->n_built + 8191) / 8192));
            const bool root = sh == ix->shards[0];
            cudaError_t e = cudaFuncSetAttribute(zero_query_ids_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(ctabuf_cap(CSGPU_MAX_K) * sizeof(uint64_t)));
            if (e != cudaSuccess) return fail_cuda(e, "zero_query_ids_kernel attr", __FILE__, __LINE__);
            zero_query_ids_kernel<<<grid, SCAN_THREADS, (size_t)cap * sizeof(uint64_t), c->stream>>>(
                sh->ids, sh->n_built, root ? ix->zero_ids_dev : nullptr, root ? (uint32_t)ix->zero_ids.size() : 0u, k, cap, c->cand,
                c->scalar + 8, c->out + (size_t)j * k);
            count_launch();
            CS_CUDA(cudaGetLastError());
            any = true;
        }
        if (any) CS_CUDA(cudaStreamSynchronize(c->stream));
    }
    ix->batch_route.store(con == Contraction::TC_TF32 ? CSGPU_ROUTE_TC_TF32 : con == Contraction::TC_BF16 ? CSGPU_ROUTE_TC_BF16 : CSGPU_ROUTE_SIMT_F32);
    if (rescore_path(ix, sh)) {
        const unsigned *st = scalar_host + 2;   // [0..1] rows rescored (u64), [2] float bits of the largest |d_filter - d_f32|
        unsigned long long nres = 0;
        memcpy(&nres, st, sizeof nres);
        ix->prefilter_rescored.fetch_add(nres);   // summed over the shards of a multi-device index (reset per chunk by batch_search)
        float err = 0.f, seen = ix->filter_max_err.load();
        memcpy(&err, st + 2, sizeof err);
        while (err > seen && !ix->filter_max_err.compare_exchange_weak(seen, err)) {}
    }
    return CSGPU_OK;
}

bool batch_gemm_available(const csgpu_index *ix)
{
    for (const Shard *sh : ix->shards)
        if (!sh->map_valid) return false;
    return !ix->shards.empty();
}

bool batch_tf32_route(const csgpu_index *ix)
{
    if (!batch_gemm_available(ix)) return false;
    for (const Shard *sh : ix->shards)
        if (contraction_of(ix, sh) != Contraction::TC_TF32) return false;
    return true;
}

// b queries through the GEMM-shaped path. zero_queries (fp32 index only) receives the batch positions of zero-norm
// queries, whose outputs are left untouched for the caller to fill in.
// Multi-device index: every shard runs the whole batch against its rows concurrently (one host thread per device:
// the phase loop reads an overflow flag back at its end), the per-shard [nq][k] lists are gathered on device 0 over
// NVLink peer copies and merged there by key — the same k-way merge as the single-query path.
int batch_search(const csgpu_index *ix, const float *q, uint32_t b, uint32_t k,
                 uint32_t *out_ids, float *out_dist, uint32_t *out_n, std::vector<uint32_t> *zero_queries)
{
    const size_t G = ix->shards.size();
    if (!batch_gemm_available(ix)) return fail(CSGPU_ERR_ARG, "batched GEMM path is unavailable for this index");
    std::vector<std::unique_lock<std::mutex>> locks;   // one GEMM batch at a time per index; shards are locked in order
    std::vector<BatchCtx *> ctx(G, nullptr);
    for (size_t g = 0; g < G; ++g) {
        locks.emplace_back(ix->shards[g]->batch_mu);
        int rc = batch_ctx(ix, ix->shards[g], &ctx[g]);
        if (rc) return rc;
    }
    Shard *sh0 = ix->shards[0];
    BatchCtx *c0 = ctx[0];
    DeviceGuard g0(sh0->device);
    const uint32_t chunk = BF_MAX_QBLOCKS * GT_BLOCK_M;
    cudaEvent_t e0 = c0->e0, e1 = c0->e1;
    CS_CUDA(cudaEventRecord(e0, c0->stream));
    uint64_t *gather = nullptr;
    if (G > 1) CS_CUDA(cudaMalloc(&gather, G * (size_t)std::min(chunk, b) * k * sizeof(uint64_t)));
    int rc = CSGPU_OK;
    for (uint32_t j = 0; j < b && !rc; j += chunk) {
        const uint32_t nq = std::min(chunk, b - j);
        ix->prefilter_rescored.store(0);
        ix->filter_max_err.store(0.f);
        std::vector<const uint8_t *> flags(G, nullptr);
        if (G == 1) {
            rc = batch_search_shard(ix, sh0, c0, q + (size_t)j * ix->dim, nq, k, &flags[0]);
        } else {
            std::vector<int> rcs(G, CSGPU_OK);
            std::vector<std::string> errs(G);
            std::vector<std::thread> th;
            for (size_t g = 0; g < G; ++g)
                th.emplace_back([&, g]() {
                    rcs[g] = batch_search_shard(ix, ix->shards[g], ctx[g], q + (size_t)j * ix->dim, nq, k, &flags[g]);
                    if (rcs[g]) errs[g] = csgpu_last_error();   // the error text is thread-local: carry it over
                });
            for (auto &t : th) t.join();
            for (size_t g = 0; g < G && !rc; ++g)
                if (rcs[g]) rc = fail(rcs[g], errs[g]);
        }
        if (rc) break;
        std::vector<uint8_t> zf(flags[0], flags[0] + nq);   // out_pin is reused for the results below
        if (G > 1) {   // every shard's stream is idle here (batch_search_shard ends synchronised)
            for (size_t g = 0; g < G; ++g)
                CS_CUDA(cudaMemcpyPeerAsync(gather + g * (size_t)nq * k, sh0->device, ctx[g]->out, ix->shards[g]->device,
                                            (size_t)nq * k * sizeof(uint64_t), c0->stream));
            const bool big = k > 32;
            const uint32_t kpad = big ? pow2_at_least(k, 64) : 32;
            const size_t smem = big ? (size_t)2 * SCAN_WARPS * kpad * sizeof(uint64_t) : (size_t)SCAN_WARPS * 32 * sizeof(uint64_t);
            cudaError_t e = cudaSuccess;
            if (big) {
                // always the ceiling (k = 1024), never this launch's own size: the attribute is per function and device, and
                // a small value set here would make a later, larger merge launched from csgpu.cu fail with "invalid argument"
                e = cudaFuncSetAttribute(merge_keys_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MERGE_SMEM_MAX);
                if (e == cudaSuccess) merge_keys_kernel<true><<<nq, SCAN_THREADS, smem, c0->stream>>>(gather, (uint32_t)G, k, kpad, c0->out);
            } else {
                merge_keys_kernel<false><<<nq, SCAN_THREADS, smem, c0->stream>>>(gather, (uint32_t)G, k, kpad, c0->out);
            }
            count_launch();
            if (e == cudaSuccess) e = cudaGetLastError();
            if (e != cudaSuccess) { rc = fail_cuda(e, "merge_keys_kernel launch", __FILE__, __LINE__); break; }
        }
        CS_CUDA(cudaMemcpyAsync(c0->out_pin, c0->out, (size_t)nq * k * sizeof(uint64_t), cudaMemcpyDeviceToHost, c0->stream));
        CS_CUDA(cudaStreamSynchronize(c0->stream));
        for (uint32_t i = 0; i < nq; ++i) {
            if (zf[i] && zero_queries) { zero_queries->push_back(j + i); continue; }   // fp32 index: the caller runs the scan kernel (qzero)
            decode_keys(c0->out_pin + (size_t)i * k, k, out_ids + (size_t)(j + i) * k, out_dist + (size_t)(j + i) * k, out_n ? out_n + j + i : nullptr);
        }
    }
    cudaEventRecord(e1, c0->stream);
    cudaStreamSynchronize(c0->stream);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess) ix->last_search_us.store(ms * 1000.f);
    cudaFree(gather);
    return rc;
}

}  // namespace csgpu
