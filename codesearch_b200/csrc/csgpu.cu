// csgpu.cu — C ABI (include/csgpu.h) + index lifecycle + kernel dispatch.
//
// Mirrors the state machine of /root/reference/src/vectordb/store.rs:
//   insert (store.rs:618-686) / delete (:548-610) flip indexed=false; build_index (:386-430)
//   flips it back; search on a dirty index is an error (:440-444); clear (:690-706) empties.
// There is no CPU fallback anywhere in this file: every search is a CUDA kernel launch.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <thread>

#include "index.h"
#include "scan.cuh"
#include "scan_multi.cuh"
#include "synth.cuh"

namespace csgpu {

static thread_local std::string t_error;
std::atomic<uint64_t> g_kernel_launches{0};

void set_error(const std::string &msg) { t_error = msg; }
int fail(int code, const std::string &msg) { t_error = msg; return code; }
int fail_cuda(cudaError_t e, const char *what, const char *file, int line)
{
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    t_error = buf;
    cudaGetLastError();  // clear sticky-less errors
    return (e == cudaErrorMemoryAllocation) ? CSGPU_ERR_OOM : CSGPU_ERR_CUDA;
}


void count_launch(uint64_t n) { g_kernel_launches.fetch_add(n, std::memory_order_relaxed); }

// ---------------------------------------------------------------------------------------
// contexts
// ---------------------------------------------------------------------------------------
constexpr uint32_t MAX_GRID = 148 * 4;   // upper bound on persistent grid size we ever launch
constexpr uint32_t MAX_BATCH = 16;       // queries per multi-query scan
constexpr size_t GATHER_KEYS_PER_SHARD = 16 * 256 > 2 * CSGPU_MAX_K ? 16 * 256 : 2 * CSGPU_MAX_K;   // SearchCtx::gather
constexpr uint32_t PREFILTER_MIN_BATCH = 1;   // tensor prefilter on: every csgpu_search_batch goes to the tensor cores (9 queries:
                                              // 1.56 ms vs two multi-query passes at 3 ms each; it reads the 2-byte shadow, not the
                                              // 4-byte rows). csgpu_search itself always stays on the fp32 scan kernel.
constexpr uint32_t I8_BEATS_MULTI = 4;        // byte prefilter on: this many single int8 queries beat one fp32 multi-query pass
constexpr uint32_t GEMM_MIN_BATCH = 33;       // round 2: the 64-query SIMT tile answers <= 64 queries in 10.2 ms flat (10M x 384, top-100);
                                              // two 16-query passes (<= 32 queries) take 9.8 ms, three 14.6
                                              // (profiles/r02_bench_batch_simt.txt, r02_bench_multi16.txt)
constexpr uint32_t GEMM_MIN_BATCH_NO_MULTI = 10;  // csgpu_search_batch switches to the SIMT GEMM path from here

static int ctx_create(const csgpu_index *ix, Shard *sh, SearchCtx **out)
{
    DeviceGuard g(sh->device);
    SearchCtx *c = new SearchCtx();
    *out = c;   // set before the first fallible call: the caller destroys a partially built context
    c->device = sh->device;
    CS_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CS_CUDA(cudaEventCreateWithFlags(&c->busy, cudaEventDisableTiming));
    CS_CUDA(cudaEventCreate(&c->ev0));
    CS_CUDA(cudaEventCreate(&c->ev1));
    CS_CUDA(cudaMalloc(&c->q_dev, (size_t)MAX_BATCH * ix->dim_pad * sizeof(float)));
    CS_CUDA(cudaHostAlloc(&c->q_pin, (size_t)MAX_BATCH * ix->dim_pad * sizeof(float), cudaHostAllocDefault));
    c->cand_cap = (size_t)MAX_GRID * CSGPU_MAX_K * 2;   // single query: grid x k <= 592 x 1024; 16-query pass: 16 x 296 x 256
    CS_CUDA(cudaMalloc(&c->cand, c->cand_cap * sizeof(uint64_t)));
    // per-shard results gathered on shard 0 (peer-copy route): up to 8 shards x max(one list of 1024 keys, a 16-query pass of
    // 256 keys each) — the 16-query pass of round 2 doubled what the multi-query route can bring back per shard
    CS_CUDA(cudaMalloc(&c->gather, (size_t)8 * GATHER_KEYS_PER_SHARD * sizeof(uint64_t)));
    CS_CUDA(cudaMalloc(&c->ticket, 64 * sizeof(unsigned)));
    CS_CUDA(cudaMemset(c->ticket, 0, 64 * sizeof(unsigned)));
    CS_CUDA(cudaMalloc(&c->out_dev, (size_t)MAX_BATCH * CSGPU_MAX_K * sizeof(uint64_t)));
    // (+ 8 words: [MAX_BATCH * CSGPU_MAX_K] receives a scan's in-kernel device time, ScanArgs::elapsed_out)
    CS_CUDA(cudaHostAlloc(&c->out_pin, ((size_t)MAX_BATCH * CSGPU_MAX_K + 8) * sizeof(uint64_t),
                          cudaHostAllocMapped | cudaHostAllocPortable));
    if (ix->byte_prefilter) return i8_prepare_ctx(c);   // the scratch is born with the context, not inside its first search
    return CSGPU_OK;
}

static void ctx_destroy(SearchCtx *c)
{
    if (!c) return;
    DeviceGuard g(c->device);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->busy) cudaEventDestroy(c->busy);
    cudaFree(c->q_dev); cudaFreeHost(c->q_pin); cudaFree(c->cand); cudaFree(c->gather); cudaFree(c->ticket);
    cudaFree(c->out_dev); cudaFreeHost(c->out_pin); cudaFree(c->bitmap_dev); cudaFree(c->timing);
    i8_free_ctx(c);
    delete c;
}

static int ctx_acquire(const csgpu_index *ix, Shard *sh, SearchCtx **out)
{
    {
        std::lock_guard<std::mutex> lk(sh->ctx_mu);
        if (!sh->free_ctx.empty()) { *out = sh->free_ctx.back(); sh->free_ctx.pop_back(); return CSGPU_OK; }
    }
    SearchCtx *c = nullptr;
    int rc = ctx_create(ix, sh, &c);
    if (rc) { ctx_destroy(c); *out = nullptr; return rc; }
    std::lock_guard<std::mutex> lk(sh->ctx_mu);
    sh->all_ctx.push_back(c);
    *out = c;
    return CSGPU_OK;
}

// Device entry points (csgpu_search_keys_device & co) only ENQUEUE on the caller's stream, so the context's scratch
// (cand / ticket / work counter / int8 scratch) is in flight after the call returns. They therefore use one dedicated
// context per shard that the host-pointer searches never touch, and order successive uses with an event: the next
// device-entry call makes ITS stream wait for the previous call's work, whatever stream that was on. (Round-1 advisor
// finding: the context went back to the shared pool while its kernels were still running.) dev_ctx_begin returns with
// sh->dev_mu held; dev_ctx_end records the event and unlocks.
static int dev_ctx_begin(const csgpu_index *ix, Shard *sh, cudaStream_t st, SearchCtx **out)
{
    sh->dev_mu.lock();
    if (sh->dev_ctx == nullptr) {
        SearchCtx *c = nullptr;
        int rc = ctx_create(ix, sh, &c);
        if (rc) { ctx_destroy(c); sh->dev_mu.unlock(); return rc; }
        sh->dev_ctx = c;
    }
    SearchCtx *c = sh->dev_ctx;
    if (ix->byte_prefilter && c->i8_scratch == nullptr) {
        DeviceGuard g(sh->device);
        int rc = i8_prepare_ctx(c);
        if (rc) { sh->dev_mu.unlock(); return rc; }
    }
    if (c->busy_recorded) {
        DeviceGuard g(sh->device);
        cudaError_t e = cudaStreamWaitEvent(st, c->busy, 0);
        if (e != cudaSuccess) { sh->dev_mu.unlock(); return fail_cuda(e, "cudaStreamWaitEvent", __FILE__, __LINE__); }
    }
    *out = c;
    return CSGPU_OK;
}

static void dev_ctx_end(Shard *sh, SearchCtx *c, cudaStream_t st)
{
    {
        DeviceGuard g(sh->device);
        if (cudaEventRecord(c->busy, st) == cudaSuccess) c->busy_recorded = true; else cudaGetLastError();
    }
    sh->dev_mu.unlock();
}

static void ctx_release(Shard *sh, SearchCtx *c)
{
    std::lock_guard<std::mutex> lk(sh->ctx_mu);
    sh->free_ctx.push_back(c);
}

// one search context per shard with its int8 scratch allocated, parked in the pool: the first query pays no cudaMalloc
static int warm_i8_contexts(csgpu_index *ix)
{
    for (Shard *sh : ix->shards) {
        SearchCtx *c = nullptr;
        int rc = ctx_acquire(ix, sh, &c);
        if (!rc) { DeviceGuard dg(sh->device); rc = i8_prepare_ctx(c); ctx_release(sh, c); }
        if (rc) return rc;
    }
    return CSGPU_OK;
}

// ---------------------------------------------------------------------------------------
// scan dispatch
// ---------------------------------------------------------------------------------------
// OCC = resident CTAs per SM. The scan needs ~96 KB of loads in flight per SM (profiles/r01_tune_scan.txt:
// occupancy 1 loses 30 %); the big-k selector is one CTA-shared buffer of <= 4096 keys (32 KB), so every k runs
// 2 CTAs/SM.
// CSGPU_SCAN_STATIC=1 (diagnostic, A/B runs): fixed-stride row split instead of the work counter (scan.cuh: DYN);
// =2: fixed stride only for k > 32
static int scan_static_mode()
{
    static const int v = [] { const char *e = getenv("CSGPU_SCAN_STATIC"); return e && *e ? atoi(e) : 0; }();
    return v;
}

static bool scan_dynamic_all()
{
    static const bool v = [] { const char *e = getenv("CSGPU_SCAN_DYNAMIC_ALL"); return e && *e && atoi(e) != 0; }();
    return v;
}

template <int V, bool EXACT, bool BIG, int OCC, bool DYN>
static cudaError_t launch_scan_vd(const ScanArgs &a, uint32_t grid, size_t smem, cudaStream_t st)
{
    constexpr int R = (V <= 2) ? 8 : (V <= 4 ? 4 : 2);
    auto kern = scan_topk_kernel<V, EXACT, R, BIG, OCC, 0, false, DYN>;
    if (grid == 0) { cudaFuncAttributes fa; return cudaFuncGetAttributes(&fa, kern); }   // preload only (preload_exchange_kernels)
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    kern<<<grid, SCAN_THREADS, smem, st>>>(a);
    count_launch();
    return cudaGetLastError();
}

template <int V, bool EXACT, bool BIG, int OCC>
static cudaError_t launch_scan_v(const ScanArgs &a, uint32_t grid, size_t smem, cudaStream_t st)
{
    const int m = scan_static_mode();
    // Small corpora (k <= 32): the fixed stride wins — every warp's first rows are in flight at once instead of behind a
    // round trip to the work counter, and there is no long run for the faster SMs to steal. build/tune_scan on one B200
    // (profiles/r02_tune_scan_small.txt): 50k / 100k / 200k / 400k rows static -10 / -9 / -7 / -2 %, 1M rows equal, from 2M
    // rows on the counter is 1-2 % ahead (10M: 1.5-2.8 %, round 1). Crossover taken at 768 MB of rows per device.
    const bool small = !BIG && m == 0 && (uint64_t)a.n_rows * a.dim4 * sizeof(float4) <= (768ull << 20);
    if (m == 1 || (m == 2 && BIG) || small) return launch_scan_vd<V, EXACT, BIG, OCC, false>(a, grid, smem, st);
    return launch_scan_vd<V, EXACT, BIG, OCC, true>(a, grid, smem, st);
}

template <bool BIG, int OCC>
static cudaError_t launch_scan_b(const ScanArgs &a, uint32_t grid, size_t smem, cudaStream_t st)
{
    const uint32_t V = (a.dim4 + 31) / 32;
    const bool exact = (a.dim4 % 32) == 0;
#define CS_CASE(v)                                                                    \
    case v: return exact ? launch_scan_v<v, true, BIG, OCC>(a, grid, smem, st)         \
                         : launch_scan_v<v, false, BIG, OCC>(a, grid, smem, st);
    switch (V) {
        CS_CASE(1) CS_CASE(2) CS_CASE(3) CS_CASE(4) CS_CASE(5) CS_CASE(6) CS_CASE(7) CS_CASE(8)
        default: return cudaErrorInvalidValue;
    }
#undef CS_CASE
}

static uint32_t rows_in_flight(uint32_t dim4)
{
    const uint32_t V = (dim4 + 31) / 32;
    return (V <= 2) ? 8 : (V <= 4 ? 4 : 2);
}

// Enqueue one single-query scan of `sh` on `st`. q_dev: [dim_pad] on the shard's device.
static int enqueue_scan(const csgpu_index *ix, const Shard *sh, SearchCtx *c, const float *q_dev, uint32_t k,
                        const uint64_t *bitmap_dev, uint64_t n_bits, bool with_zero_ids, uint64_t *out_keys,
                        cudaStream_t st, const ExchangeDev *xchg = nullptr, uint32_t seq = 0,
                        const csgpu_predicate_t *pred = nullptr /* row-tag predicate; bitmap_dev is then its FILE bitmap */,
                        const unsigned *run_if = nullptr /* device word: the launch is a no-op while it is 0 */,
                        bool stamp_time = false /* in-kernel device time -> c->out_pin[MAX_BATCH * CSGPU_MAX_K] (ns) */)
{
    ScanArgs a;
    if (stamp_time) {
        a.t0_slot = reinterpret_cast<unsigned long long *>(c->ticket + 4);
        a.elapsed_out = reinterpret_cast<unsigned long long *>(c->out_pin + (size_t)MAX_BATCH * CSGPU_MAX_K);
    }
    a.run_if = run_if;
    a.rows = reinterpret_cast<const float4 *>(sh->rows);
    a.ids = sh->ids;
    a.n_rows = sh->n_built;
    a.dim4 = ix->dim4;
    a.q = q_dev;
    a.k = k;
    const bool big = k > 32;
    a.kpad = big ? ctabuf_cap(k) : 32;   // big k: capacity of the CTA-shared candidate buffer (topk.cuh CtaBuf)
    a.bitmap = bitmap_dev;
    a.n_bits = n_bits;
    a.zero_ids = with_zero_ids ? ix->zero_ids_dev : nullptr;
    a.n_zero = with_zero_ids ? (uint32_t)ix->zero_ids.size() : 0;
    a.cand = c->cand;
    a.ticket = c->ticket;
    a.out_keys = out_keys;
    a.xchg = xchg;
    a.seq = seq;
    if (pred) { a.tags = sh->tags; a.lang_mask = pred->lang_mask; a.file_lo = pred->file_lo; a.file_hi = pred->file_hi; }
    // filtered scans keep the fixed-stride block split: the work counter was built and measured for them in round 2 and is
    // NOT faster (768-d, top-200, density 1.0: 2.152 vs 2.141 ms; density 0.25: 0.70 vs 0.62 ms — most blocks are masked out
    // and cost one tag load, so the grabs' atomics and the pipeline refill at chunk boundaries outweigh the balancing;
    // profiles/r02_static_vs_dynamic_filtered_multi.txt). CSGPU_SCAN_DYNAMIC_ALL=1 switches it on for A/B runs.
    a.static_split = scan_dynamic_all() ? 0u : 1u;
    static const bool timing = getenv("CSGPU_SCAN_TIMING") != nullptr;   // diagnostic: where a scan's time goes (see scan_timing_report)
    if (timing) {
        if (c->timing == nullptr) cudaMalloc(&c->timing, (size_t)(MAX_GRID + 1) * 4 * sizeof(unsigned long long));
        a.timing = c->timing;
    }
    const uint32_t per_cta_rows = SCAN_WARPS * rows_in_flight(ix->dim4);
    uint64_t want = (sh->n_built + per_cta_rows - 1) / per_cta_rows;
    if (bitmap_dev != nullptr || pred != nullptr) {   // filtered: pre-filter variant (scan_filtered.cu)
        const uint64_t want32 = (sh->n_built + 32 * SCAN_WARPS - 1) / (32 * SCAN_WARPS);
        const uint32_t gridf = (uint32_t)std::min<uint64_t>((uint64_t)sh->sm_count * 2, std::max<uint64_t>(want32, 1));
        const size_t smemf = big ? (size_t)a.kpad * sizeof(uint64_t) : SCAN_SMALL_SMEM;
        cudaError_t ef = launch_scan_filtered(a, gridf, smemf, st);
        if (ef != cudaSuccess) return fail_cuda(ef, "scan_topk_kernel (filtered) launch", __FILE__, __LINE__);
        return CSGPU_OK;
    }
    uint32_t grid = (uint32_t)std::min<uint64_t>((uint64_t)sh->sm_count * 2, std::max<uint64_t>(want, 1));
    if (grid > MAX_GRID) grid = MAX_GRID;
    const size_t smem = big ? (size_t)a.kpad * sizeof(uint64_t) : SCAN_SMALL_SMEM;
    cudaError_t e = !big ? launch_scan_b<false, 2>(a, grid, smem, st) : launch_scan_b<true, 2>(a, grid, smem, st);
    if (e != cudaSuccess) return fail_cuda(e, "scan_topk_kernel launch", __FILE__, __LINE__);
    return CSGPU_OK;
}

// CUDA loads a kernel lazily at its first launch, and that load can wait for running kernels. A rank whose scan kernel
// is already spinning in the exchange (waiting for a peer's flag) therefore must never be the reason a peer in the SAME
// process cannot load the kernel it needs to answer — that is a 4-second timeout. Everything an exchange search can
// launch is loaded up front, when the exchange (or the byte prefilter) is set up.
static void preload_exchange_kernels(const csgpu_index *ix)
{
    for (const Shard *sh : ix->shards) {
        DeviceGuard dg(sh->device);
        ScanArgs a{};
        a.dim4 = ix->dim4;
        launch_scan_b<false, 2>(a, 0, 0, nullptr);
        launch_scan_b<true, 2>(a, 0, 0, nullptr);
        a.k = 10;  launch_scan_filtered(a, 0, 0, nullptr);
        a.k = 100; launch_scan_filtered(a, 0, 0, nullptr);
        cudaFuncAttributes fa;
        cudaFuncGetAttributes(&fa, exchange_keys_kernel<true>);
        cudaFuncGetAttributes(&fa, exchange_keys_kernel<false>);
        i8_preload(ix);
        cudaGetLastError();
    }
}

// The fused exchange as a launch of its own: `local` (this rank's sorted top-k, device) -> global top-k in out_keys.
static int enqueue_exchange(SearchCtx *c, const uint64_t *local, uint32_t k, uint64_t *out_keys, cudaStream_t st,
                            const ExchangeDev *xchg, uint32_t seq)
{
    ScanArgs a{};
    a.k = k;
    const bool big = k > 32;
    a.kpad = big ? ctabuf_cap(k) : 32;
    a.cand = c->cand;
    a.out_keys = out_keys;
    a.xchg = xchg;
    a.seq = seq;
    const size_t smem = big ? (size_t)a.kpad * sizeof(uint64_t) : SCAN_SMALL_SMEM;
    if (big) exchange_keys_kernel<true><<<1, SCAN_THREADS, smem, st>>>(a, local);
    else exchange_keys_kernel<false><<<1, SCAN_THREADS, smem, st>>>(a, local);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail_cuda(e, "exchange_keys_kernel launch", __FILE__, __LINE__);
    return CSGPU_OK;
}

// Device-resident single query on one shard: local top-k (xchg == nullptr) or the global one. With the byte prefilter
// on, the int8 kernel runs first; the fp32 scan is enqueued behind it as a launch that does nothing unless the int8
// kernel raised its device-side status word (the host cannot look without synchronising), then the exchange.
static int enqueue_keys_device(const csgpu_index *ix, const Shard *sh, SearchCtx *c, const float *q_dev, uint32_t k,
                               uint64_t *out_keys, cudaStream_t st, const ExchangeDev *xchg, uint32_t seq,
                               const uint64_t *bitmap_dev = nullptr, uint64_t n_bits = 0, const csgpu_predicate_t *pred = nullptr)
{
    if (!i8_eligible(ix, k)) return enqueue_scan(ix, sh, c, q_dev, k, bitmap_dev, n_bits, true, out_keys, st, xchg, seq, pred);
    uint64_t *local = xchg ? c->out_dev : out_keys;
    int rc = enqueue_scan_i8(ix, sh, c, q_dev, k, true, local, st, /*host_status=*/false, bitmap_dev, n_bits, pred);
    if (!rc) rc = enqueue_scan(ix, sh, c, q_dev, k, bitmap_dev, n_bits, true, local, st, nullptr, 0, pred, i8_status_dev(c));
    if (!rc && xchg) rc = enqueue_exchange(c, local, k, out_keys, st, xchg, seq);
    ix->byte_searches.fetch_add(1, std::memory_order_relaxed);
    return rc;
}

static int enqueue_merge(const uint64_t *keys_dev, uint32_t n_lists, uint32_t nq, uint32_t k, uint64_t *out, cudaStream_t st)
{
    const bool big = k > 32;
    const uint32_t kpad = big ? pow2_at_least(k, 64) : 32;
    const size_t smem = big ? (size_t)2 * SCAN_WARPS * kpad * sizeof(uint64_t) : (size_t)SCAN_WARPS * 32 * sizeof(uint64_t);
    cudaError_t e;
    if (big) {
        // always the ceiling (k = 1024): the attribute is per function and device, so per-launch values set by concurrent
        // callers (or a smaller one set elsewhere) must never undercut a launch in flight
        e = cudaFuncSetAttribute(merge_keys_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MERGE_SMEM_MAX);
        if (e != cudaSuccess) return fail_cuda(e, "merge attr", __FILE__, __LINE__);
        merge_keys_kernel<true><<<nq, SCAN_THREADS, smem, st>>>(keys_dev, n_lists, k, kpad, out);
    } else {
        merge_keys_kernel<false><<<nq, SCAN_THREADS, smem, st>>>(keys_dev, n_lists, k, kpad, out);
    }
    count_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail_cuda(e, "merge_keys_kernel launch", __FILE__, __LINE__);
    return CSGPU_OK;
}

// Enqueue one multi-query (nq <= 8) scan of `sh`. q_dev: [nq, dim_pad]; out_keys: [nq, k].
static int enqueue_scan_multi(const csgpu_index *ix, const Shard *sh, SearchCtx *c, const float *q_dev, uint32_t nq,
                              uint32_t k, bool with_zero_ids, uint64_t *out_keys, cudaStream_t st,
                              const csgpu_predicate_t *pred = nullptr /* row-tag predicate */,
                              const uint64_t *file_bitmap_dev = nullptr, uint64_t n_file_bits = 0,
                              bool stamp_time = false /* in-kernel device time -> c->out_pin[MAX_BATCH * CSGPU_MAX_K] (ns) */)
{
    MultiArgs a;
    if (stamp_time) {
        a.t0_slot = reinterpret_cast<unsigned long long *>(c->ticket + 4);
        a.elapsed_out = reinterpret_cast<unsigned long long *>(c->out_pin + (size_t)MAX_BATCH * CSGPU_MAX_K);
    }
    a.rows = reinterpret_cast<const float4 *>(sh->rows);
    a.ids = sh->ids;
    a.n_rows = sh->n_built;
    a.dim4 = ix->dim4;
    a.q = q_dev;
    a.nq = nq;
    a.k = k;
    a.bitmap = nullptr;
    a.n_bits = 0;
    if (pred) {
        a.tags = sh->tags; a.lang_mask = pred->lang_mask; a.file_lo = pred->file_lo; a.file_hi = pred->file_hi;
        a.bitmap = file_bitmap_dev; a.n_bits = n_file_bits;
    }
    a.zero_ids = with_zero_ids ? ix->zero_ids_dev : nullptr;
    a.n_zero = with_zero_ids ? (uint32_t)ix->zero_ids.size() : 0;
    a.cand = c->cand;
    a.ticket = c->ticket;
    a.out_keys = out_keys;
    // measured (profiles/r02_static_vs_dynamic_filtered_multi.txt): the work counter is worth ~1.5 % with the CTA-shared
    // buffers (8 queries x top-100: 3.26 vs 3.31 ms) and costs ~2 % with the per-warp lists (top-10: 2.81 vs 2.74 ms), so
    // each kernel gets what is faster
    // k > 32 (CTA buffers): work counter, except on small corpora (same crossover as the single-query scan, launch_scan_v)
    const bool small_corpus = (uint64_t)sh->n_built * ix->dim4 * sizeof(float4) <= (768ull << 20);
    a.static_split = scan_static_mode() == 1 ? 1u : (scan_dynamic_all() ? 0u : ((k > 32 && !small_corpus) ? 0u : 1u));
    const uint32_t R = multi_scan_rows_per_iter(ix->dim4, nq);
    const uint64_t want = (sh->n_built + SCAN_WARPS * R - 1) / (SCAN_WARPS * R);
    // the buffer capacity depends on the grid (the last CTA takes one key per CTA and column) and the grid on what fits an SM
    uint32_t per_sm = 2;
    a.kpad = multi_scan_cap(k, nq, (uint32_t)sh->sm_count * per_sm);
    per_sm = multi_scan_ctas_per_sm(ix->dim4, k, nq, a.kpad);
    if (per_sm == 1) a.kpad = multi_scan_cap(k, nq, (uint32_t)sh->sm_count);
    uint32_t grid = (uint32_t)std::min<uint64_t>((uint64_t)sh->sm_count * per_sm, std::max<uint64_t>(want, 1));
    // cand holds nq x grid x k keys: never launch more CTAs than the scratch was sized for (a part with more SMs than 148)
    grid = (uint32_t)std::min<uint64_t>(grid, std::max<uint64_t>(c->cand_cap / ((uint64_t)nq * std::max(k, 1u)), 1));
    cudaError_t e = launch_scan_multi(a, grid, st);
    if (e != cudaSuccess) return fail_cuda(e, "scan_multi_topk_kernel launch", __FILE__, __LINE__);
    return CSGPU_OK;
}

// ---------------------------------------------------------------------------------------
// storage
// ---------------------------------------------------------------------------------------
static int shard_reserve(const csgpu_index *ix, Shard *sh, uint64_t rows)
{
    if (ix->dtype == CSGPU_DTYPE_BF16) return bf16_reserve_rows(ix, sh, rows);
    if (rows <= sh->cap) return CSGPU_OK;
    DeviceGuard g(sh->device);
    uint64_t ncap = std::max<uint64_t>(rows, sh->cap + sh->cap / 2);
    ncap = std::max<uint64_t>(ncap, 1024);
    float *nrows = nullptr; uint32_t *nids = nullptr, *ntags = nullptr; uint8_t *nst = nullptr;
    const size_t row_bytes = (size_t)ix->dim_pad * sizeof(float);
    cudaError_t e = cudaMalloc(&nrows, ncap * row_bytes);
    if (e != cudaSuccess && ncap > rows) {  // retry with the exact size before giving up
        cudaGetLastError();
        ncap = rows;
        e = cudaMalloc(&nrows, ncap * row_bytes);
    }
    if (e != cudaSuccess) return fail_cuda(e, "cudaMalloc(rows)", __FILE__, __LINE__);
    if ((e = cudaMalloc(&nids, ncap * sizeof(uint32_t))) != cudaSuccess) { cudaFree(nrows); return fail_cuda(e, "cudaMalloc(ids)", __FILE__, __LINE__); }
    if ((e = cudaMalloc(&nst, ncap)) != cudaSuccess) { cudaFree(nrows); cudaFree(nids); return fail_cuda(e, "cudaMalloc(status)", __FILE__, __LINE__); }
    if ((e = cudaMalloc(&ntags, ncap * sizeof(uint32_t))) != cudaSuccess) { cudaFree(nrows); cudaFree(nids); cudaFree(nst); return fail_cuda(e, "cudaMalloc(tags)", __FILE__, __LINE__); }
    CS_CUDA(cudaMemsetAsync(nst, 0, ncap, sh->stream));
    CS_CUDA(cudaMemsetAsync(ntags, 0xFF, ncap * sizeof(uint32_t), sh->stream));   // CSGPU_TAG_NONE
    if (sh->n_total) {
        CS_CUDA(cudaMemcpyAsync(nrows, sh->rows, sh->n_total * row_bytes, cudaMemcpyDeviceToDevice, sh->stream));
        CS_CUDA(cudaMemcpyAsync(nids, sh->ids, sh->n_total * sizeof(uint32_t), cudaMemcpyDeviceToDevice, sh->stream));
        CS_CUDA(cudaMemcpyAsync(ntags, sh->tags, sh->n_total * sizeof(uint32_t), cudaMemcpyDeviceToDevice, sh->stream));
        CS_CUDA(cudaMemcpyAsync(nst, sh->status, sh->n_total, cudaMemcpyDeviceToDevice, sh->stream));
    }
    CS_CUDA(cudaStreamSynchronize(sh->stream));
    cudaFree(sh->rows); cudaFree(sh->ids); cudaFree(sh->status); cudaFree(sh->tags);
    sh->rows = nrows; sh->ids = nids; sh->status = nst; sh->tags = ntags; sh->cap = ncap;
    return CSGPU_OK;
}

// Kill rows (all shards) whose id is in `ids` (any order). Returns rows killed (+ zero-norm ids dropped).
static int kill_ids(csgpu_index *ix, const uint32_t *ids, uint64_t n, uint64_t *killed)
{
    *killed = 0;
    if (n == 0) return CSGPU_OK;
    std::vector<uint32_t> sorted(ids, ids + n);
    std::sort(sorted.begin(), sorted.end());
    sorted.erase(std::unique(sorted.begin(), sorted.end()), sorted.end());
    // zero-norm rows live on the host list only
    if (!ix->zero_ids.empty()) {
        std::vector<uint32_t> keep, keep_tags;
        keep.reserve(ix->zero_ids.size());
        for (size_t i = 0; i < ix->zero_ids.size(); ++i) {
            const uint32_t z = ix->zero_ids[i];
            if (std::binary_search(sorted.begin(), sorted.end(), z)) ++*killed;
            else { keep.push_back(z); keep_tags.push_back(ix->zero_tags[i]); }
        }
        ix->zero_ids.swap(keep);
        ix->zero_tags.swap(keep_tags);
    }
    for (Shard *sh : ix->shards) {
        if (sh->n_total == 0) continue;
        DeviceGuard g(sh->device);
        uint32_t *kill_dev = nullptr;
        unsigned long long *cnt_dev = nullptr;
        CS_CUDA(cudaMalloc(&kill_dev, sorted.size() * sizeof(uint32_t)));
        CS_CUDA(cudaMalloc(&cnt_dev, sizeof(unsigned long long)));
        CS_CUDA(cudaMemsetAsync(cnt_dev, 0, sizeof(unsigned long long), sh->stream));
        CS_CUDA(cudaMemcpyAsync(kill_dev, sorted.data(), sorted.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, sh->stream));
        const uint32_t grid = (uint32_t)std::min<uint64_t>((sh->n_total + 255) / 256, (uint64_t)sh->sm_count * 8);
        mark_dead_kernel<<<grid, 256, 0, sh->stream>>>(sh->ids, sh->status, sh->n_total, kill_dev, (uint32_t)sorted.size(), cnt_dev);
        count_launch();
        unsigned long long cnt = 0;
        CS_CUDA(cudaMemcpyAsync(&cnt, cnt_dev, sizeof cnt, cudaMemcpyDeviceToHost, sh->stream));
        CS_CUDA(cudaStreamSynchronize(sh->stream));
        cudaFree(kill_dev); cudaFree(cnt_dev);
        *killed += cnt;
    }
    return CSGPU_OK;
}

static Shard *least_loaded(csgpu_index *ix)
{
    Shard *best = ix->shards[0];
    for (Shard *s : ix->shards) if (s->n_total < best->n_total) best = s;
    return best;
}

static int upload_zero_ids(csgpu_index *ix)
{
    Shard *s0 = ix->shards[0];
    DeviceGuard g(s0->device);
    cudaFree(ix->zero_ids_dev);
    ix->zero_ids_dev = nullptr;
    if (!ix->zero_ids.empty()) {   // [n_zero] ids, then [n_zero] tags
        const size_t nz = ix->zero_ids.size();
        ix->zero_tags.resize(nz, CSGPU_TAG_NONE);
        CS_CUDA(cudaMalloc(&ix->zero_ids_dev, 2 * nz * sizeof(uint32_t)));
        CS_CUDA(cudaMemcpy(ix->zero_ids_dev, ix->zero_ids.data(), nz * sizeof(uint32_t), cudaMemcpyHostToDevice));
        CS_CUDA(cudaMemcpy(ix->zero_ids_dev + nz, ix->zero_tags.data(), nz * sizeof(uint32_t), cudaMemcpyHostToDevice));
    }
    return CSGPU_OK;
}

__global__ void gather_rows_bounce_kernel(const float4 *__restrict__ rows, float4 *__restrict__ bounce,
                                          const uint32_t *__restrict__ keep_src, uint64_t n, uint32_t dim4)
{
    const int lane = threadIdx.x & 31;
    const uint64_t w0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t i = w0; i < n; i += nw) {
        const float4 *s = rows + (uint64_t)keep_src[i] * dim4;
        float4 *o = bounce + i * dim4;
        for (uint32_t c = lane; c < dim4; c += 32) o[c] = s[c];
    }
}
__global__ void gather_u32_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out,
                                  const uint32_t *__restrict__ keep_src, uint64_t n)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        out[i] = in[keep_src[i]];
}

// Normalise pending rows, drop dead/zero/non-finite rows (stable), leave [0, n_built) all live.
static int shard_build(csgpu_index *ix, Shard *sh, std::vector<std::pair<uint32_t, uint32_t>> &new_zero_ids /* (id, tag) */)
{
    DeviceGuard g(sh->device);
    const uint64_t n = sh->n_total;
    if (n == 0) { sh->n_built = 0; return CSGPU_OK; }
    const uint64_t pending = n - sh->n_built;
    const bool bf16 = ix->dtype == CSGPU_DTYPE_BF16;
    char *rowbase = bf16 ? reinterpret_cast<char *>(sh->rows_bf16) : reinterpret_cast<char *>(sh->rows);
    const size_t row_bytes = bf16 ? (size_t)ix->dim * 2 : (size_t)ix->dim_pad * sizeof(float);
    const uint32_t row_u4 = (uint32_t)(row_bytes / 16);
    if (bf16) {
        int rc = bf16_convert_pending(ix, sh);
        if (rc) return rc;
    } else if (pending) {
        const uint32_t grid = (uint32_t)std::min<uint64_t>((pending + 7) / 8, (uint64_t)sh->sm_count * 8);
        normalise_rows_kernel<<<grid, 256, 0, sh->stream>>>(reinterpret_cast<float4 *>(sh->rows), sh->status, sh->n_built, pending, ix->dim4);
        count_launch();
        CS_CUDA(cudaGetLastError());
    }
    std::vector<uint8_t> st(n);
    CS_CUDA(cudaMemcpyAsync(st.data(), sh->status, n, cudaMemcpyDeviceToHost, sh->stream));
    CS_CUDA(cudaStreamSynchronize(sh->stream));
    uint64_t first_bad = n;
    for (uint64_t i = 0; i < n; ++i) if (st[i] != ROW_OK) { first_bad = i; break; }
    if (first_bad == n) { sh->n_built = n; return CSGPU_OK; }

    // ids of zero-norm rows go to the host list; count non-finite
    std::vector<uint32_t> keep_src;
    keep_src.reserve(n - first_bad);
    std::vector<uint64_t> zero_rows;
    for (uint64_t i = first_bad; i < n; ++i) {
        if (st[i] == ROW_OK) keep_src.push_back((uint32_t)i);
        else if (st[i] == ROW_ZERO) zero_rows.push_back(i);
        else if (st[i] == ROW_NONFINITE) ix->nonfinite_rows++;
    }
    for (uint64_t r : zero_rows) {
        uint32_t id, tag;
        CS_CUDA(cudaMemcpy(&id, sh->ids + r, sizeof id, cudaMemcpyDeviceToHost));
        CS_CUDA(cudaMemcpy(&tag, sh->tags + r, sizeof tag, cudaMemcpyDeviceToHost));
        new_zero_ids.push_back({id, tag});
    }
    const uint64_t moved = keep_src.size();
    if (moved) {
        const uint64_t CH = 65536;  // rows per bounce chunk
        uint32_t *keep_dev = nullptr; float *bounce = nullptr; uint32_t *bounce_ids = nullptr, *bounce_tags = nullptr;
        CS_CUDA(cudaMalloc(&keep_dev, moved * sizeof(uint32_t)));
        CS_CUDA(cudaMalloc(&bounce, std::min(CH, moved) * row_bytes));
        CS_CUDA(cudaMalloc(&bounce_ids, std::min(CH, moved) * sizeof(uint32_t)));
        CS_CUDA(cudaMalloc(&bounce_tags, std::min(CH, moved) * sizeof(uint32_t)));
        CS_CUDA(cudaMemcpyAsync(keep_dev, keep_src.data(), moved * sizeof(uint32_t), cudaMemcpyHostToDevice, sh->stream));
        for (uint64_t c0 = 0; c0 < moved; c0 += CH) {
            const uint64_t cn = std::min(CH, moved - c0);
            const uint32_t grid = (uint32_t)std::min<uint64_t>((cn + 7) / 8, (uint64_t)sh->sm_count * 8);
            gather_rows_bounce_kernel<<<grid, 256, 0, sh->stream>>>(reinterpret_cast<const float4 *>(rowbase),
                                                                     reinterpret_cast<float4 *>(bounce), keep_dev + c0, cn, row_u4);
            gather_u32_kernel<<<(uint32_t)((cn + 255) / 256), 256, 0, sh->stream>>>(sh->ids, bounce_ids, keep_dev + c0, cn);
            gather_u32_kernel<<<(uint32_t)((cn + 255) / 256), 256, 0, sh->stream>>>(sh->tags, bounce_tags, keep_dev + c0, cn);
            count_launch(3);
            CS_CUDA(cudaMemcpyAsync(rowbase + (first_bad + c0) * row_bytes, bounce, cn * row_bytes, cudaMemcpyDeviceToDevice, sh->stream));
            CS_CUDA(cudaMemcpyAsync(sh->ids + first_bad + c0, bounce_ids, cn * sizeof(uint32_t), cudaMemcpyDeviceToDevice, sh->stream));
            CS_CUDA(cudaMemcpyAsync(sh->tags + first_bad + c0, bounce_tags, cn * sizeof(uint32_t), cudaMemcpyDeviceToDevice, sh->stream));
        }
        CS_CUDA(cudaStreamSynchronize(sh->stream));
        cudaFree(keep_dev); cudaFree(bounce); cudaFree(bounce_ids); cudaFree(bounce_tags);
    }
    sh->n_built = sh->n_total = first_bad + moved;
    CS_CUDA(cudaMemsetAsync(sh->status, 0, sh->cap, sh->stream));
    CS_CUDA(cudaStreamSynchronize(sh->stream));
    return CSGPU_OK;
}

// CSGPU_SCAN_TIMING=1: after a host-pointer single-device search, print when (relative to the first CTA's start) the last CTA
// started, the first / last CTA finished streaming, the last CTA list was written, the last ticket was taken, the last
// CTA had merged, and the kernel was done. The search has been synchronised; a diagnostic, never a benchmark.
static void scan_timing_report(const Shard *sh, SearchCtx *c, uint64_t n_rows, uint32_t k)
{
    if (c->timing == nullptr) return;
    const uint32_t per_cta_rows = SCAN_WARPS * 4;
    uint32_t grid = (uint32_t)std::min<uint64_t>((uint64_t)sh->sm_count * 2, std::max<uint64_t>((n_rows + per_cta_rows - 1) / per_cta_rows, 1));
    std::vector<unsigned long long> t((size_t)(MAX_GRID + 1) * 4);
    if (cudaMemcpy(t.data(), c->timing, t.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess) return;
    unsigned long long t0 = ~0ull, st_hi = 0, s_lo = ~0ull, s_hi = 0, w_hi = 0;
    for (uint32_t b = 0; b < grid; ++b) {
        t0 = std::min(t0, t[b * 4]); st_hi = std::max(st_hi, t[b * 4]);
        s_lo = std::min(s_lo, t[b * 4 + 1]); s_hi = std::max(s_hi, t[b * 4 + 1]); w_hi = std::max(w_hi, t[b * 4 + 2]);
    }
    fprintf(stderr, "[scan timing] rows %llu k %u grid %u: last CTA start +%.1f us | streaming ends first +%.1f last +%.1f | last CTA list +%.1f | "
                    "last ticket seen +%.1f | merged +%.1f | done +%.1f us\n",
            (unsigned long long)n_rows, k, grid, (st_hi - t0) / 1e3, (s_lo - t0) / 1e3, (s_hi - t0) / 1e3, (w_hi - t0) / 1e3,
            (t[grid * 4] - t0) / 1e3, (t[grid * 4 + 1] - t0) / 1e3, (t[grid * 4 + 2] - t0) / 1e3);
}

static bool all_finite(const float *q, uint32_t n)
{
    for (uint32_t i = 0; i < n; ++i) if (!std::isfinite(q[i])) return false;
    return true;
}

static int check_search_args(const csgpu_index *ix, const float *q, uint32_t q_len, uint32_t k)
{
    if (!ix) return fail(CSGPU_ERR_ARG, "null index");
    if (q_len != ix->dim) {
        char buf[160];
        snprintf(buf, sizeof buf, "Query embedding dimension mismatch: expected %u, got %u", ix->dim, q_len);
        return fail(CSGPU_ERR_DIM, buf);
    }
    if (!ix->built) return fail(CSGPU_ERR_NOT_BUILT, "Index not built. Call build_index() after inserting chunks.");
    if (!q) return fail(CSGPU_ERR_ARG, "null query");
    if (k > CSGPU_MAX_K) return fail(CSGPU_ERR_ARG, "k exceeds CSGPU_MAX_K (1024)");
    return CSGPU_OK;
}

void decode_keys(const uint64_t *keys, uint32_t k, uint32_t *out_ids, float *out_dist, uint32_t *out_n)
{
    uint32_t m = 0;
    for (uint32_t i = 0; i < k; ++i) {
        if (keys[i] == KEY_EMPTY) break;
        const uint32_t bits = bits_from_okey((uint32_t)(keys[i] >> 32));
        float d;
        memcpy(&d, &bits, sizeof d);
        if (out_ids) out_ids[m] = (uint32_t)keys[i];
        if (out_dist) out_dist[m] = d;
        ++m;
    }
    if (out_n) *out_n = m;
}

// ---------------------------------------------------------------------------------------
// in-process multi-device search: wired context groups (GroupCtx in index.h)
// ---------------------------------------------------------------------------------------
static uint64_t default_exchange_timeout_ns()
{
    static const uint64_t v = [] {
        const char *e = getenv("CSGPU_EXCHANGE_TIMEOUT_MS");
        const uint64_t ms = e && *e ? (uint64_t)strtoull(e, nullptr, 10) : 4000;
        return std::max<uint64_t>(ms, 1) * 1000000ull;
    }();
    return v;
}

// every pair of shard devices can reach each other's HBM (or is the same device)? CSGPU_NO_FUSED_LOCAL=1 forces the
// peer-copy + merge-launch route for A/B runs.
static bool fused_local_ok(const csgpu_index *ix)
{
    int v = ix->fused_local.load(std::memory_order_relaxed);
    if (v >= 0) return v != 0;
    bool ok = getenv("CSGPU_NO_FUSED_LOCAL") == nullptr && ix->dtype == CSGPU_DTYPE_F32;
    for (const Shard *a : ix->shards)
        for (const Shard *b : ix->shards) {
            if (a->device == b->device) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, a->device, b->device);
            ok = ok && can;
        }
    ix->fused_local.store(ok ? 1 : 0, std::memory_order_relaxed);
    return ok;
}

static void group_destroy(GroupCtx *g)
{
    if (!g) return;
    for (size_t s = 0; s < g->ctx.size(); ++s) {
        if (g->ctx[s]) {
            DeviceGuard dg(g->ctx[s]->device);
            cudaStreamSynchronize(g->ctx[s]->stream);
            cudaFree(g->xbase[s]);
            cudaFree(g->xdev[s]);
        }
        ctx_destroy(g->ctx[s]);
    }
    if (g->status_pin) cudaFreeHost(g->status_pin);
    delete g;
}

static int group_create(const csgpu_index *ix, GroupCtx *g)
{
    const size_t G = ix->shards.size();
    g->ctx.assign(G, nullptr); g->xbase.assign(G, nullptr); g->xdev.assign(G, nullptr);
    Exchange geo;   // geometry only
    geo.world = (uint32_t)G;
    {
        DeviceGuard dg(ix->shards[0]->device);
        CS_CUDA(cudaHostAlloc(&g->status_pin, 64, cudaHostAllocMapped | cudaHostAllocPortable));
        memset(g->status_pin, 0, 64);
    }
    for (size_t s = 0; s < G; ++s) {
        Shard *sh = ix->shards[s];
        int rc = ctx_create(ix, sh, &g->ctx[s]);
        if (rc) return rc;
        DeviceGuard dg(sh->device);
        CS_CUDA(cudaMalloc(&g->xbase[s], geo.block_bytes()));
        CS_CUDA(cudaMemset(g->xbase[s], 0, geo.block_bytes()));
        CS_CUDA(cudaMalloc(reinterpret_cast<void **>(&g->xdev[s]), sizeof(ExchangeDev)));
    }
    for (size_t s = 0; s < G; ++s) {
        ExchangeDev d;
        memset(&d, 0, sizeof d);
        for (size_t p = 0; p < G; ++p) {
            d.slots[p] = reinterpret_cast<uint64_t *>(g->xbase[p]);
            d.flags[p] = reinterpret_cast<unsigned *>(reinterpret_cast<char *>(g->xbase[p]) + geo.slots_bytes());
        }
        DeviceGuard dg(ix->shards[s]->device);
        void *st_dev = nullptr;
        CS_CUDA(cudaHostGetDevicePointer(&st_dev, g->status_pin, 0));
        d.status = reinterpret_cast<unsigned *>(st_dev);
        d.wait_ring = reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(g->xbase[s]) + geo.stats_off());
        d.timeout_ns = ix->exchange_timeout_ns ? ix->exchange_timeout_ns : default_exchange_timeout_ns();
        d.world = (uint32_t)G; d.rank = (uint32_t)s; d.kmax = geo.kmax;
        d.root = 0;
        CS_CUDA(cudaMemcpy(g->xdev[s], &d, sizeof d, cudaMemcpyHostToDevice));
    }
    preload_exchange_kernels(ix);
    return CSGPU_OK;
}

static int group_acquire(const csgpu_index *ix, GroupCtx **out)
{
    {
        std::lock_guard<std::mutex> lk(ix->group_mu);
        if (!ix->free_groups.empty()) { *out = ix->free_groups.back(); ix->free_groups.pop_back(); return CSGPU_OK; }
    }
    GroupCtx *g = new GroupCtx();
    int rc = group_create(ix, g);
    if (rc) { group_destroy(g); *out = nullptr; return rc; }
    std::lock_guard<std::mutex> lk(ix->group_mu);
    ix->all_groups.push_back(g);
    *out = g;
    return CSGPU_OK;
}

static void group_release(const csgpu_index *ix, GroupCtx *g)
{
    std::lock_guard<std::mutex> lk(ix->group_mu);
    ix->free_groups.push_back(g);
}

// One query on a multi-device index, fused: N concurrent scan launches (one per shard, each on its own stream); the
// tails of shards 1..N-1 store their k keys into shard 0's exchange block over NVLink and raise a flag, shard 0's last
// CTA waits for the flags, merges and writes the global top-k straight into mapped host memory. No cudaMemcpyPeer, no
// merge launch; the host synchronises ONE stream. A group whose exchange failed (launch error part-way, timed-out
// wait) is retired, never reused: its sequence numbers are out of step.
static int search_one_fused(const csgpu_index *ix, const float *q, uint32_t k, const uint64_t *bitmap, uint64_t n_bits,
                            uint32_t *out_ids, float *out_dist, uint32_t *out_n, const csgpu_predicate_t *pred)
{
    const size_t G = ix->shards.size();
    GroupCtx *grp = nullptr;
    int rc = group_acquire(ix, &grp);
    if (rc) return rc;
    const size_t qbytes = (size_t)ix->dim_pad * sizeof(float);
    const size_t bm_words = bitmap ? (size_t)((n_bits + 63) / 64) : 0;
    const uint32_t seq = ++grp->seq;
    SearchCtx *c0 = grp->ctx[0];
    // test hook (tests/test_gpu_multishard_one_gpu.py): CSGPU_FAULT_SKIP_SHARD=<g>, g >= 1, drops shard g's launch, which is
    // what a lost device looks like to the root's wait
    const char *fault = getenv("CSGPU_FAULT_SKIP_SHARD");
    long skip = fault && *fault ? strtol(fault, nullptr, 10) : -1;
    if (skip == 0) skip = -1;   // the root itself cannot be skipped: nobody would be left to notice
    auto body = [&]() -> int {
        memset(c0->q_pin, 0, qbytes);
        memcpy(c0->q_pin, q, (size_t)ix->dim * sizeof(float));   // one pinned copy feeds every device's H2D
        for (size_t gi = 0; gi < G; ++gi) {
            const size_t g = G - 1 - gi;   // the root (shard 0) last: should anything serialise the launches (a debugger, the
                                           // sanitizer, time slicing), the kernel that waits finds the others' keys already there
            if ((long)g == skip) continue;
            Shard *sh = ix->shards[g];
            SearchCtx *c = grp->ctx[g];
            DeviceGuard dg(sh->device);
            CS_CUDA(cudaMemcpyAsync(c->q_dev, c0->q_pin, qbytes, cudaMemcpyHostToDevice, c->stream));
            const uint64_t *bm_dev = nullptr;
            if (bitmap) {
                if (c->bitmap_dev == nullptr || c->bitmap_cap < bm_words) {
                    CS_CUDA(cudaStreamSynchronize(c->stream));
                    cudaFree(c->bitmap_dev); c->bitmap_dev = nullptr; c->bitmap_cap = 0;
                    CS_CUDA(cudaMalloc(&c->bitmap_dev, std::max<size_t>(bm_words, 1) * sizeof(uint64_t)));
                    c->bitmap_cap = std::max<size_t>(bm_words, 1);
                }
                if (bm_words) CS_CUDA(cudaMemcpyAsync(c->bitmap_dev, bitmap, bm_words * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
                bm_dev = c->bitmap_dev;
            }
            if (g == 0) CS_CUDA(cudaEventRecord(c->ev0, c->stream));
            int r = enqueue_scan(ix, sh, c, c->q_dev, k, bm_dev, n_bits, /*with_zero_ids=*/g == 0, g == 0 ? c->out_pin : c->out_dev,
                                 c->stream, grp->xdev[g], seq, pred);
            if (r) return r;
        }
        DeviceGuard dg(ix->shards[0]->device);
        CS_CUDA(cudaEventRecord(c0->ev1, c0->stream));
        CS_CUDA(cudaStreamSynchronize(c0->stream));
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, c0->ev0, c0->ev1) == cudaSuccess) ix->last_search_us.store(ms * 1000.f);
        if (*reinterpret_cast<volatile unsigned *>(grp->status_pin) != 0)
            return fail(CSGPU_ERR_NCCL, "cross-GPU exchange timed out: a shard's scan never delivered its keys (device lost or hung)");
        decode_keys(c0->out_pin, k, out_ids, out_dist, out_n);
        return CSGPU_OK;
    };
    rc = body();
    if (rc) {   // drain what was launched (the root gives up after the exchange timeout), then retire the group for good
        const std::string keep = t_error;
        for (size_t g = 0; g < G; ++g) { DeviceGuard dg(ix->shards[g]->device); cudaStreamSynchronize(grp->ctx[g]->stream); }
        cudaGetLastError();
        {
            std::lock_guard<std::mutex> lk(ix->group_mu);
            auto it = std::find(ix->all_groups.begin(), ix->all_groups.end(), grp);
            if (it != ix->all_groups.end()) ix->all_groups.erase(it);
        }
        group_destroy(grp);
        t_error = keep;
        return rc;
    }
    group_release(ix, grp);
    return rc;
}

// One query (optionally filtered) through every shard; result keys land in ctx0->out_pin[0..k).
static int search_one(const csgpu_index *ix, const float *q, uint32_t k, const uint64_t *bitmap, uint64_t n_bits,
                      uint32_t *out_ids, float *out_dist, uint32_t *out_n, const csgpu_predicate_t *pred = nullptr)
{
    // byte prefilter on (csgpu_set_byte_prefilter): the unfiltered query streams the int8 shadow and rescoring makes it
    // exact (scan_i8.cuh); if that launch reports a case it cannot bound, the query is answered again by the fp32 scan
    // (round 2: filtered and tagged searches too — the int8 kernel's FILT instantiation streams only row groups the filter
    //  allows, so they move a quarter of the fp32 filtered scan's bytes at every density)
    bool use_i8 = i8_eligible(ix, k);
    if (pred) { bitmap = pred->file_bitmap; n_bits = pred->file_bitmap ? pred->n_file_bits : 0; }
    const size_t G = ix->shards.size();
    if (G > 1 && !use_i8 && fused_local_ok(ix)) return search_one_fused(ix, q, k, bitmap, n_bits, out_ids, out_dist, out_n, pred);
    std::vector<SearchCtx *> ctx(G, nullptr);
    int rc = CSGPU_OK;
    auto release_all = [&]() { for (size_t g = 0; g < G; ++g) if (ctx[g]) ctx_release(ix->shards[g], ctx[g]); };
    for (size_t g = 0; g < G && !rc; ++g) rc = ctx_acquire(ix, ix->shards[g], &ctx[g]);
    if (rc) { release_all(); return rc; }
    const size_t qbytes = (size_t)ix->dim_pad * sizeof(float);
    const size_t bm_words = bitmap ? (size_t)((n_bits + 63) / 64) : 0;

    auto body = [&]() -> int {
        for (size_t g = 0; g < G; ++g) {
            Shard *sh = ix->shards[g];
            SearchCtx *c = ctx[g];
            DeviceGuard dg(sh->device);
            memset(c->q_pin, 0, qbytes);
            memcpy(c->q_pin, q, (size_t)ix->dim * sizeof(float));
            CS_CUDA(cudaMemcpyAsync(c->q_dev, c->q_pin, qbytes, cudaMemcpyHostToDevice, c->stream));
            const uint64_t *bm_dev = nullptr;
            if (bitmap) {
                if (c->bitmap_dev == nullptr || c->bitmap_cap < bm_words) {   // n_bits == 0 still needs a (never read) non-null pointer: it selects the filtered kernel, which then excludes every id
                    CS_CUDA(cudaStreamSynchronize(c->stream));
                    cudaFree(c->bitmap_dev); c->bitmap_dev = nullptr; c->bitmap_cap = 0;
                    CS_CUDA(cudaMalloc(&c->bitmap_dev, std::max<size_t>(bm_words, 1) * sizeof(uint64_t)));
                    c->bitmap_cap = std::max<size_t>(bm_words, 1);
                }
                if (bm_words) CS_CUDA(cudaMemcpyAsync(c->bitmap_dev, bitmap, bm_words * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
                bm_dev = c->bitmap_dev;
            }
            // one device, fp32 scan: the kernel stamps its own device time (two event records cost a 100k-row query 8 of its
            // 49 us); every other shape is timed with CUDA events
            const bool stamp = G == 1 && !use_i8;
            if (g == 0 && !stamp) CS_CUDA(cudaEventRecord(c->ev0, c->stream));
            uint64_t *dst = (G == 1) ? c->out_pin : c->out_dev;
            int r = use_i8 ? enqueue_scan_i8(ix, sh, c, c->q_dev, k, /*with_zero_ids=*/g == 0, dst, c->stream, /*host_status=*/true, bm_dev, n_bits, pred)
                           : enqueue_scan(ix, sh, c, c->q_dev, k, bm_dev, n_bits, /*with_zero_ids=*/g == 0, dst, c->stream, nullptr, 0, pred, nullptr, stamp);
            if (r) return r;
        }
        SearchCtx *c0 = ctx[0];
        if (G > 1) {
            // gather every shard's k keys on shard 0 (peer copies over NVLink), merge there
            DeviceGuard dg(ix->shards[0]->device);
            uint64_t *gather = c0->gather;
            for (size_t g = 1; g < G; ++g) {
                DeviceGuard dg2(ix->shards[g]->device);
                CS_CUDA(cudaMemcpyPeerAsync(gather + g * k, ix->shards[0]->device,
                                            ctx[g]->out_dev, ix->shards[g]->device, (size_t)k * sizeof(uint64_t), ctx[g]->stream));
                CS_CUDA(cudaEventRecord(ctx[g]->ev1, ctx[g]->stream));
            }
            for (size_t g = 1; g < G; ++g) CS_CUDA(cudaStreamWaitEvent(c0->stream, ctx[g]->ev1, 0));
            CS_CUDA(cudaMemcpyAsync(gather, c0->out_dev, (size_t)k * sizeof(uint64_t), cudaMemcpyDeviceToDevice, c0->stream));
            int r = enqueue_merge(gather, (uint32_t)G, 1, k, c0->out_pin, c0->stream);
            if (r) return r;
        }
        {
            DeviceGuard dg(ix->shards[0]->device);
            const bool stamp = G == 1 && !use_i8;
            if (!stamp) CS_CUDA(cudaEventRecord(c0->ev1, c0->stream));
            CS_CUDA(cudaStreamSynchronize(c0->stream));
            float ms = 0.f;
            if (stamp) ix->last_search_us.store((float)c0->out_pin[(size_t)MAX_BATCH * CSGPU_MAX_K] * 1e-3f);
            else if (cudaEventElapsedTime(&ms, c0->ev0, c0->ev1) == cudaSuccess) ix->last_search_us.store(ms * 1000.f);
        }
        if (use_i8) {
            bool again = false;
            uint64_t cand = 0, resc = 0;
            for (size_t g = 0; g < G; ++g) {
                again = again || ctx[g]->i8_status[0] != 0;
                cand += ctx[g]->i8_status[1] & 0xFFFFFFFFull;
                resc += ctx[g]->i8_status[1] >> 32;
            }
            ix->byte_searches.fetch_add(1, std::memory_order_relaxed);
            ix->byte_candidates.store(cand, std::memory_order_relaxed);
            ix->byte_rescored.store(resc, std::memory_order_relaxed);
            if (again) { ix->byte_fallbacks.fetch_add(1, std::memory_order_relaxed); return -1; }
        }
        decode_keys(c0->out_pin, k, out_ids, out_dist, out_n);
        if (G == 1 && !use_i8 && bitmap == nullptr && pred == nullptr) scan_timing_report(ix->shards[0], c0, ix->shards[0]->n_built, k);
        return CSGPU_OK;
    };
    rc = body();
    if (rc == -1) { use_i8 = false; rc = body(); }
    if (rc) for (size_t g = 0; g < G; ++g) { DeviceGuard dg(ix->shards[g]->device); cudaStreamSynchronize(ctx[g]->stream); }
    release_all();
    return rc;
}

// nq (2..8) queries in ONE pass over every shard; results [nq][k] land in ctx0->out_pin.
// Bounds the multi-query scans in flight on an index (csgpu_index::MULTI_SLOTS): held from before the first launch until the
// results are back.
struct MultiSlot {
    const csgpu_index *ix;
    explicit MultiSlot(const csgpu_index *ix_) : ix(ix_)
    {
        std::unique_lock<std::mutex> lk(ix->multi_mu);
        ix->multi_cv.wait(lk, [&] { return ix->multi_running < csgpu_index::MULTI_SLOTS; });
        ++ix->multi_running;
    }
    ~MultiSlot()
    {
        { std::lock_guard<std::mutex> lk(ix->multi_mu); --ix->multi_running; }
        ix->multi_cv.notify_one();
    }
    MultiSlot(const MultiSlot &) = delete;
    MultiSlot &operator=(const MultiSlot &) = delete;
};

static int search_multi(const csgpu_index *ix, const float *q, uint32_t nq, uint32_t k,
                        uint32_t *out_ids, float *out_dist, uint32_t *out_n)
{
    MultiSlot slot(ix);
    const size_t G = ix->shards.size();
    std::vector<SearchCtx *> ctx(G, nullptr);
    int rc = CSGPU_OK;
    auto release_all = [&]() { for (size_t g = 0; g < G; ++g) if (ctx[g]) ctx_release(ix->shards[g], ctx[g]); };
    for (size_t g = 0; g < G && !rc; ++g) rc = ctx_acquire(ix, ix->shards[g], &ctx[g]);
    if (rc) { release_all(); return rc; }
    const size_t qbytes = (size_t)nq * ix->dim_pad * sizeof(float);
    auto body = [&]() -> int {
        for (size_t g = 0; g < G; ++g) {
            Shard *sh = ix->shards[g];
            SearchCtx *c = ctx[g];
            DeviceGuard dg(sh->device);
            memset(c->q_pin, 0, qbytes);
            for (uint32_t j = 0; j < nq; ++j) memcpy(c->q_pin + (size_t)j * ix->dim_pad, q + (size_t)j * ix->dim, (size_t)ix->dim * sizeof(float));
            CS_CUDA(cudaMemcpyAsync(c->q_dev, c->q_pin, qbytes, cudaMemcpyHostToDevice, c->stream));
            if (g == 0 && G > 1) CS_CUDA(cudaEventRecord(c->ev0, c->stream));   // one device: the kernel times itself (see search_one)
            uint64_t *dst = (G == 1) ? c->out_pin : c->out_dev;
            int r = enqueue_scan_multi(ix, sh, c, c->q_dev, nq, k, g == 0, dst, c->stream, nullptr, nullptr, 0, /*stamp_time=*/G == 1);
            if (r) return r;
        }
        SearchCtx *c0 = ctx[0];
        if (G > 1) {
            DeviceGuard dg(ix->shards[0]->device);
            const size_t per = (size_t)nq * k;
            for (size_t g = 1; g < G; ++g) {
                DeviceGuard dg2(ix->shards[g]->device);
                CS_CUDA(cudaMemcpyPeerAsync(c0->gather + g * per, ix->shards[0]->device, ctx[g]->out_dev, ix->shards[g]->device,
                                            per * sizeof(uint64_t), ctx[g]->stream));
                CS_CUDA(cudaEventRecord(ctx[g]->ev1, ctx[g]->stream));
            }
            for (size_t g = 1; g < G; ++g) CS_CUDA(cudaStreamWaitEvent(c0->stream, ctx[g]->ev1, 0));
            CS_CUDA(cudaMemcpyAsync(c0->gather, c0->out_dev, per * sizeof(uint64_t), cudaMemcpyDeviceToDevice, c0->stream));
            int r = enqueue_merge(c0->gather, (uint32_t)G, nq, k, c0->out_pin, c0->stream);
            if (r) return r;
        }
        {
            DeviceGuard dg(ix->shards[0]->device);
            if (G > 1) CS_CUDA(cudaEventRecord(c0->ev1, c0->stream));
            CS_CUDA(cudaStreamSynchronize(c0->stream));
            float ms = 0.f;
            if (G == 1) ix->last_search_us.store((float)c0->out_pin[(size_t)MAX_BATCH * CSGPU_MAX_K] * 1e-3f);
            else if (cudaEventElapsedTime(&ms, c0->ev0, c0->ev1) == cudaSuccess) ix->last_search_us.store(ms * 1000.f);
        }
        for (uint32_t j = 0; j < nq; ++j)
            decode_keys(c0->out_pin + (size_t)j * k, k, out_ids + (size_t)j * k, out_dist + (size_t)j * k, out_n ? out_n + j : nullptr);
        return CSGPU_OK;
    };
    rc = body();
    if (rc) for (size_t g = 0; g < G; ++g) { DeviceGuard dg(ix->shards[g]->device); cudaStreamSynchronize(ctx[g]->stream); }
    release_all();
    return rc;
}

// Rows per language id (32 bins; untagged rows = lang 31) and the largest file id over a shard's built rows.
__global__ void tag_stats_kernel(const uint32_t *__restrict__ tags, uint64_t n, unsigned long long *__restrict__ lang_rows, unsigned *__restrict__ max_file)
{
    __shared__ unsigned s_lang[32];
    __shared__ unsigned s_max;
    if (threadIdx.x < 32) s_lang[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_max = 0;
    __syncthreads();
    unsigned mx = 0;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t t = tags[r];
        atomicAdd(&s_lang[t >> 27], 1u);
        if (t != CSGPU_TAG_NONE) mx = max(mx, t & 0x07FFFFFFu);
    }
    atomicMax(&s_max, mx);
    __syncthreads();
    if (threadIdx.x < 32 && s_lang[threadIdx.x]) atomicAdd(&lang_rows[threadIdx.x], (unsigned long long)s_lang[threadIdx.x]);
    if (threadIdx.x == 0) atomicMax(max_file, s_max);
}

int tag_stats_refresh(csgpu_index *ix)
{
    memset(ix->lang_rows, 0, sizeof ix->lang_rows);
    ix->tagged_rows = 0;
    ix->max_file_id = 0;
    if (ix->dtype != CSGPU_DTYPE_F32) return CSGPU_OK;
    for (Shard *sh : ix->shards) {
        if (sh->n_built == 0 || sh->tags == nullptr) continue;
        DeviceGuard dg(sh->device);
        unsigned long long *d = nullptr;
        CS_CUDA(cudaMalloc(&d, 33 * sizeof(unsigned long long)));
        CS_CUDA(cudaMemsetAsync(d, 0, 33 * sizeof(unsigned long long), sh->stream));
        const uint32_t grid = (uint32_t)std::min<uint64_t>((sh->n_built + 255) / 256, (uint64_t)sh->sm_count * 8);
        tag_stats_kernel<<<grid, 256, 0, sh->stream>>>(sh->tags, sh->n_built, d, reinterpret_cast<unsigned *>(d + 32));
        count_launch();
        unsigned long long h[33];
        cudaError_t e = cudaMemcpyAsync(h, d, sizeof h, cudaMemcpyDeviceToHost, sh->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(sh->stream);
        cudaFree(d);
        if (e != cudaSuccess) return fail_cuda(e, "tag_stats_kernel", __FILE__, __LINE__);
        for (int l = 0; l < 32; ++l) ix->lang_rows[l] += h[l];
        ix->max_file_id = std::max<uint32_t>(ix->max_file_id, (uint32_t)(h[32] & 0xFFFFFFFFull));
    }
    for (int l = 0; l < 31; ++l) ix->tagged_rows += ix->lang_rows[l];
    return CSGPU_OK;
}

// Rough fraction of the rows a predicate lets through, from the build-time tag statistics: language shares are exact, files
// are taken as equally large and independent of the language. Good enough to pick a route; never affects a result.
static double predicate_density_estimate(const csgpu_index *ix, const csgpu_predicate_t *p)
{
    uint64_t rows = 0, pass = 0;
    for (int l = 0; l < 32; ++l) { rows += ix->lang_rows[l]; if ((p->lang_mask >> l) & 1u) pass += ix->lang_rows[l]; }
    if (rows == 0) return 1.0;
    double dens = (double)pass / (double)rows;
    const double n_files = (double)ix->max_file_id + 1.0;
    const double lo = (double)p->file_lo, hi = std::min((double)p->file_hi, n_files - 1.0);
    dens *= hi >= lo ? std::min(1.0, (hi - lo + 1.0) / n_files) : 0.0;
    if (p->file_bitmap) {
        const uint64_t nb = std::min<uint64_t>(p->n_file_bits, (uint64_t)ix->max_file_id + 1);
        uint64_t set = 0;
        for (uint64_t w = 0; w < (nb + 63) / 64; ++w) {
            uint64_t word = p->file_bitmap[w];
            if ((w + 1) * 64 > nb) word &= (nb % 64) ? ((1ull << (nb % 64)) - 1) : ~0ull;
            set += (uint64_t)__builtin_popcountll(word);
        }
        dens *= (double)set / n_files;
    }
    return dens;
}

// b (<= MAX_BATCH) variants of one user query: searched in ceil(b/8) passes, lists kept on the device, deduplicated
// by chunk id (best distance wins) and cut to the best k by dedup_variants_kernel. On a multi-device index every
// shard does that for its own rows (a chunk id lives on exactly one shard, so per-shard dedup + a k-way merge of the
// shards' lists on device 0 is the global dedup).
// pred != nullptr (csgpu_search_variants_tagged): every variant is searched under the row-tag predicate. Small corpora
// (<= 768 MB of rows per device: latency-bound) take the multi-query passes with the predicate applied to the results that
// beat a threshold; larger ones one filtered scan per variant, which never reads a masked row — all lists stay on the device.
static int search_variants(const csgpu_index *ix, const float *q, uint32_t b, uint32_t k,
                           uint32_t *out_ids, float *out_dist, uint32_t *out_n, const csgpu_predicate_t *pred = nullptr)
{
    MultiSlot slot(ix);
    const size_t G = ix->shards.size();
    std::vector<SearchCtx *> ctx(G, nullptr);
    int rc = CSGPU_OK;
    auto release_all = [&]() { for (size_t g = 0; g < G; ++g) if (ctx[g]) ctx_release(ix->shards[g], ctx[g]); };
    for (size_t g = 0; g < G && !rc; ++g) rc = ctx_acquire(ix, ix->shards[g], &ctx[g]);
    if (rc) { release_all(); return rc; }
    auto body = [&]() -> int {
        const size_t qbytes = (size_t)b * ix->dim_pad * sizeof(float);
        const uint32_t MQ = multi_scan_max_queries();
        const uint32_t total = b * k, npad = pow2_at_least(total, 64);
        const size_t smem = (size_t)npad * sizeof(uint64_t);
        const double pred_density = pred ? predicate_density_estimate(ix, pred) : 1.0;
        const uint64_t *file_bitmap = pred ? pred->file_bitmap : nullptr;
        const uint64_t n_file_bits = file_bitmap ? pred->n_file_bits : 0;
        const size_t bm_words = file_bitmap ? (size_t)((n_file_bits + 63) / 64) : 0;
        for (size_t g = 0; g < G; ++g) {
            Shard *sh = ix->shards[g];
            SearchCtx *c = ctx[g];
            DeviceGuard dg(sh->device);
            // under a predicate: a multi-query pass reads every row once for <= 16 variants (~1.25 x a dense single scan for <= 8
            // of them, ~2.2 x for 9-16), a filtered scan per variant reads only what passes — b x density dense scans. Small
            // corpora are latency-bound: always the one launch.
            const bool small_corpus = (uint64_t)sh->n_built * ix->dim4 * sizeof(float4) <= (768ull << 20);
            const bool dense_enough = pred != nullptr && (double)b * pred_density > (b <= 8 ? 1.25 : 2.2);
            const bool multi_ok = multi_scan_supported(ix->dim4, k) && (pred == nullptr || small_corpus || dense_enough);
            memset(c->q_pin, 0, qbytes);
            for (uint32_t j = 0; j < b; ++j) memcpy(c->q_pin + (size_t)j * ix->dim_pad, q + (size_t)j * ix->dim, (size_t)ix->dim * sizeof(float));
            CS_CUDA(cudaMemcpyAsync(c->q_dev, c->q_pin, qbytes, cudaMemcpyHostToDevice, c->stream));
            const uint64_t *bm_dev = nullptr;
            if (file_bitmap) {   // as in search_one: n_file_bits == 0 still needs a (never read) non-null pointer — it excludes every file
                if (c->bitmap_dev == nullptr || c->bitmap_cap < bm_words) {
                    CS_CUDA(cudaStreamSynchronize(c->stream));
                    cudaFree(c->bitmap_dev); c->bitmap_dev = nullptr; c->bitmap_cap = 0;
                    CS_CUDA(cudaMalloc(&c->bitmap_dev, std::max<size_t>(bm_words, 1) * sizeof(uint64_t)));
                    c->bitmap_cap = std::max<size_t>(bm_words, 1);
                }
                if (bm_words) CS_CUDA(cudaMemcpyAsync(c->bitmap_dev, file_bitmap, bm_words * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
                bm_dev = c->bitmap_dev;
            }
            if (g == 0) CS_CUDA(cudaEventRecord(c->ev0, c->stream));
            for (uint32_t j = 0; j < b;) {
                const uint32_t nq = std::min(MQ, b - j);
                int r;
                if (multi_ok && nq >= 2) {
                    r = enqueue_scan_multi(ix, sh, c, c->q_dev + (size_t)j * ix->dim_pad, nq, k, g == 0, c->out_dev + (size_t)j * k, c->stream,
                                           pred, bm_dev, n_file_bits);
                    j += nq;
                } else {
                    r = enqueue_scan(ix, sh, c, c->q_dev + (size_t)j * ix->dim_pad, k, bm_dev, n_file_bits, g == 0, c->out_dev + (size_t)j * k, c->stream,
                                     nullptr, 0, pred);
                    j += 1;
                }
                if (r) return r;
            }
            // always the ceiling (16 variants x k = 1024 keys): per-launch values from concurrent callers must not undercut each other
            if (total <= 8192) {   // hash-table dedup: only the distinct ids are sorted (scan.cuh)
                const uint32_t tbl = pow2_at_least(2 * total, 64), cpad = pow2_at_least(total, 32);
                CS_CUDA(cudaFuncSetAttribute(dedup_variants_hash_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)((size_t)(16384 + 8192) * sizeof(uint64_t))));   // the ceiling, never this launch's size
                dedup_variants_hash_kernel<<<1, SCAN_THREADS, (size_t)(tbl + cpad) * sizeof(uint64_t), c->stream>>>(
                    c->out_dev, total, tbl, cpad, k, G == 1 ? c->out_pin : c->cand);
            } else {
            CS_CUDA(cudaFuncSetAttribute(dedup_variants_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)((size_t)MAX_BATCH * CSGPU_MAX_K * sizeof(uint64_t))));
            // one device: straight into the pinned result; several: the shard's deduplicated list parks in its scratch
            dedup_variants_kernel<<<1, SCAN_THREADS, smem, c->stream>>>(c->out_dev, total, npad, k, G == 1 ? c->out_pin : c->cand);
            }
            count_launch();
            CS_CUDA(cudaGetLastError());
        }
        SearchCtx *c0 = ctx[0];
        if (G > 1) {
            DeviceGuard dg(ix->shards[0]->device);
            for (size_t g = 1; g < G; ++g) {
                DeviceGuard dg2(ix->shards[g]->device);
                CS_CUDA(cudaMemcpyPeerAsync(c0->gather + g * k, ix->shards[0]->device, ctx[g]->cand, ix->shards[g]->device,
                                            (size_t)k * sizeof(uint64_t), ctx[g]->stream));
                CS_CUDA(cudaEventRecord(ctx[g]->ev1, ctx[g]->stream));
            }
            for (size_t g = 1; g < G; ++g) CS_CUDA(cudaStreamWaitEvent(c0->stream, ctx[g]->ev1, 0));
            CS_CUDA(cudaMemcpyAsync(c0->gather, c0->cand, (size_t)k * sizeof(uint64_t), cudaMemcpyDeviceToDevice, c0->stream));
            int r = enqueue_merge(c0->gather, (uint32_t)G, 1, k, c0->out_pin, c0->stream);
            if (r) return r;
        }
        DeviceGuard dg(ix->shards[0]->device);
        CS_CUDA(cudaEventRecord(c0->ev1, c0->stream));
        CS_CUDA(cudaStreamSynchronize(c0->stream));
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, c0->ev0, c0->ev1) == cudaSuccess) ix->last_search_us.store(ms * 1000.f);
        decode_keys(c0->out_pin, k, out_ids, out_dist, out_n);
        return CSGPU_OK;
    };
    rc = body();
    if (rc) for (size_t g = 0; g < G; ++g) { DeviceGuard dg(ix->shards[g]->device); cudaStreamSynchronize(ctx[g]->stream); }
    release_all();
    return rc;
}

// Same contract with the tensor prefilter on: the variants are ONE batch on the tensor cores (bf16 shadow as a filter +
// exact fp32 rescoring: every list bit-identical to csgpu_search), and the few thousand result entries are
// deduplicated on the host — the same rule as dedup_variants_kernel / src/search/mod.rs:513-590: per chunk id the
// smallest distance, then ascending (distance, id), first k. 9 variants x top-200 over 10M rows: one 1.6 ms batch
// instead of two multi-query passes (5.5 ms).
static int search_variants_prefiltered(const csgpu_index *ix, const float *q, uint32_t b, uint32_t k,
                                       uint32_t *out_ids, float *out_dist, uint32_t *out_n)
{
    std::vector<uint32_t> ids((size_t)b * k), ns(b, 0);
    std::vector<float> dd((size_t)b * k);
    std::vector<uint32_t> zero_q;
    int rc = batch_search(ix, q, b, k, ids.data(), dd.data(), ns.data(), &zero_q);
    for (size_t z = 0; z < zero_q.size() && !rc; ++z) {   // zero-norm variants: distance 0.0 everywhere (scan kernel)
        const uint32_t j = zero_q[z];
        rc = search_one(ix, q + (size_t)j * ix->dim, k, nullptr, 0, ids.data() + (size_t)j * k, dd.data() + (size_t)j * k, &ns[j]);
    }
    if (rc) return rc;
    std::vector<uint64_t> by_id;
    by_id.reserve((size_t)b * k);
    for (uint32_t j = 0; j < b; ++j)
        for (uint32_t i = 0; i < ns[j]; ++i) {
            uint32_t bits;
            memcpy(&bits, &dd[(size_t)j * k + i], sizeof bits);
            by_id.push_back(((uint64_t)ids[(size_t)j * k + i] << 32) | okey_from_bits(bits));
        }
    std::sort(by_id.begin(), by_id.end());   // (id, distance): the first entry of every id run is its best
    std::vector<uint64_t> keys;
    keys.reserve(by_id.size());
    for (size_t t = 0; t < by_id.size(); ++t)
        if (t == 0 || (by_id[t] >> 32) != (by_id[t - 1] >> 32)) keys.push_back((by_id[t] << 32) | (by_id[t] >> 32));
    std::sort(keys.begin(), keys.end());     // (distance, id)
    keys.resize(std::min<size_t>(keys.size(), k), KEY_EMPTY);
    keys.resize(k, KEY_EMPTY);
    decode_keys(keys.data(), k, out_ids, out_dist, out_n);
    return CSGPU_OK;
}

// One query through the micro-batcher (see Coalescer in index.h).
static int search_coalesced(const csgpu_index *ix, const float *q, uint32_t k, uint32_t *out_ids, float *out_dist, uint32_t *out_n)
{
    Coalescer &co = ix->coalescer;
    PendingSearch me{q, k, out_ids, out_dist, out_n};
    std::unique_lock<std::mutex> lk(co.mu);
    co.queue.push_back(&me);
    // group size: one multi-query pass (16), or — where the group becomes a tensor-core batch (tf32 filter of an fp32 index,
    // gemm_tf32.cuh, or the tensor prefilter) — one 128-query block, which costs about as much as a single query at 10M rows
    const uint32_t MQ = (ix->dtype == CSGPU_DTYPE_F32 && (batch_tf32_route(ix) || ix->tensor_prefilter)) ? 128u : multi_scan_max_queries();
    while (!me.done) {
        if (co.leader_active) { co.cv.wait(lk); continue; }
        co.leader_active = true;
        if (co.window_us && co.queue.size() < MQ) {
            lk.unlock();
            std::this_thread::sleep_for(std::chrono::microseconds(co.window_us));
            lk.lock();
        }
        // one pass: up to MQ queued requests that share the k of the oldest one
        std::vector<PendingSearch *> batch;
        const uint32_t bk = co.queue.front()->k;
        for (size_t i = 0; i < co.queue.size() && batch.size() < MQ;) {
            if (co.queue[i]->k == bk) { batch.push_back(co.queue[i]); co.queue.erase(co.queue.begin() + i); }
            else ++i;
        }
        lk.unlock();
        int rc;
        const uint32_t nb = (uint32_t)batch.size();
        if (nb == 1) {
            PendingSearch *r = batch[0];
            rc = search_one(ix, r->q, bk, nullptr, 0, r->out_ids, r->out_dist, r->out_n);
        } else {
            // the group is a small batch: csgpu_search_batch picks the route (one multi-query pass; with the byte
            // prefilter on, <= 4 queries one after the other through the int8 kernel; with the tensor prefilter on, one
            // tensor-core batch) — every route returns what csgpu_search would
            std::vector<float> qs((size_t)nb * ix->dim);
            std::vector<uint32_t> ids((size_t)nb * bk), ns(nb);
            std::vector<float> dd((size_t)nb * bk);
            for (uint32_t j = 0; j < nb; ++j) memcpy(qs.data() + (size_t)j * ix->dim, batch[j]->q, (size_t)ix->dim * sizeof(float));
            rc = csgpu_search_batch(ix, qs.data(), ix->dim, nb, bk, ids.data(), dd.data(), ns.data());
            for (uint32_t j = 0; j < nb && !rc; ++j) {
                memcpy(batch[j]->out_ids, ids.data() + (size_t)j * bk, (size_t)ns[j] * sizeof(uint32_t));
                memcpy(batch[j]->out_dist, dd.data() + (size_t)j * bk, (size_t)ns[j] * sizeof(float));
                if (batch[j]->out_n) *batch[j]->out_n = ns[j];
            }
        }
        co.passes.fetch_add(1, std::memory_order_relaxed);
        co.queries.fetch_add(nb, std::memory_order_relaxed);
        const std::string err = rc ? t_error : std::string();
        lk.lock();
        for (PendingSearch *r : batch) { r->rc = rc; r->err = err; r->done = true; }
        co.leader_active = false;
        co.cv.notify_all();
    }
    if (me.rc) t_error = me.err;   // the error text is thread-local: hand the leader's message to this caller
    return me.rc;
}

}  // namespace csgpu

using namespace csgpu;

// =========================================================================================
// C ABI
// =========================================================================================
extern "C" {

uint32_t csgpu_abi_version(void) { return CSGPU_ABI_VERSION; }
const char *csgpu_last_error(void) { return t_error.c_str(); }
uint64_t csgpu_kernel_launches(void) { return g_kernel_launches.load(); }

int csgpu_create(csgpu_index **out, uint32_t dim, uint32_t dtype, const int32_t *devices, uint32_t n_devices)
{
    if (!out) return fail(CSGPU_ERR_ARG, "out is null");
    *out = nullptr;
    if (dim == 0 || dim > CSGPU_MAX_DIM) return fail(CSGPU_ERR_ARG, "dim must be in [1, 4096]");
    if (((dim + 3) / 4 + 31) / 32 > 8) return fail(CSGPU_ERR_ARG, "dim > 1024 is not supported by the scan kernels yet");
    if (dtype != CSGPU_DTYPE_F32 && dtype != CSGPU_DTYPE_BF16) return fail(CSGPU_ERR_ARG, "dtype must be CSGPU_DTYPE_F32 or CSGPU_DTYPE_BF16");
    if (dtype == CSGPU_DTYPE_BF16 && !bf16_dim_supported(dim)) return fail(CSGPU_ERR_ARG, "bf16 index needs dim % 64 == 0 and 64 <= dim <= 512");
    if (n_devices == 0) n_devices = 1;
    if (n_devices > 8) return fail(CSGPU_ERR_ARG, "n_devices must be <= 8");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(CSGPU_ERR_CUDA, std::string("no usable CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
    }
    csgpu_index *ix = new csgpu_index();
    ix->dim = dim;
    ix->dim_pad = (dim + 3) & ~3u;
    ix->dim4 = ix->dim_pad / 4;
    ix->dtype = dtype;
    for (uint32_t g = 0; g < n_devices; ++g) {
        const int dev = devices ? devices[g] : (int)g;
        if (dev < 0 || dev >= count) { csgpu_destroy(ix); return fail(CSGPU_ERR_ARG, "device ordinal out of range"); }
        cudaDeviceProp prop;
        if ((e = cudaGetDeviceProperties(&prop, dev)) != cudaSuccess) { csgpu_destroy(ix); return fail_cuda(e, "cudaGetDeviceProperties", __FILE__, __LINE__); }
        if (prop.major != 10) {
            csgpu_destroy(ix);
            return fail(CSGPU_ERR_CUDA, std::string("device is not sm_100 (Blackwell B200); this library ships sm_100a code only: ") + prop.name);
        }
        Shard *sh = new Shard();
        sh->device = dev;
        sh->sm_count = prop.multiProcessorCount;
        ix->shards.push_back(sh);
        DeviceGuard dg(dev);
        if ((e = cudaStreamCreateWithFlags(&sh->stream, cudaStreamNonBlocking)) != cudaSuccess) { csgpu_destroy(ix); return fail_cuda(e, "cudaStreamCreate", __FILE__, __LINE__); }
    }
    // peer access for the in-process multi-GPU gather
    for (size_t a = 0; a < ix->shards.size(); ++a)
        for (size_t b = 0; b < ix->shards.size(); ++b) {
            if (a == b) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, ix->shards[a]->device, ix->shards[b]->device);
            if (can) { DeviceGuard dg(ix->shards[a]->device); cudaError_t pe = cudaDeviceEnablePeerAccess(ix->shards[b]->device, 0); if (pe != cudaSuccess) cudaGetLastError(); }
        }
    *out = ix;
    return CSGPU_OK;
}

static void exchange_free(csgpu_index *ix);
static int exchange_healthy(const csgpu_index *ix);

void csgpu_destroy(csgpu_index *ix)
{
    if (!ix) return;
    exchange_free(ix);
    for (GroupCtx *g : ix->all_groups) group_destroy(g);
    for (Shard *sh : ix->shards) {
        DeviceGuard dg(sh->device);
        for (SearchCtx *c : sh->all_ctx) ctx_destroy(c);
        ctx_destroy(sh->dev_ctx);
        batch_free_ctx(sh);
        i8_free_shard(sh);
        cudaFree(sh->rows_bf16); cudaFree(sh->stage); cudaFree(sh->shadow_bf16);
        cudaFree(sh->rows); cudaFree(sh->ids); cudaFree(sh->status); cudaFree(sh->tags);
        if (sh->stream) cudaStreamDestroy(sh->stream);
        delete sh;
    }
    if (ix->zero_ids_dev && !ix->shards.empty()) { cudaFree(ix->zero_ids_dev); }
    delete ix;
}

int csgpu_reserve(csgpu_index *ix, uint64_t total_rows)
{
    if (!ix) return fail(CSGPU_ERR_ARG, "null index");
    const uint64_t G = ix->shards.size();
    for (Shard *sh : ix->shards) {
        int rc = shard_reserve(ix, sh, (total_rows + G - 1) / G);
        if (rc) return rc;
    }
    return CSGPU_OK;
}

static int append_rows(csgpu_index *ix, const float *rows, const uint32_t *ids, const uint32_t *tags, uint64_t n)
{
    if (!ix) return fail(CSGPU_ERR_ARG, "null index");
    if (n == 0) return CSGPU_OK;
    if (!rows || !ids) return fail(CSGPU_ERR_ARG, "rows/ids is null");
    // replace semantics (LMDB put): an id that is already present dies; within the batch the last wins
    {
        uint64_t killed = 0;
        bool any_rows = !ix->zero_ids.empty();
        for (Shard *sh : ix->shards) any_rows = any_rows || sh->n_total;
        if (any_rows) { int rc = kill_ids(ix, ids, n, &killed); if (rc) return rc; }
    }
    std::vector<uint8_t> dup;  // rows of this batch superseded by a later row with the same id
    {
        bool increasing = true;
        for (uint64_t i = 1; i < n && increasing; ++i) increasing = ids[i] > ids[i - 1];
        if (!increasing) {
            std::vector<uint64_t> order(n);
            std::iota(order.begin(), order.end(), 0);
            std::stable_sort(order.begin(), order.end(), [&](uint64_t a, uint64_t b) { return ids[a] < ids[b]; });
            for (uint64_t i = 0; i + 1 < n; ++i)
                if (ids[order[i]] == ids[order[i + 1]]) { if (dup.empty()) dup.assign(n, 0); dup[order[i]] = ROW_DEAD; }
        }
    }
    const uint64_t G = ix->shards.size();
    // big batches are split evenly; small ones go to the least-loaded shard
    std::vector<std::pair<Shard *, std::pair<uint64_t, uint64_t>>> parts;
    if (G > 1 && n >= 4096 * G) {
        for (uint64_t g = 0; g < G; ++g) parts.push_back({ix->shards[g], {n * g / G, n * (g + 1) / G}});
    } else {
        parts.push_back({least_loaded(ix), {0, n}});
    }
    std::vector<float> padded;
    for (auto &p : parts) {
        Shard *sh = p.first;
        const uint64_t a = p.second.first, b = p.second.second, m = b - a;
        if (!m) continue;
        int rc = shard_reserve(ix, sh, sh->n_total + m);
        if (rc) return rc;
        const bool bf16 = ix->dtype == CSGPU_DTYPE_BF16;
        if (bf16 && (rc = bf16_reserve_stage(ix, sh, sh->n_total - sh->n_built + m))) return rc;
        DeviceGuard dg(sh->device);
        float *dst = bf16 ? sh->stage + (sh->n_total - sh->n_built) * ix->dim_pad : sh->rows + sh->n_total * ix->dim_pad;
        if (ix->dim_pad == ix->dim) {
            CS_CUDA(cudaMemcpyAsync(dst, rows + a * ix->dim, m * (size_t)ix->dim * sizeof(float), cudaMemcpyHostToDevice, sh->stream));
        } else {
            CS_CUDA(cudaMemsetAsync(dst, 0, m * (size_t)ix->dim_pad * sizeof(float), sh->stream));
            CS_CUDA(cudaMemcpy2DAsync(dst, (size_t)ix->dim_pad * sizeof(float), rows + a * ix->dim, (size_t)ix->dim * sizeof(float),
                                      (size_t)ix->dim * sizeof(float), m, cudaMemcpyHostToDevice, sh->stream));
        }
        CS_CUDA(cudaMemcpyAsync(sh->ids + sh->n_total, ids + a, m * sizeof(uint32_t), cudaMemcpyHostToDevice, sh->stream));
        if (tags) CS_CUDA(cudaMemcpyAsync(sh->tags + sh->n_total, tags + a, m * sizeof(uint32_t), cudaMemcpyHostToDevice, sh->stream));
        else CS_CUDA(cudaMemsetAsync(sh->tags + sh->n_total, 0xFF, m * sizeof(uint32_t), sh->stream));
        if (!dup.empty()) CS_CUDA(cudaMemcpyAsync(sh->status + sh->n_total, dup.data() + a, m, cudaMemcpyHostToDevice, sh->stream));
        else CS_CUDA(cudaMemsetAsync(sh->status + sh->n_total, 0, m, sh->stream));
        CS_CUDA(cudaStreamSynchronize(sh->stream));
        sh->n_total += m;
    }
    ix->built = false;
    return CSGPU_OK;
}

int csgpu_append(csgpu_index *ix, const float *rows, const uint32_t *ids, uint64_t n)
{
    return append_rows(ix, rows, ids, nullptr, n);
}

int csgpu_append_tagged(csgpu_index *ix, const float *rows, const uint32_t *ids, const uint32_t *tags, uint64_t n)
{
    if (n && !tags) return fail(CSGPU_ERR_ARG, "tags is null");
    return append_rows(ix, rows, ids, tags, n);
}

static int append_synthetic(csgpu_index *ix, uint64_t seed, uint64_t first_row, uint64_t n, uint32_t id_base, bool tagged)
{
    if (!ix) return fail(CSGPU_ERR_ARG, "null index");
    if (n == 0) return CSGPU_OK;
    if (ix->dim != ix->dim_pad) return fail(CSGPU_ERR_ARG, "synthetic rows need dim % 4 == 0");
    const uint64_t G = ix->shards.size();
    for (uint64_t g = 0; g < G; ++g) {
        Shard *sh = ix->shards[g];
        const uint64_t a = n * g / G, b = n * (g + 1) / G, m = b - a;
        if (!m) continue;
        int rc = shard_reserve(ix, sh, sh->n_total + m);
        if (rc) return rc;
        const bool bf16 = ix->dtype == CSGPU_DTYPE_BF16;
        if (bf16 && (rc = bf16_reserve_stage(ix, sh, sh->n_total - sh->n_built + m))) return rc;
        DeviceGuard dg(sh->device);
        const uint64_t total = m * ix->dim4;
        const uint32_t grid = (uint32_t)std::min<uint64_t>((total + 255) / 256, (uint64_t)sh->sm_count * 16);
        float4 *synth_dst = bf16 ? reinterpret_cast<float4 *>(sh->stage) + (sh->n_total - sh->n_built) * ix->dim4
                                 : reinterpret_cast<float4 *>(sh->rows) + sh->n_total * ix->dim4;
        synth_rows_kernel<<<grid, 256, 0, sh->stream>>>(synth_dst,
                                                        sh->ids + sh->n_total, seed, first_row + a, m, ix->dim4, id_base);
        count_launch();
        CS_CUDA(cudaGetLastError());
        if (tagged) {
            synth_tags_kernel<<<(uint32_t)std::min<uint64_t>((m + 255) / 256, (uint64_t)sh->sm_count * 16), 256, 0, sh->stream>>>(sh->tags + sh->n_total, first_row + a, m);
            count_launch();
            CS_CUDA(cudaGetLastError());
        } else {
            CS_CUDA(cudaMemsetAsync(sh->tags + sh->n_total, 0xFF, m * sizeof(uint32_t), sh->stream));
        }
        CS_CUDA(cudaMemsetAsync(sh->status + sh->n_total, 0, m, sh->stream));
        CS_CUDA(cudaStreamSynchronize(sh->stream));
        sh->n_total += m;
    }
    ix->built = false;
    return CSGPU_OK;
}

int csgpu_append_synthetic(csgpu_index *ix, uint64_t seed, uint64_t first_row, uint64_t n, uint32_t id_base)
{
    return append_synthetic(ix, seed, first_row, n, id_base, false);
}

int csgpu_append_synthetic_tagged(csgpu_index *ix, uint64_t seed, uint64_t first_row, uint64_t n, uint32_t id_base)
{
    return append_synthetic(ix, seed, first_row, n, id_base, true);
}

int csgpu_synth_rows_host(const csgpu_index *ix, uint64_t seed, uint64_t first_row, uint64_t n, float *out_rows)
{
    if (!ix || !out_rows) return fail(CSGPU_ERR_ARG, "null argument");
    if (ix->dim != ix->dim_pad) return fail(CSGPU_ERR_ARG, "synthetic rows need dim % 4 == 0");
    if (n == 0) return CSGPU_OK;
    Shard *sh = ix->shards[0];
    DeviceGuard dg(sh->device);
    float *tmp = nullptr;
    CS_CUDA(cudaMalloc(&tmp, n * (size_t)ix->dim_pad * sizeof(float)));
    const uint64_t total = n * ix->dim4;
    const uint32_t grid = (uint32_t)std::min<uint64_t>((total + 255) / 256, (uint64_t)sh->sm_count * 16);
    synth_rows_kernel<<<grid, 256, 0, sh->stream>>>(reinterpret_cast<float4 *>(tmp), nullptr, seed, first_row, n, ix->dim4, 0);
    count_launch();
    cudaError_t e = cudaMemcpyAsync(out_rows, tmp, n * (size_t)ix->dim_pad * sizeof(float), cudaMemcpyDeviceToHost, sh->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(sh->stream);
    cudaFree(tmp);
    if (e != cudaSuccess) return fail_cuda(e, "synth_rows_host", __FILE__, __LINE__);
    return CSGPU_OK;
}

int csgpu_remove(csgpu_index *ix, const uint32_t *ids, uint64_t n, uint64_t *n_removed)
{
    if (!ix) return fail(CSGPU_ERR_ARG, "null index");
    if (n_removed) *n_removed = 0;
    if (n == 0) return CSGPU_OK;
    if (!ids) return fail(CSGPU_ERR_ARG, "ids is null");
    uint64_t killed = 0;
    int rc = kill_ids(ix, ids, n, &killed);
    if (rc) return rc;
    if (n_removed) *n_removed = killed;
    ix->tombstones += killed;
    ix->built = false;  // store.rs:605-607 — even a no-op delete leaves the index needing a rebuild
    return CSGPU_OK;
}

int csgpu_build(csgpu_index *ix)
{
    if (!ix) return fail(CSGPU_ERR_ARG, "null index");
    std::vector<std::pair<uint32_t, uint32_t>> new_zero;
    for (Shard *sh : ix->shards) {
        int rc = shard_build(ix, sh, new_zero);
        if (rc) return rc;
    }
    if (!new_zero.empty()) {
        ix->zero_tags.resize(ix->zero_ids.size(), CSGPU_TAG_NONE);
        for (size_t i = 0; i < ix->zero_ids.size(); ++i) new_zero.push_back({ix->zero_ids[i], ix->zero_tags[i]});
        std::sort(new_zero.begin(), new_zero.end());
        new_zero.erase(std::unique(new_zero.begin(), new_zero.end(), [](const auto &x, const auto &y) { return x.first == y.first; }), new_zero.end());
        ix->zero_ids.clear(); ix->zero_tags.clear();
        for (const auto &z : new_zero) { ix->zero_ids.push_back(z.first); ix->zero_tags.push_back(z.second); }
    }
    int rc = upload_zero_ids(ix);
    if (rc) return rc;
    if ((rc = tag_stats_refresh(ix))) return rc;
    for (Shard *sh : ix->shards) if ((rc = batch_after_build(ix, sh))) return rc;
    for (Shard *sh : ix->shards) if ((rc = i8_refresh(ix, sh))) return rc;
    if (ix->byte_prefilter && (rc = warm_i8_contexts(ix))) return rc;
    ix->tombstones = 0;
    ix->built = true;
    return CSGPU_OK;
}

int csgpu_clear(csgpu_index *ix)
{
    if (!ix) return fail(CSGPU_ERR_ARG, "null index");
    for (Shard *sh : ix->shards) {
        DeviceGuard dg(sh->device);
        cudaFree(sh->rows); cudaFree(sh->ids); cudaFree(sh->status); cudaFree(sh->tags);
        cudaFree(sh->rows_bf16); cudaFree(sh->stage); cudaFree(sh->shadow_bf16);
        sh->shadow_bf16 = nullptr; sh->shadow_valid = false; sh->shadow_rows = 0;
        i8_free_shard(sh);
        sh->rows_bf16 = nullptr; sh->stage = nullptr; sh->stage_cap = 0; sh->map_valid = false;
        sh->rows = nullptr; sh->ids = nullptr; sh->status = nullptr; sh->tags = nullptr;
        sh->n_built = sh->n_total = sh->cap = 0;
    }
    ix->zero_ids.clear();
    ix->zero_tags.clear();
    upload_zero_ids(ix);
    ix->nonfinite_rows = 0;
    ix->tombstones = 0;
    ix->built = false;
    return CSGPU_OK;
}

int csgpu_search(const csgpu_index *ix, const float *q, uint32_t q_len, uint32_t k,
                 uint32_t *out_ids, float *out_dist, uint32_t *out_n)
{
    if (out_n) *out_n = 0;
    int rc = check_search_args(ix, q, q_len, k);
    if (rc) return rc;
    if (!all_finite(q, q_len)) return fail(CSGPU_ERR_ARG, "query contains NaN/Inf");
    if (k == 0) return CSGPU_OK;
    if (ix->dtype == CSGPU_DTYPE_BF16) return batch_search(ix, q, 1, k, out_ids, out_dist, out_n, nullptr);
    if (ix->coalescer.enabled.load(std::memory_order_relaxed) && multi_scan_supported(ix->dim4, k))
        return search_coalesced(ix, q, k, out_ids, out_dist, out_n);
    return search_one(ix, q, k, nullptr, 0, out_ids, out_dist, out_n);
}

int csgpu_set_coalescing(csgpu_index *ix, uint32_t enabled, uint32_t window_us)
{
    if (!ix) return fail(CSGPU_ERR_ARG, "null index");
    std::lock_guard<std::mutex> lk(ix->coalescer.mu);
    ix->coalescer.window_us = window_us;
    ix->coalescer.enabled.store(enabled ? 1 : 0);
    return CSGPU_OK;
}

int csgpu_search_filtered(const csgpu_index *ix, const float *q, uint32_t q_len, uint32_t k,
                          const uint64_t *id_bitmap, uint64_t n_bits,
                          uint32_t *out_ids, float *out_dist, uint32_t *out_n)
{
    if (out_n) *out_n = 0;
    int rc = check_search_args(ix, q, q_len, k);
    if (rc) return rc;
    if (!id_bitmap && n_bits) return fail(CSGPU_ERR_ARG, "id_bitmap is null");
    if (ix->dtype == CSGPU_DTYPE_BF16) return fail(CSGPU_ERR_ARG, "filtered search is not implemented for the bf16 index yet");
    if (!all_finite(q, q_len)) return fail(CSGPU_ERR_ARG, "query contains NaN/Inf");
    if (k == 0) return CSGPU_OK;
    static const uint64_t empty_word = 0;
    return search_one(ix, q, k, id_bitmap ? id_bitmap : &empty_word, n_bits, out_ids, out_dist, out_n);
}

int csgpu_search_tagged(const csgpu_index *ix, const float *q, uint32_t q_len, uint32_t k, const csgpu_predicate_t *pred,
                        uint32_t *out_ids, float *out_dist, uint32_t *out_n)
{
    if (out_n) *out_n = 0;
    int rc = check_search_args(ix, q, q_len, k);
    if (rc) return rc;
    if (!pred) return fail(CSGPU_ERR_ARG, "pred is null");
    if (!pred->file_bitmap && pred->n_file_bits) return fail(CSGPU_ERR_ARG, "file_bitmap is null");
    if (ix->dtype == CSGPU_DTYPE_BF16) return fail(CSGPU_ERR_ARG, "filtered search is not implemented for the bf16 index yet");
    if (!all_finite(q, q_len)) return fail(CSGPU_ERR_ARG, "query contains NaN/Inf");
    if (k == 0) return CSGPU_OK;
    csgpu_predicate_t p = *pred;
    static const uint64_t empty_word = 0;
    if (p.file_bitmap && p.n_file_bits == 0) p.file_bitmap = &empty_word;   // an empty bitmap allows nothing
    return search_one(ix, q, k, nullptr, 0, out_ids, out_dist, out_n, &p);
}

// tags[i] of every live row whose id is in `ids` (sorted copy on the device; one thread per row)
__global__ void lookup_tags_kernel(const uint32_t *__restrict__ ids, const uint32_t *__restrict__ tags, uint64_t n_rows,
                                   const uint32_t *__restrict__ want_sorted, const uint32_t *__restrict__ want_pos, uint32_t n_want,
                                   uint32_t *__restrict__ out)
{
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t id = ids[r];
        uint32_t lo = 0, hi = n_want;
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (want_sorted[mid] < id) lo = mid + 1; else hi = mid; }
        for (; lo < n_want && want_sorted[lo] == id; ++lo) out[want_pos[lo]] = tags[r];
    }
}

int csgpu_get_tags(const csgpu_index *ix, const uint32_t *ids, uint64_t n, uint32_t *out_tags)
{
    if (!ix) return fail(CSGPU_ERR_ARG, "null index");
    if (n == 0) return CSGPU_OK;
    if (!ids || !out_tags) return fail(CSGPU_ERR_ARG, "null argument");
    if (!ix->built) return fail(CSGPU_ERR_NOT_BUILT, "Index not built. Call build_index() after inserting chunks.");
    if (n > 0xFFFFFFFFull) return fail(CSGPU_ERR_ARG, "too many ids");
    std::vector<uint32_t> order(n), sorted(n);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return ids[a] < ids[b]; });
    for (uint64_t i = 0; i < n; ++i) sorted[i] = ids[order[i]];
    std::vector<uint32_t> result(n, CSGPU_TAG_NONE);
    for (Shard *sh : ix->shards) {
        if (!sh->n_built) continue;
        DeviceGuard dg(sh->device);
        uint32_t *d = nullptr;
        CS_CUDA(cudaMalloc(&d, 3 * n * sizeof(uint32_t)));
        CS_CUDA(cudaMemcpyAsync(d, sorted.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, sh->stream));
        CS_CUDA(cudaMemcpyAsync(d + n, order.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, sh->stream));
        CS_CUDA(cudaMemcpyAsync(d + 2 * n, result.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, sh->stream));
        const uint32_t grid = (uint32_t)std::min<uint64_t>((sh->n_built + 255) / 256, (uint64_t)sh->sm_count * 8);
        lookup_tags_kernel<<<grid, 256, 0, sh->stream>>>(sh->ids, sh->tags, sh->n_built, d, d + n, (uint32_t)n, d + 2 * n);
        count_launch();
        cudaError_t e = cudaMemcpyAsync(result.data(), d + 2 * n, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, sh->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(sh->stream);
        cudaFree(d);
        if (e != cudaSuccess) return fail_cuda(e, "csgpu_get_tags", __FILE__, __LINE__);
    }
    for (uint64_t i = 0; i < n; ++i) {   // zero-norm rows live on the host list
        auto it = std::lower_bound(ix->zero_ids.begin(), ix->zero_ids.end(), ids[i]);
        if (it != ix->zero_ids.end() && *it == ids[i]) {
            const size_t z = it - ix->zero_ids.begin();
            result[i] = z < ix->zero_tags.size() ? ix->zero_tags[z] : CSGPU_TAG_NONE;
        }
    }
    memcpy(out_tags, result.data(), n * sizeof(uint32_t));
    return CSGPU_OK;
}

// fp32 index, tf32 route available: does the tensor-core filter (+ exact rescoring) beat the scan kernels for b queries?
// A cost model fitted to measurements on one B200 (profiles/r02_tf32_probe_*.txt, r02_multi_tail_ab.txt), times in us, sizes
// in MB per device:
//   tensor-core route  ~205 us fixed (260 for k > 32: query prep, 3-5 phases of contraction + select, one host round trip)
//                      + one pass over the rows at ~6.4 TB/s per <= 128 queries; above, the tensor pipe is the bound
//                      (1024 queries: 12.0 ms at 15.36 GB);
//   multi-query scan   one launch per pass of <= 16 queries: ~45 us + ~8 us per query of tail; <= 8 queries stream at
//                      ~6.3 TB/s, 9-16 at ~3.7 TB/s (issue-bound);
//   single-query scan  ~30 us + the rows at ~7.2 TB/s.
// At 10M x 384 the tensor cores win from 2 queries on (2.36 vs 2.43 ms; 16 queries: 2.4 vs 4.3 ms); at the reference's own
// scale (100k rows, src/constants.rs:93-95) up to 16 query variants stay one multi-query launch (8 variants: 103 us).
static bool tf32_route_is_faster(const csgpu_index *ix, uint32_t b, uint32_t k)
{
    if (b < 2) return false;
    uint64_t rows = 0;
    for (const Shard *sh : ix->shards) rows = std::max<uint64_t>(rows, sh->n_built);
    const double mb = (double)rows * ix->dim_pad * sizeof(float) / 1e6;
    const uint32_t nqb = (b + 127) / 128;
    const double t_tc = (k > 32 ? 260.0 : 205.0) - (rows <= 512 * 1024 ? 20.0 : 0.0) /* two phases instead of three */
                        + mb / 6.4 * (nqb == 1 ? 1.0 : 0.3 + 0.57 * nqb);
    double t_scan = 0.0;
    if (multi_scan_supported(ix->dim4, k)) {
        const uint32_t MQ = multi_scan_max_queries();
        for (uint32_t j = 0; j < b; j += MQ) {
            const uint32_t nq = std::min(MQ, b - j);
            t_scan += nq == 1 ? 30.0 + mb / 7.2 : 45.0 + 8.0 * nq + mb / (nq <= 8 ? 6.3 : 3.7) * (k > 32 ? 1.2 : 1.0);
        }
    } else {
        t_scan = b * (30.0 + mb / 7.2);
    }
    if (i8_eligible(ix, k)) t_scan = std::min(t_scan, b * (60.0 + mb / 4.0 / 6.5));   // byte prefilter on: int8 singles, a quarter of the bytes each
    return t_tc < t_scan;
}

int csgpu_search_batch(const csgpu_index *ix, const float *q, uint32_t q_len, uint32_t b, uint32_t k,
                       uint32_t *out_ids, float *out_dist, uint32_t *out_n)
{
    for (uint32_t j = 0; j < b && out_n; ++j) out_n[j] = 0;
    int rc = check_search_args(ix, q, q_len, k);
    if (rc) return rc;
    if (!all_finite(q, q_len * b)) return fail(CSGPU_ERR_ARG, "query contains NaN/Inf");
    if (k == 0 || b == 0) return CSGPU_OK;
    if (ix->dtype == CSGPU_DTYPE_BF16) return batch_search(ix, q, b, k, out_ids, out_dist, out_n, nullptr);
    // large batches: register-tiled fp32 SIMT GEMM + fused threshold filter (gemm_simt.cuh): 128-query blocks, or one
    // 64-query block for <= 64 queries; up to 32 queries the multi-query scan (8 queries per 3.0 ms pass, 9..16 per 4.9 ms
    // pass at k = 100) is faster.
    bool prefilter = ix->tensor_prefilter;
    for (const Shard *sh : ix->shards) prefilter = prefilter && (sh->shadow_valid || sh->n_built == 0);
    // where the multi-query scan cannot serve (dim % 128 != 0 or k > 256) the alternative is one 2 ms scan per query,
    // which a 128-query SIMT block (20 ms) beats from ~10 queries on
    const char *env_s = getenv("CSGPU_GEMM_MIN_BATCH");   // tuning runs and tests; read per call
    const uint32_t env_min = env_s && *env_s ? (uint32_t)atoi(env_s) : 0u;
    const uint32_t gemm_min = env_min ? env_min : (multi_scan_supported(ix->dim4, k) ? GEMM_MIN_BATCH : GEMM_MIN_BATCH_NO_MULTI);
    // round 2: without a shadow the GEMM-shaped route is the tf32 tensor-core filter straight off the fp32 rows (gemm_tf32.cuh),
    // which takes over wherever the cost model says so — from 2 queries on at 10M rows
    const bool tf32 = !prefilter && !env_min && batch_tf32_route(ix) && tf32_route_is_faster(ix, b, k);
    if ((b >= gemm_min || tf32 || (prefilter && b >= PREFILTER_MIN_BATCH)) && batch_gemm_available(ix)) {
        std::vector<uint32_t> zero_q;
        rc = batch_search(ix, q, b, k, out_ids, out_dist, out_n, &zero_q);
        for (size_t z = 0; z < zero_q.size() && !rc; ++z) {   // zero-norm queries: distance 0.0 everywhere (scan kernel)
            const uint32_t j = zero_q[z];
            rc = search_one(ix, q + (size_t)j * q_len, k, nullptr, 0, out_ids + (size_t)j * k, out_dist + (size_t)j * k,
                            out_n ? out_n + j : nullptr);
        }
        return rc;
    }
    // chunks of up to 16 queries share ONE pass over the corpus (scan_multi.cuh); leftovers of one query,
    // k > 256 or dims that are not a multiple of 128 take the single-query kernel.
    const uint32_t MQ = multi_scan_max_queries();
    const bool multi_ok = multi_scan_supported(ix->dim4, k);
    for (uint32_t j = 0; j < b;) {
        const uint32_t nq = std::min(MQ, b - j);
        if (multi_ok && nq >= 2 && !(nq <= I8_BEATS_MULTI && i8_eligible(ix, k))) {
            rc = search_multi(ix, q + (size_t)j * q_len, nq, k, out_ids + (size_t)j * k, out_dist + (size_t)j * k,
                              out_n ? out_n + j : nullptr);
            if (rc) return rc;
            j += nq;
        } else {
            rc = search_one(ix, q + (size_t)j * q_len, k, nullptr, 0, out_ids + (size_t)j * k, out_dist + (size_t)j * k,
                            out_n ? out_n + j : nullptr);
            if (rc) return rc;
            j += 1;
        }
    }
    return CSGPU_OK;
}

int csgpu_set_tensor_prefilter(csgpu_index *ix, uint32_t enabled)
{
    if (!ix) return fail(CSGPU_ERR_ARG, "null index");
    if (ix->dtype != CSGPU_DTYPE_F32) return fail(CSGPU_ERR_ARG, "the tensor prefilter belongs to an fp32 index (a bf16 index is already on the tensor cores)");
    if (enabled && !bf16_dim_supported(ix->dim)) return fail(CSGPU_ERR_ARG, "tensor prefilter needs dim % 64 == 0 and 64 <= dim <= 512");
    ix->tensor_prefilter = enabled != 0;
    if (ix->built)
        for (Shard *sh : ix->shards) { int rc = shadow_refresh(ix, sh); if (rc) { ix->tensor_prefilter = false; return rc; } }
    return CSGPU_OK;
}

int csgpu_set_byte_prefilter(csgpu_index *ix, uint32_t enabled)
{
    if (!ix) return fail(CSGPU_ERR_ARG, "null index");
    if (ix->dtype != CSGPU_DTYPE_F32) return fail(CSGPU_ERR_ARG, "the byte prefilter belongs to an fp32 index");
    if (enabled && ix->dim_pad > 1024) return fail(CSGPU_ERR_ARG, "byte prefilter needs dim <= 1024");
    ix->byte_prefilter = enabled != 0;
    if (ix->built)
        for (Shard *sh : ix->shards) { int rc = i8_refresh(ix, sh); if (rc) { ix->byte_prefilter = false; return rc; } }
    if (enabled) {
        preload_exchange_kernels(ix);
        int rc = warm_i8_contexts(ix);
        if (rc) return rc;
    }
    return CSGPU_OK;
}

int csgpu_search_variants(const csgpu_index *ix, const float *q, uint32_t q_len, uint32_t b, uint32_t k,
                          uint32_t *out_ids, float *out_dist, uint32_t *out_n)
{
    if (out_n) *out_n = 0;
    int rc = check_search_args(ix, q, q_len, k);
    if (rc) return rc;
    if (b == 0 || b > MAX_BATCH) return fail(CSGPU_ERR_ARG, "b must be in [1, 16] query variants");
    if (ix->dtype != CSGPU_DTYPE_F32) return fail(CSGPU_ERR_ARG, "csgpu_search_variants needs an fp32 index");
    if (!all_finite(q, q_len * b)) return fail(CSGPU_ERR_ARG, "query contains NaN/Inf");
    if (k == 0) return CSGPU_OK;
    bool prefilter = ix->tensor_prefilter && b >= 2;
    for (const Shard *sh : ix->shards) prefilter = prefilter && (sh->shadow_valid || sh->n_built == 0);
    if (prefilter && batch_gemm_available(ix)) return search_variants_prefiltered(ix, q, b, k, out_ids, out_dist, out_n);
    // no shadow: the tf32 tensor-core filter answers the variants as one batch where it beats the multi-query passes
    // (9 variants x top-200 over 10M x 384: 2.5 ms instead of 4.8); same lists bit for bit, deduplicated on the host
    const char *env_s = getenv("CSGPU_GEMM_MIN_BATCH");
    const bool tf32 = b >= 2 && batch_tf32_route(ix) && (env_s && *env_s ? b >= (uint32_t)atoi(env_s) : tf32_route_is_faster(ix, b, k));
    if (tf32) return search_variants_prefiltered(ix, q, b, k, out_ids, out_dist, out_n);
    return search_variants(ix, q, b, k, out_ids, out_dist, out_n);
}

int csgpu_search_variants_tagged(const csgpu_index *ix, const float *q, uint32_t q_len, uint32_t b, uint32_t k,
                                 const csgpu_predicate_t *pred, uint32_t *out_ids, float *out_dist, uint32_t *out_n)
{
    if (out_n) *out_n = 0;
    int rc = check_search_args(ix, q, q_len, k);
    if (rc) return rc;
    if (b == 0 || b > MAX_BATCH) return fail(CSGPU_ERR_ARG, "b must be in [1, 16] query variants");
    if (ix->dtype != CSGPU_DTYPE_F32) return fail(CSGPU_ERR_ARG, "csgpu_search_variants_tagged needs an fp32 index");
    if (!pred) return fail(CSGPU_ERR_ARG, "pred is null");
    if (!pred->file_bitmap && pred->n_file_bits) return fail(CSGPU_ERR_ARG, "file_bitmap is null");
    if (!all_finite(q, q_len * b)) return fail(CSGPU_ERR_ARG, "query contains NaN/Inf");
    if (k == 0) return CSGPU_OK;
    csgpu_predicate_t p = *pred;
    static const uint64_t empty_word = 0;
    if (p.file_bitmap && p.n_file_bits == 0) p.file_bitmap = &empty_word;   // an empty bitmap allows nothing
    return search_variants(ix, q, b, k, out_ids, out_dist, out_n, &p);
}

int csgpu_search_keys_device(const csgpu_index *ix, const float *q_dev, uint32_t k, uint64_t *out_keys_dev, void *stream)
{
    if (!ix) return fail(CSGPU_ERR_ARG, "null index");
    if (!ix->built) return fail(CSGPU_ERR_NOT_BUILT, "Index not built. Call build_index() after inserting chunks.");
    if (ix->shards.size() != 1) return fail(CSGPU_ERR_ARG, "device entry points need a single-device index");
    if (ix->dtype != CSGPU_DTYPE_F32) return fail(CSGPU_ERR_ARG, "device entry points need an fp32 index");
    if (!q_dev || !out_keys_dev) return fail(CSGPU_ERR_ARG, "null device pointer");
    if (k == 0 || k > CSGPU_MAX_K) return fail(CSGPU_ERR_ARG, "k must be in [1, 1024]");
    Shard *sh = ix->shards[0];
    SearchCtx *c = nullptr;
    int rc = dev_ctx_begin(ix, sh, (cudaStream_t)stream, &c);
    if (rc) return rc;
    {
        DeviceGuard dg(sh->device);
        rc = enqueue_keys_device(ix, sh, c, q_dev, k, out_keys_dev, (cudaStream_t)stream, nullptr, 0);
    }
    dev_ctx_end(sh, c, (cudaStream_t)stream);
    return rc;
}

int csgpu_merge_keys_device(const csgpu_index *ix, const uint64_t *keys_dev, uint32_t n_lists, uint32_t k,
                            uint64_t *out_keys_dev, void *stream)
{
    if (!ix) return fail(CSGPU_ERR_ARG, "null index");
    if (!keys_dev || !out_keys_dev) return fail(CSGPU_ERR_ARG, "null device pointer");
    if (k == 0 || k > CSGPU_MAX_K || n_lists == 0) return fail(CSGPU_ERR_ARG, "bad k / n_lists");
    DeviceGuard dg(ix->shards[0]->device);
    return enqueue_merge(keys_dev, n_lists, 1, k, out_keys_dev, (cudaStream_t)stream);
}

int csgpu_merge_keys_batch_device(const csgpu_index *ix, const uint64_t *keys_dev, uint32_t n_lists, uint32_t nq, uint32_t k,
                                  uint64_t *out_keys_dev, void *stream)
{
    if (!ix) return fail(CSGPU_ERR_ARG, "null index");
    if (!keys_dev || !out_keys_dev) return fail(CSGPU_ERR_ARG, "null device pointer");
    if (k == 0 || k > CSGPU_MAX_K || n_lists == 0 || nq == 0) return fail(CSGPU_ERR_ARG, "bad k / n_lists / nq");
    DeviceGuard dg(ix->shards[0]->device);
    return enqueue_merge(keys_dev, n_lists, nq, k, out_keys_dev, (cudaStream_t)stream);
}

void csgpu_encode_keys(const uint32_t *ids, const float *dist, uint32_t n, uint32_t k, uint64_t *out_keys)
{
    for (uint32_t i = 0; i < k; ++i) {
        if (i >= n) { out_keys[i] = KEY_EMPTY; continue; }
        uint32_t bits;
        memcpy(&bits, &dist[i], sizeof bits);
        out_keys[i] = ((uint64_t)okey_from_bits(bits) << 32) | ids[i];
    }
}

void csgpu_decode_keys(const uint64_t *keys, uint32_t k, uint32_t *out_ids, float *out_dist, uint32_t *out_n)
{
    decode_keys(keys, k, out_ids, out_dist, out_n);
}

// ---- snapshot / hydrate (snapshot.cu) -----------------------------------------------------------------
static int load_finish(csgpu_index *ix)
{
    int rc = upload_zero_ids(ix);
    if (rc) return rc;
    if ((rc = tag_stats_refresh(ix))) return rc;
    for (Shard *sh : ix->shards) if ((rc = batch_after_build(ix, sh))) return rc;
    for (Shard *sh : ix->shards) if ((rc = i8_refresh(ix, sh))) return rc;
    if (ix->byte_prefilter && (rc = warm_i8_contexts(ix))) return rc;
    ix->tombstones = 0;
    ix->built = true;
    return CSGPU_OK;
}

int csgpu_save(const csgpu_index *ix, const char *dir) { return snapshot_save(ix, dir); }

int csgpu_load(csgpu_index *ix, const char *dir) { return snapshot_load(ix, dir, csgpu_reserve, load_finish); }

// ---- fused cross-GPU exchange (rank-per-GPU) ---------------------------------------------------------
static void exchange_free(csgpu_index *ix)
{
    Exchange *x = ix->xchg;
    if (!x) return;
    DeviceGuard dg(ix->shards[0]->device);
    for (uint32_t p = 0; p < x->world; ++p)
        if (x->peer_ipc[p] && x->peer_base[p]) cudaIpcCloseMemHandle(x->peer_base[p]);
    cudaFree(x->dev);
    cudaFree(x->base);
    if (x->status_pin) cudaFreeHost(x->status_pin);
    delete x;
    ix->xchg = nullptr;
}

// A timed-out wait leaves the ranks' sequence numbers out of step and the results undefined: every later exchange
// search fails with CSGPU_ERR_NCCL until the exchange is set up again (the FFI returns errors, never aborts).
static int exchange_healthy(const csgpu_index *ix)
{
    const Exchange *x = ix->xchg;
    if (x && x->status_pin && *reinterpret_cast<volatile unsigned *>(x->status_pin) != 0)
        return fail(CSGPU_ERR_NCCL, "cross-GPU exchange timed out earlier (a peer rank never delivered its keys); "
                                    "results since then are undefined — csgpu_exchange_create/connect again on every rank");
    return CSGPU_OK;
}

int csgpu_exchange_create(csgpu_index *ix, uint32_t world, uint32_t rank, void *out_handle)
{
    if (!ix || !out_handle) return fail(CSGPU_ERR_ARG, "null argument");
    if (ix->shards.size() != 1) return fail(CSGPU_ERR_ARG, "the fused exchange needs a single-device index (one rank per GPU)");
    if (world == 0 || world > (uint32_t)XCHG_MAX_WORLD || rank >= world) return fail(CSGPU_ERR_ARG, "world must be in [1, 8] and rank < world");
    exchange_free(ix);
    DeviceGuard dg(ix->shards[0]->device);
    Exchange *x = new Exchange();
    x->world = world; x->rank = rank;
    ix->xchg = x;
    CS_CUDA(cudaMalloc(&x->base, x->block_bytes()));
    CS_CUDA(cudaMemset(x->base, 0, x->block_bytes()));
    CS_CUDA(cudaMalloc(&x->dev, sizeof(ExchangeDev)));
    CS_CUDA(cudaHostAlloc(&x->status_pin, 64, cudaHostAllocMapped | cudaHostAllocPortable));
    memset(x->status_pin, 0, 64);
    x->timeout_ns = ix->exchange_timeout_ns ? ix->exchange_timeout_ns : default_exchange_timeout_ns();
    static_assert(sizeof(cudaIpcMemHandle_t) == CSGPU_EXCHANGE_HANDLE_BYTES, "handle size");
    cudaIpcMemHandle_t h;
    CS_CUDA(cudaIpcGetMemHandle(&h, x->base));
    memcpy(out_handle, &h, sizeof h);
    return CSGPU_OK;
}

static int exchange_finish_connect(csgpu_index *ix)
{
    Exchange *x = ix->xchg;
    ExchangeDev d;
    memset(&d, 0, sizeof d);
    for (uint32_t p = 0; p < x->world; ++p) {
        d.slots[p] = reinterpret_cast<uint64_t *>(x->peer_base[p]);
        d.flags[p] = reinterpret_cast<unsigned *>(reinterpret_cast<char *>(x->peer_base[p]) + x->slots_bytes());
    }
    DeviceGuard dg(ix->shards[0]->device);
    void *st_dev = nullptr;
    CS_CUDA(cudaHostGetDevicePointer(&st_dev, x->status_pin, 0));
    d.status = reinterpret_cast<unsigned *>(st_dev);
    d.wait_ring = reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(x->base) + x->stats_off());
    d.timeout_ns = x->timeout_ns;
    d.world = x->world; d.rank = x->rank; d.kmax = x->kmax;
    d.root = -1;   // rank-per-GPU processes: every rank ends with the global top-k
    CS_CUDA(cudaMemcpy(x->dev, &d, sizeof d, cudaMemcpyHostToDevice));
    preload_exchange_kernels(ix);
    x->connected = true;
    return CSGPU_OK;
}

int csgpu_exchange_connect(csgpu_index *ix, const void *handles)
{
    if (!ix || !ix->xchg || !handles) return fail(CSGPU_ERR_ARG, "csgpu_exchange_create first");
    Exchange *x = ix->xchg;
    DeviceGuard dg(ix->shards[0]->device);
    for (uint32_t p = 0; p < x->world; ++p) {
        if (p == x->rank) { x->peer_base[p] = x->base; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, reinterpret_cast<const char *>(handles) + (size_t)p * CSGPU_EXCHANGE_HANDLE_BYTES, sizeof h);
        CS_CUDA(cudaIpcOpenMemHandle(&x->peer_base[p], h, cudaIpcMemLazyEnablePeerAccess));
        x->peer_ipc[p] = true;
    }
    return exchange_finish_connect(ix);
}

int csgpu_exchange_connect_local(csgpu_index *ix, csgpu_index *const *peers)
{
    if (!ix || !ix->xchg || !peers) return fail(CSGPU_ERR_ARG, "csgpu_exchange_create first");
    Exchange *x = ix->xchg;
    for (uint32_t p = 0; p < x->world; ++p) {
        if (p == x->rank) { x->peer_base[p] = x->base; continue; }
        const csgpu_index *o = peers[p];
        if (!o || !o->xchg || o->xchg->world != x->world || o->xchg->rank != p)
            return fail(CSGPU_ERR_ARG, "peer index has no matching exchange (world/rank)");
        const int da = ix->shards[0]->device, db = o->shards[0]->device;
        if (da != db) {   // same process, different GPUs: plain peer access
            int can = 0;
            cudaDeviceCanAccessPeer(&can, da, db);
            if (!can) return fail(CSGPU_ERR_CUDA, "no peer access between the devices of the exchange");
            DeviceGuard dg(da);
            cudaError_t pe = cudaDeviceEnablePeerAccess(db, 0);
            if (pe != cudaSuccess) cudaGetLastError();   // already enabled
        }
        x->peer_base[p] = o->xchg->base;
    }
    return exchange_finish_connect(ix);
}

int csgpu_search_keys_exchange_device(const csgpu_index *ix, const float *q_dev, uint32_t k, uint64_t *out_keys_dev, void *stream)
{
    if (!ix) return fail(CSGPU_ERR_ARG, "null index");
    if (!ix->built) return fail(CSGPU_ERR_NOT_BUILT, "Index not built. Call build_index() after inserting chunks.");
    if (!ix->xchg || !ix->xchg->connected) return fail(CSGPU_ERR_ARG, "exchange is not connected (csgpu_exchange_create/connect)");
    if (ix->dtype != CSGPU_DTYPE_F32) return fail(CSGPU_ERR_ARG, "device entry points need an fp32 index");
    if (!q_dev || !out_keys_dev) return fail(CSGPU_ERR_ARG, "null device pointer");
    if (k == 0 || k > CSGPU_MAX_K) return fail(CSGPU_ERR_ARG, "k must be in [1, 1024]");
    if (int rc = exchange_healthy(ix)) return rc;
    Shard *sh = ix->shards[0];
    SearchCtx *c = nullptr;
    int rc = dev_ctx_begin(ix, sh, (cudaStream_t)stream, &c);
    if (rc) return rc;
    {
        DeviceGuard dg(sh->device);
        const uint32_t seq = ix->xchg->seq.fetch_add(1) + 1;   // every rank issues the same sequence of searches
        rc = enqueue_keys_device(ix, sh, c, q_dev, k, out_keys_dev, (cudaStream_t)stream, ix->xchg->dev, seq);
    }
    dev_ctx_end(sh, c, (cudaStream_t)stream);
    return rc;
}

// Host-pointer form of the exchange search for rank-per-GPU processes: pinned staging of the query, ONE fused launch
// (scan + peer stores + flag wait + global merge) whose last CTA writes the global top-k into mapped host memory, one
// stream synchronised. A collective: every rank calls it for every query, in the same order, one call at a time per rank.
int csgpu_search_exchange(const csgpu_index *ix, const float *q, uint32_t q_len, uint32_t k,
                          uint32_t *out_ids, float *out_dist, uint32_t *out_n)
{
    if (out_n) *out_n = 0;
    int rc = check_search_args(ix, q, q_len, k);
    if (rc) return rc;
    if (ix->shards.size() != 1) return fail(CSGPU_ERR_ARG, "csgpu_search_exchange needs a single-device index (one rank per GPU)");
    if (ix->dtype != CSGPU_DTYPE_F32) return fail(CSGPU_ERR_ARG, "the exchange needs an fp32 index");
    if (!ix->xchg || !ix->xchg->connected) return fail(CSGPU_ERR_ARG, "exchange is not connected (csgpu_exchange_create/connect)");
    if (!all_finite(q, q_len)) return fail(CSGPU_ERR_ARG, "query contains NaN/Inf");
    if (k == 0) return CSGPU_OK;
    if ((rc = exchange_healthy(ix))) return rc;
    Shard *sh = ix->shards[0];
    SearchCtx *c = nullptr;
    if ((rc = ctx_acquire(ix, sh, &c))) return rc;
    auto body = [&]() -> int {
        DeviceGuard dg(sh->device);
        const size_t qbytes = (size_t)ix->dim_pad * sizeof(float);
        memset(c->q_pin, 0, qbytes);
        memcpy(c->q_pin, q, (size_t)ix->dim * sizeof(float));
        CS_CUDA(cudaMemcpyAsync(c->q_dev, c->q_pin, qbytes, cudaMemcpyHostToDevice, c->stream));
        CS_CUDA(cudaEventRecord(c->ev0, c->stream));
        const uint32_t seq = ix->xchg->seq.fetch_add(1) + 1;
        int r = enqueue_keys_device(ix, sh, c, c->q_dev, k, c->out_pin, c->stream, ix->xchg->dev, seq);
        if (r) return r;
        CS_CUDA(cudaEventRecord(c->ev1, c->stream));
        CS_CUDA(cudaStreamSynchronize(c->stream));
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) == cudaSuccess) ix->last_search_us.store(ms * 1000.f);
        if ((r = exchange_healthy(ix))) return r;
        decode_keys(c->out_pin, k, out_ids, out_dist, out_n);
        return CSGPU_OK;
    };
    rc = body();
    if (rc) { DeviceGuard dg(sh->device); cudaStreamSynchronize(c->stream); }
    ctx_release(sh, c);
    return rc;
}

int csgpu_search_tagged_keys_device(const csgpu_index *ix, const float *q_dev, uint32_t k, const csgpu_predicate_t *pred,
                                    uint32_t exchange, uint64_t *out_keys_dev, void *stream)
{
    if (!ix) return fail(CSGPU_ERR_ARG, "null index");
    if (!ix->built) return fail(CSGPU_ERR_NOT_BUILT, "Index not built. Call build_index() after inserting chunks.");
    if (ix->shards.size() != 1) return fail(CSGPU_ERR_ARG, "device entry points need a single-device index");
    if (ix->dtype != CSGPU_DTYPE_F32) return fail(CSGPU_ERR_ARG, "device entry points need an fp32 index");
    if (exchange && (!ix->xchg || !ix->xchg->connected)) return fail(CSGPU_ERR_ARG, "exchange is not connected (csgpu_exchange_create/connect)");
    if (!q_dev || !out_keys_dev || !pred) return fail(CSGPU_ERR_ARG, "null pointer");
    if (k == 0 || k > CSGPU_MAX_K) return fail(CSGPU_ERR_ARG, "k must be in [1, 1024]");
    if (pred->file_bitmap && pred->n_file_bits == 0) return fail(CSGPU_ERR_ARG, "file_bitmap with n_file_bits == 0");
    if (exchange) if (int rc = exchange_healthy(ix)) return rc;
    Shard *sh = ix->shards[0];
    SearchCtx *c = nullptr;
    int rc = dev_ctx_begin(ix, sh, (cudaStream_t)stream, &c);
    if (rc) return rc;
    {
        DeviceGuard dg(sh->device);
        const uint32_t seq = exchange ? ix->xchg->seq.fetch_add(1) + 1 : 0;
        rc = enqueue_keys_device(ix, sh, c, q_dev, k, out_keys_dev, (cudaStream_t)stream, exchange ? ix->xchg->dev : nullptr, seq,
                                 pred->file_bitmap, pred->file_bitmap ? pred->n_file_bits : 0, pred);
    }
    dev_ctx_end(sh, c, (cudaStream_t)stream);
    return rc;
}

int csgpu_exchange_status(const csgpu_index *ix, uint32_t *timed_out)
{
    if (!ix || !ix->xchg || !timed_out) return fail(CSGPU_ERR_ARG, "no exchange");
    *timed_out = *reinterpret_cast<volatile unsigned *>(ix->xchg->status_pin);   // pinned host word the kernel writes: no device round trip
    return CSGPU_OK;
}

int csgpu_exchange_set_timeout_ms(csgpu_index *ix, uint32_t ms)
{
    if (!ix) return fail(CSGPU_ERR_ARG, "null index");
    ix->exchange_timeout_ns = (uint64_t)std::max(ms, 1u) * 1000000ull;
    Exchange *x = ix->xchg;
    if (x) {
        x->timeout_ns = ix->exchange_timeout_ns;
        if (x->connected) {   // live exchange: patch the device copy of the table (callers quiesce searches first, as for connect)
            DeviceGuard dg(ix->shards[0]->device);
            CS_CUDA(cudaMemcpy(reinterpret_cast<char *>(x->dev) + offsetof(ExchangeDev, timeout_ns), &x->timeout_ns,
                               sizeof x->timeout_ns, cudaMemcpyHostToDevice));
        }
    }
    // in-process groups are re-created with the new bound
    std::lock_guard<std::mutex> lk(ix->group_mu);
    for (GroupCtx *g : ix->free_groups) {
        ix->all_groups.erase(std::find(ix->all_groups.begin(), ix->all_groups.end(), g));
        group_destroy(g);
    }
    ix->free_groups.clear();
    return CSGPU_OK;
}

int csgpu_exchange_wait_stats(const csgpu_index *ix, uint64_t *out_ns, uint32_t max_queries, uint32_t *n_queries)
{
    if (!ix || !ix->xchg || !out_ns || !n_queries) return fail(CSGPU_ERR_ARG, "no exchange / null argument");
    const Exchange *x = ix->xchg;
    const uint32_t seq = x->seq.load();
    const uint32_t n = std::min(std::min(seq, (uint32_t)XCHG_STATS_RING), max_queries);
    *n_queries = n;
    if (n == 0) return CSGPU_OK;
    std::vector<unsigned long long> ring((size_t)XCHG_STATS_RING * XCHG_MAX_WORLD);
    DeviceGuard dg(ix->shards[0]->device);
    CS_CUDA(cudaMemcpy(ring.data(), reinterpret_cast<const char *>(x->base) + x->stats_off(), ring.size() * sizeof(unsigned long long),
                       cudaMemcpyDeviceToHost));
    for (uint32_t i = 0; i < n; ++i) {   // oldest first: queries seq-n+1 .. seq (1-based), ring slot = (query - 1) % RING
        const uint32_t query = seq - n + 1 + i;
        for (uint32_t p = 0; p < x->world; ++p)
            out_ns[(size_t)i * x->world + p] = ring[(size_t)((query - 1) % XCHG_STATS_RING) * XCHG_MAX_WORLD + p];
    }
    return CSGPU_OK;
}

void csgpu_exchange_destroy(csgpu_index *ix)
{
    if (ix) exchange_free(ix);
}

int csgpu_stats(const csgpu_index *ix, csgpu_stats_t *out)
{
    if (!ix || !out) return fail(CSGPU_ERR_ARG, "null argument");
    memset(out, 0, sizeof *out);
    out->dim = ix->dim;
    out->dtype = ix->dtype;
    out->n_devices = (uint32_t)ix->shards.size();
    out->built = ix->built ? 1 : 0;
    out->abi_version = CSGPU_ABI_VERSION;
    out->zero_norm_rows = ix->zero_ids.size();
    out->nonfinite_rows = ix->nonfinite_rows;
    out->tombstones = ix->tombstones;
    out->last_search_us = ix->last_search_us.load();
    for (size_t g = 0; g < ix->shards.size(); ++g) {
        const Shard *sh = ix->shards[g];
        out->live_rows += sh->n_built;
        out->pending_rows += sh->n_total - sh->n_built;
        out->rows_per_device[g] = sh->n_built;
        const size_t row_bytes = ix->dtype == CSGPU_DTYPE_BF16 ? (size_t)ix->dim * 2 : (size_t)ix->dim_pad * sizeof(float);
        out->bytes_on_device += sh->cap * (row_bytes + 2 * sizeof(uint32_t) + 1) + sh->stage_cap * (size_t)ix->dim_pad * sizeof(float);
    }
    out->live_rows += ix->zero_ids.size();
    out->coalesced_passes = ix->coalescer.passes.load();
    out->coalesced_queries = ix->coalescer.queries.load();
    out->prefilter_rescored = ix->prefilter_rescored.load();
    for (const Shard *sh : ix->shards) out->shadow_bytes += sh->shadow_rows * (uint64_t)ix->dim * 2;
    out->bytes_on_device += out->shadow_bytes;
    for (const Shard *sh : ix->shards) out->byte_shadow_bytes += sh->i8_rows * ((uint64_t)((ix->dim4 + 31) / 32) * 128 + sizeof(uint32_t));
    out->bytes_on_device += out->byte_shadow_bytes;
    out->byte_searches = ix->byte_searches.load();
    out->byte_fallbacks = ix->byte_fallbacks.load();
    out->byte_candidates = ix->byte_candidates.load();
    out->byte_rescored = ix->byte_rescored.load();
    out->batch_route = ix->batch_route.load();
    out->filter_max_err = ix->filter_max_err.load();
    return CSGPU_OK;
}

}  // extern "C"
