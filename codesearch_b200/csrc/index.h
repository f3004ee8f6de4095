// index.h — host-side state of a csgpu_index (library-private; the public surface is include/csgpu.h).
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <mutex>
#include <string>
#include <vector>

#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/csgpu.h"

struct csgpu_index;

namespace csgpu {

// thread-local error text (csgpu_last_error)
void set_error(const std::string &msg);
int fail(int code, const std::string &msg);
int fail_cuda(cudaError_t e, const char *what, const char *file, int line);
extern std::atomic<uint64_t> g_kernel_launches;

void count_launch(uint64_t n = 1);
void decode_keys(const uint64_t *keys, uint32_t k, uint32_t *out_ids, float *out_dist, uint32_t *out_n);

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int d) { cudaGetDevice(&prev); if (prev != d) cudaSetDevice(d); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

#define CS_CUDA(call)                                                                  \
    do {                                                                               \
        cudaError_t e__ = (call);                                                      \
        if (e__ != cudaSuccess) return ::csgpu::fail_cuda(e__, #call, __FILE__, __LINE__); \
    } while (0)

// Per-search scratch: one stream + buffers, so searches from different host threads overlap
// (VectorStore::search is &self and is called from rayon/tokio threads concurrently,
// /root/reference/src/search/mod.rs:508-511, src/server/mod.rs:545-548).
struct SearchCtx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float *q_dev = nullptr;         // [max_b * dim_pad]
    float *q_pin = nullptr;         // pinned staging for the query
    uint64_t *cand = nullptr;       // [grid_max * CSGPU_MAX_K]
    uint64_t *gather = nullptr;     // [8 * 4096] per-shard results gathered for the cross-GPU merge (peer-copy route)
    unsigned *ticket = nullptr;
    uint64_t *out_dev = nullptr;    // [max_b * CSGPU_MAX_K]
    uint64_t *out_pin = nullptr;    // pinned, device-mapped; kernels write results straight here
    uint64_t *bitmap_dev = nullptr; // filter bitmap staging
    size_t bitmap_cap = 0;          // in u64 words
    size_t cand_cap = 0;            // in keys
    void *i8_scratch = nullptr;     // byte prefilter (scan_i8.cu): per-warp minima, counters, candidate regions; allocated with the
                                    // context when the prefilter is on (csgpu_set_byte_prefilter warms one per shard)
    uint64_t *i8_status = nullptr;  // pinned, device-mapped: [0] fallback flag, [1] statistics of the last launch
    unsigned long long *timing = nullptr;   // CSGPU_SCAN_TIMING=1: per-CTA globaltimer stamps of the last fp32 scan (diagnostic)
    cudaEvent_t busy = nullptr;     // device entry points: recorded on the caller's stream after the enqueue, waited on by the next
    bool busy_recorded = false;
};

// In-process multi-device search (csgpu_create with n_devices > 1, the one-process host of the reference:
// /root/reference/src/search/mod.rs:508-511, src/server/mod.rs:27): one SearchCtx per shard wired together by exchange
// blocks in every shard's HBM, so a query is N concurrent scan launches whose tails push their k keys into the ROOT
// shard's block over NVLink; the root's last CTA merges and writes the global top-k straight into mapped host memory.
// One group per concurrent host search (pooled like SearchCtx).
struct ExchangeDev;
struct GroupCtx {
    std::vector<SearchCtx *> ctx;         // [n_shards]
    std::vector<void *> xbase;            // [n_shards] exchange block on shard g's device
    std::vector<ExchangeDev *> xdev;      // [n_shards] device copy of the pointer table as seen from shard g
    unsigned *status_pin = nullptr;       // pinned, device-mapped: != 0 after a timed-out wait (sticky)
    uint32_t seq = 0;
};

// Scratch of the batched GEMM-shaped path (gemm_topk.cu): one per shard, serialised by Shard::batch_mu.
struct BatchCtx {
    cudaStream_t stream = nullptr;
    float *q_f32 = nullptr;        // [1024][dim] raw queries
    void *q_prep = nullptr;        // [1024][dim_pad] unit-normalised queries: fp32, or bf16 on a bf16 index
    uint8_t *flags = nullptr;      // [1024] zero-norm query flags
    float *thr = nullptr;          // [1024] threshold distance per query
    unsigned *count = nullptr, *count_saved = nullptr;
    uint64_t *cand = nullptr;      // [1024][BF_CAP]
    uint64_t *out = nullptr;       // [1024][CSGPU_MAX_K]
    unsigned *scalar = nullptr;     // [0] sticky overflow flag, [2..3] rows rescored (u64)
    unsigned *seg_count = nullptr;  // [1024][2 * 148] per-(query, segment) candidate counts of the tensor-core kernel
    float *q_pin = nullptr;
    uint64_t *out_pin = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;   // device time of a batch (created once: create + destroy per call cost ~10 us)
};

// Fused cross-GPU exchange state of a single-device index (scan.cuh: ExchangeDev). One cudaMalloc block:
// [2][world][kmax] keys, then [2][world] flags, then one status word — shared with peers through cudaIpc.
struct Exchange {
    uint32_t world = 0, rank = 0, kmax = CSGPU_MAX_K;
    unsigned *status_pin = nullptr;       // pinned, device-mapped status word: the host reads it without a device round trip
    uint64_t timeout_ns = 4000000000ull;  // bound of the in-kernel flag wait (csgpu_exchange_set_timeout_ms / CSGPU_EXCHANGE_TIMEOUT_MS)
    void *base = nullptr;                 // local block
    void *peer_base[8] = {};              // peers' blocks as mapped here (own entry = base)
    bool peer_ipc[8] = {};                // opened with cudaIpcOpenMemHandle (needs cudaIpcCloseMemHandle)
    ExchangeDev *dev = nullptr;           // device copy of the pointer table
    bool connected = false;
    std::atomic<uint32_t> seq{0};
    size_t slots_bytes() const { return (size_t)2 * world * kmax * sizeof(uint64_t); }
    size_t flags_bytes() const { return (size_t)2 * world * sizeof(unsigned) + 64; }
    size_t stats_off() const { return slots_bytes() + flags_bytes(); }   // [XCHG_STATS_RING] u64 wait times (ns) of the last queries
    size_t block_bytes() const { return stats_off() + (size_t)1024 * sizeof(uint64_t); }
};

// Host micro-batcher in front of csgpu_search. Concurrent single-query searches with the same k
// (the <= 9 query variants of /root/reference/src/search/mod.rs:508-511 arrive from rayon threads; MCP/HTTP readers
// from tokio tasks, src/mcp/mod.rs:251-252, src/server/mod.rs:545-548) are coalesced into ONE multi-query pass over
// the corpus (scan_multi.cuh, up to 8 queries per pass). Group-commit style: the first caller to find no pass in
// flight becomes the leader, takes whatever is queued and launches; callers arriving during a pass queue up and
// ride the next one. An idle index therefore adds no latency to a lone query.
struct PendingSearch {
    const float *q; uint32_t k;
    uint32_t *out_ids; float *out_dist; uint32_t *out_n;
    int rc = 0; bool done = false; std::string err;
};
struct Coalescer {
    std::mutex mu;
    std::condition_variable cv;
    std::vector<PendingSearch *> queue;
    bool leader_active = false;
    std::atomic<uint32_t> enabled{0};
    uint32_t window_us = 0;               // optional: the leader lingers this long for company before launching
    std::atomic<uint64_t> passes{0}, queries{0};
};

struct Shard {
    int device = 0;
    // bf16 index (dtype == CSGPU_DTYPE_BF16): built rows live here, pending rows in `stage` as fp32
    void *rows_bf16 = nullptr;     // [cap, dim] bf16 unit vectors
    float *stage = nullptr;        // [stage_cap, dim] fp32, rows appended since the last build
    uint64_t stage_cap = 0;
    CUtensorMap map_c;             // TMA map over the built rows (bf16 or fp32), re-encoded at every build
    CUtensorMap map_c2;            // bf16 only: same matrix, box = 128 rows (one CTA's half tile in the CTA-pair kernel)
    bool map_valid = false;
    void *shadow_bf16 = nullptr;   // fp32 index + tensor prefilter: bf16 image of the built rows (rescore.cuh)
    CUtensorMap map_shadow;        // TMA map over the shadow (box = GT_BLOCK_N rows)
    bool shadow_valid = false;
    uint64_t shadow_rows = 0;
    uint8_t *shadow_i8 = nullptr;  // fp32 index + byte prefilter: int8 image of the built rows, [n_built][128 * ceil(dim/128)] (scan_i8.cuh)
    uint32_t *meta_i8 = nullptr;   // [n_built] half2(scale, error bound) per row
    bool i8_valid = false;
    uint64_t i8_rows = 0;
    BatchCtx *batch = nullptr;
    std::mutex batch_mu;
    float *rows = nullptr;     // [cap, dim_pad] fp32; rows [0, n_built) are unit vectors
    uint32_t *ids = nullptr;   // [cap]
    uint32_t *tags = nullptr;  // [cap] packed (lang_id << 27 | file_id) row tags, CSGPU_TAG_NONE when untagged (§8f N4)
    uint8_t *status = nullptr; // [cap] ROW_* (all ROW_OK in [0, n_built) after build)
    uint64_t n_built = 0;      // searchable rows
    uint64_t n_total = 0;      // built + pending
    uint64_t cap = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;  // build/append stream
    std::mutex ctx_mu;
    std::vector<SearchCtx *> free_ctx;
    std::vector<SearchCtx *> all_ctx;
    SearchCtx *dev_ctx = nullptr;   // device entry points only (never in free_ctx): its scratch is in flight on the CALLER's stream
    std::mutex dev_mu;
};

// gemm_topk.cu (batched GEMM-shaped path for both index dtypes + bf16 storage hooks)
bool bf16_dim_supported(uint32_t dim);
int bf16_reserve_rows(const csgpu_index *ix, Shard *sh, uint64_t rows);
int bf16_reserve_stage(const csgpu_index *ix, Shard *sh, uint64_t pending);
int bf16_convert_pending(const csgpu_index *ix, Shard *sh);
bool batch_f32_dim_supported(uint32_t dim_pad);
int batch_after_build(const csgpu_index *ix, Shard *sh);
int tag_stats_refresh(csgpu_index *ix);   // csgpu.cu
int shadow_refresh(const csgpu_index *ix, Shard *sh);
void batch_free_ctx(Shard *sh);
bool batch_gemm_available(const csgpu_index *ix);
bool batch_tf32_route(const csgpu_index *ix);   // fp32 index whose batches run the tf32 tensor-core filter (gemm_tf32.cuh)
int batch_search(const csgpu_index *ix, const float *q, uint32_t b, uint32_t k,
                 uint32_t *out_ids, float *out_dist, uint32_t *out_n, std::vector<uint32_t> *zero_queries);

// scan_i8.cu (byte prefilter of the single-query path)
int i8_refresh(const csgpu_index *ix, Shard *sh);
void i8_free_shard(Shard *sh);
void i8_free_ctx(SearchCtx *c);
bool i8_eligible(const csgpu_index *ix, uint32_t k);
int enqueue_scan_i8(const csgpu_index *ix, const Shard *sh, SearchCtx *c, const float *q_dev, uint32_t k,
                    bool with_zero_ids, uint64_t *out_keys, cudaStream_t st, bool host_status = true,
                    const uint64_t *bitmap_dev = nullptr, uint64_t n_bits = 0,
                    const csgpu_predicate_t *pred = nullptr /* filters as in enqueue_scan: id bitmap, or tag predicate (+ file bitmap) */);
void i8_preload(const csgpu_index *ix);
int i8_prepare_ctx(SearchCtx *c);   // allocates the context's scratch (device of the current context = c->device)
const unsigned *i8_status_dev(const SearchCtx *c);   // device word: != 0 after a launch that needs the fp32 scan instead

// snapshot.cu
int snapshot_save(const csgpu_index *ix, const char *dir);
int snapshot_load(csgpu_index *ix, const char *dir, int (*reserve)(csgpu_index *, uint64_t), int (*finish)(csgpu_index *));

// scan_filtered.cu
struct ScanArgs;
cudaError_t launch_scan_filtered(const ScanArgs &a, uint32_t grid, size_t smem, cudaStream_t st);

// scan_multi.cu
struct MultiArgs;
bool multi_scan_supported(uint32_t dim4, uint32_t k);
uint32_t multi_scan_max_queries();
uint32_t multi_scan_ctas_per_sm(uint32_t dim4, uint32_t k, uint32_t nq, uint32_t kpad);
uint32_t multi_scan_rows_per_iter(uint32_t dim4, uint32_t nq);
uint32_t multi_scan_cap(uint32_t k, uint32_t nq, uint32_t grid);
cudaError_t launch_scan_multi(const MultiArgs &a, uint32_t grid, cudaStream_t st);

}  // namespace csgpu

struct csgpu_index {
    uint32_t dim = 0, dim_pad = 0, dim4 = 0;
    uint32_t dtype = 0;
    bool built = false;
    bool tensor_prefilter = false;        // csgpu_set_tensor_prefilter: batches run tcgen05 on a bf16 shadow + exact fp32 rescoring
    std::vector<csgpu::Shard *> shards;
    std::vector<uint32_t> zero_ids;       // ascending; rows with |v| = 0 (host truth)
    std::vector<uint32_t> zero_tags;      // parallel to zero_ids
    uint32_t *zero_ids_dev = nullptr;     // on shards[0]->device: [n_zero] ids then [n_zero] tags
    uint64_t nonfinite_rows = 0;
    uint64_t tombstones = 0;
    mutable std::atomic<float> last_search_us{0.f};
    // Multi-query scans in flight on this index (scan_multi.cuh). Their last nq <= 16 CTAs wait, resident, for the rest of
    // their grid; with at most MULTI_SLOTS such kernels running at once the waiting CTAs can never fill a device's CTA slots
    // (6 x 16 = 96 < 148 even at one CTA per SM), whatever else runs — so every grid always gets the slots it needs to finish.
    static constexpr int MULTI_SLOTS = 6;
    mutable std::mutex multi_mu;
    mutable std::condition_variable multi_cv;
    mutable int multi_running = 0;
    // Tag statistics of the built rows, refreshed at every build / load (tag_stats_refresh, csgpu.cu): rows per language id and
    // the largest file id. Only used to ESTIMATE a predicate's density when a search has to pick a route.
    uint64_t lang_rows[32] = {};
    uint64_t tagged_rows = 0;
    uint32_t max_file_id = 0;
    mutable std::atomic<uint32_t> batch_route{0};          // CSGPU_ROUTE_* of the last GEMM-shaped batch
    mutable std::atomic<float> filter_max_err{0.f};        // largest |d_filter - d_f32| its rescoring saw
    mutable std::atomic<uint64_t> prefilter_rescored{0};   // fp32 rows read by the last tensor-prefilter batch chunk
    bool byte_prefilter = false;          // csgpu_set_byte_prefilter: csgpu_search streams an int8 shadow + exact fp32 rescoring
    mutable std::atomic<uint64_t> byte_searches{0}, byte_fallbacks{0}, byte_candidates{0}, byte_rescored{0};
    csgpu::Exchange *xchg = nullptr;      // rank-per-GPU fused exchange (csgpu_exchange_*)
    mutable std::mutex group_mu;          // in-process multi-device search: pool of wired context groups
    mutable std::vector<csgpu::GroupCtx *> free_groups, all_groups;
    uint64_t exchange_timeout_ns = 0;     // csgpu_exchange_set_timeout_ms; 0 = default (CSGPU_EXCHANGE_TIMEOUT_MS or 4 s)
    mutable std::atomic<int> fused_local{-1};   // -1 unknown, 0 = no peer access between some pair (peer copies + merge launch), 1 = fused
    mutable csgpu::Coalescer coalescer;   // csgpu_set_coalescing
};
