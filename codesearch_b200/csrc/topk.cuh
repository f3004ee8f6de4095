// topk.cuh — sortable result keys and the warp / CTA top-k selection used by every scan kernel.
//
// A result is one 64-bit key:  (okey(distance) << 32) | chunk id,  okey = order-preserving map
// of the f32 bit pattern. Ascending u64 order == ascending (distance, id), which is the order
// arroy 0.5.0 returns results in and therefore the order src/vectordb/store.rs:459-483 hands to
// its callers (rank = position, src/rerank/mod.rs:57-59). Keys are unique (ids are), so every
// comparison below is strict and selection is deterministic.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace csgpu {

constexpr uint64_t KEY_EMPTY = 0xFFFFFFFFFFFFFFFFull;
constexpr unsigned FULL = 0xFFFFFFFFu;

__host__ __device__ __forceinline__ uint32_t okey_from_bits(uint32_t u)
{
    return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__host__ __device__ __forceinline__ uint32_t bits_from_okey(uint32_t o)
{
    return o ^ ((o >> 31) ? 0x80000000u : 0xFFFFFFFFu);
}
__device__ __forceinline__ uint32_t okey(float d) { return okey_from_bits(__float_as_uint(d)); }
__device__ __forceinline__ uint64_t make_key(float d, uint32_t id)
{
    return ((uint64_t)okey(d) << 32) | id;
}

__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src)
{
    uint32_t lo = __shfl_sync(FULL, (uint32_t)v, src);
    uint32_t hi = __shfl_sync(FULL, (uint32_t)(v >> 32), src);
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint64_t shfl64_xor(uint64_t v, int m)
{
    uint32_t lo = __shfl_xor_sync(FULL, (uint32_t)v, m);
    uint32_t hi = __shfl_xor_sync(FULL, (uint32_t)(v >> 32), m);
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint64_t shfl64_up(uint64_t v, int d)
{
    uint32_t lo = __shfl_up_sync(FULL, (uint32_t)v, d);
    uint32_t hi = __shfl_up_sync(FULL, (uint32_t)(v >> 32), d);
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint64_t umin64(uint64_t a, uint64_t b) { return a < b ? a : b; }
__device__ __forceinline__ uint64_t umax64(uint64_t a, uint64_t b) { return a < b ? b : a; }

// One key per lane -> ascending by lane.
__device__ __forceinline__ uint64_t warp_sort32(uint64_t v, int lane)
{
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            uint64_t o = shfl64_xor(v, j);
            bool up = (lane & k) == 0;
            bool lower = (lane & j) == 0;
            v = (lower == up) ? umin64(v, o) : umax64(v, o);
        }
    }
    return v;
}

// a, b: one key per lane, each ascending by lane -> the 32 smallest of the 64, ascending by lane. min(a[i], b[31 - i]) is a
// bitonic sequence holding exactly those 32; five compare-exchange stages sort it.
__device__ __forceinline__ uint64_t warp_merge_low32(uint64_t a, uint64_t b, int lane)
{
    uint64_t v = umin64(a, shfl64(b, 31 - lane));
#pragma unroll
    for (int j = 16; j > 0; j >>= 1) {
        const uint64_t o = shfl64_xor(v, j);
        v = ((lane & j) == 0) ? umin64(v, o) : umax64(v, o);
    }
    return v;
}

// ----------------------------------------------------------------------------------------
// k <= 32: the warp's best 32 keys live one per lane, ascending; exact threshold after
// every insertion. Insertion is 1 ballot + 2 shuffles; it only runs for rows that beat the
// current k-th best, i.e. ~k*ln(rows_per_warp/k) times per warp.
// ----------------------------------------------------------------------------------------
struct WarpSel32 {
    static constexpr bool CTA_SHARED = false;
    uint64_t v;    // lane's slot
    uint64_t thr;  // warp-uniform: key of the k-th best so far (KEY_EMPTY until k rows seen)
    int km1;

    __device__ __forceinline__ void init(uint32_t k) { v = KEY_EMPTY; thr = KEY_EMPTY; km1 = (int)k - 1; }
    // precondition: key < thr, call is warp-uniform
    __device__ __forceinline__ void insert(uint64_t key, int lane)
    {
        unsigned m = __ballot_sync(FULL, v < key);
        int pos = __popc(m);
        uint64_t up = shfl64_up(v, 1);
        if (lane == pos) v = key;
        else if (lane > pos) v = up;
        thr = shfl64(v, km1);
    }
    __device__ __forceinline__ void flush(int) {}
};

// ----------------------------------------------------------------------------------------
// 32 < k <= 1024: sorted list of kpad (pow2, multiple of 32) keys in shared memory, ping-pong
// buffers, plus a 32-entry pending buffer in registers. When the pending buffer fills it is
// sorted across the warp and merged by rank (merge-path positions; no data-dependent loops
// over the list). The threshold is refreshed at every merge, so it is at most 32 candidates
// stale - that only lets a few extra rows through, never loses one.
// ----------------------------------------------------------------------------------------
struct WarpSelBig {
    static constexpr bool CTA_SHARED = false;
    uint64_t *cur, *nxt;  // [kpad] each, this warp's
    uint64_t pend;
    uint64_t thr;
    int npend;
    uint32_t kpad, km1;

    __device__ __forceinline__ void init(uint64_t *a, uint64_t *b, uint32_t k, uint32_t kpad_, int lane)
    {
        cur = a; nxt = b; kpad = kpad_; km1 = k - 1; pend = KEY_EMPTY; npend = 0; thr = KEY_EMPTY;
        for (uint32_t j = lane; j < kpad; j += 32) cur[j] = KEY_EMPTY;
        __syncwarp();
    }
    __device__ __forceinline__ void insert(uint64_t key, int lane)
    {
        if (lane == npend) pend = key;
        if (++npend == 32) flush(lane);
    }
    __device__ __noinline__ void flush(int lane)
    {
        if (npend == 0) return;
        uint64_t p = (lane < npend) ? pend : KEY_EMPTY;
        p = warp_sort32(p, lane);
        // list elements: new position = j + #{pending < L[j]}
        for (uint32_t j = lane; j < kpad; j += 32) {
            uint64_t l = cur[j];
            int c = 0;
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) {
                uint64_t pv = shfl64(p, c + s - 1);
                if (pv < l) c += s;
            }
            uint64_t pv = shfl64(p, c);  // c <= 31
            if (pv < l) c += 1;
            uint32_t pos = j + (uint32_t)c;
            if (pos < kpad) nxt[pos] = l;
        }
        // pending elements: new position = lane + #{L <= p}   (list first on ties: only EMPTY ties)
        {
            uint32_t c = 0;
            for (uint32_t s = kpad >> 1; s > 0; s >>= 1)
                if (cur[c + s - 1] <= p) c += s;
            if (cur[c] <= p) c += 1;  // c <= kpad-1 here
            uint32_t pos = (uint32_t)lane + c;
            if (pos < kpad) nxt[pos] = p;
        }
        __syncwarp();
        uint64_t *t = cur; cur = nxt; nxt = t;
        thr = cur[km1];
        pend = KEY_EMPTY;
        npend = 0;
        __syncwarp();
    }
};

// CTA-wide ascending bitonic sort of n (pow2) keys in shared memory.
__device__ __forceinline__ void cta_sort(uint64_t *s, uint32_t n)
{
    for (uint32_t k = 2; k <= n; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            __syncthreads();
            for (uint32_t t = threadIdx.x; t < n / 2; t += blockDim.x) {
                uint32_t i = 2 * t - (t & (j - 1));  // index with bit j clear
                uint32_t l = i ^ j;
                uint64_t a = s[i], b = s[l];
                bool up = (i & k) == 0;
                if ((a > b) == up) { s[i] = b; s[l] = a; }
            }
        }
    }
    __syncthreads();
}

// ----------------------------------------------------------------------------------------
// 32 < k <= 1024, CTA-shared: ONE candidate buffer of `cap` (pow2) keys per query for the whole CTA instead of
// a sorted list per warp (8 warps x 2 x kpad keys = 128 KB at k = 1000, i.e. one CTA per SM and 30 % of the
// bandwidth gone). A row that beats the CTA's threshold is appended with one shared-memory atomic; when the
// buffer comes within `slack` of full the whole CTA sorts it (bitonic, in place), keeps the best k and
// publishes the k-th key as the new threshold. The threshold is always the k-th best of a subset of the rows,
// so nothing that belongs to the final top-k is ever rejected; it is merely up to one refill stale.
// Expected refills per CTA ~ k ln(rows_per_cta / k) / (cap - k - slack): a handful.
// ----------------------------------------------------------------------------------------
struct CtaBuf {
    uint64_t *buf;            // [cap] shared
    unsigned *cnt;            // shared: entries appended (never exceeds cap when the caller honours `slack`)
    volatile uint64_t *thr;   // shared: current threshold key (KEY_EMPTY until k entries have been seen)

    // warp-uniform call, key already tested against *thr by the caller
    __device__ __forceinline__ void append(uint64_t key, int lane)
    {
        if (lane == 0) {
            const unsigned pos = atomicAdd(cnt, 1u);
            buf[pos] = key;
        }
    }
};

// CTA-wide: sort buf[0..cap) (entries >= *cnt are holes), keep the best k, publish the threshold.
// Every thread of the CTA must call it (it synchronises).
__device__ __forceinline__ void cta_buf_compact(uint64_t *buf, unsigned *cnt, volatile uint64_t *thr, uint32_t cap, uint32_t k)
{
    __syncthreads();
    const unsigned n = min(*cnt, cap);
    // sort only as many keys as there are (round 2): on a small corpus a CTA's buffer holds a few hundred keys when it finishes,
    // and the last CTA's first compaction one key per list — a 512-key sort instead of a cap-sized (1024+) one. Slots up to k
    // are padded so that buf[0..k) is always "best k ascending, KEY_EMPTY padded".
    uint32_t spad = 32;
    while (spad < n) spad <<= 1;                              // <= cap: cap is a power of two >= n
    const uint32_t fill_to = max(spad, min(k, cap));
    for (uint32_t t = n + threadIdx.x; t < fill_to; t += blockDim.x) buf[t] = KEY_EMPTY;
    cta_sort(buf, spad);   // leading + trailing __syncthreads inside
    if (threadIdx.x == 0) {
        *cnt = min(n, k);
        if (n >= k) *thr = buf[k - 1];
    }
    __syncthreads();
}

// CTA-uniform: offer `total` keys (key = load(t)) to the buffer, all threads in lockstep, compacting whenever
// fewer than blockDim.x free slots remain. Precondition: *cnt + blockDim.x <= cap.
template <class F>
__device__ __forceinline__ void cta_buf_stream(CtaBuf &cb, uint32_t cap, uint32_t k, uint64_t total, F load)
{
    for (uint64_t b0 = 0; b0 < total; b0 += blockDim.x) {
        const uint64_t t = b0 + threadIdx.x;
        const uint64_t key = t < total ? load(t) : KEY_EMPTY;
        if (key < *cb.thr) cb.buf[atomicAdd(cb.cnt, 1u)] = key;
        if (__syncthreads_or(*reinterpret_cast<volatile unsigned *>(cb.cnt) + blockDim.x > cap))
            cta_buf_compact(cb.buf, cb.cnt, cb.thr, cap, k);
    }
}

// Selector facade over CtaBuf with the interface the scan loops use (thr / insert / sync_point).
struct CtaSel {
    static constexpr bool CTA_SHARED = true;
    uint64_t thr;   // register copy of the CTA threshold, refreshed at sync points
    CtaBuf cb;
    uint32_t cap, k;

    __device__ __forceinline__ void reset()   // CTA-uniform
    {
        __syncthreads();
        if (threadIdx.x == 0) { *cb.cnt = 0; *cb.thr = KEY_EMPTY; }
        __syncthreads();
        thr = KEY_EMPTY;
    }
    __device__ __forceinline__ void insert(uint64_t key, int lane) { cb.append(key, lane); }
    // CTA-uniform; `slack` = the most keys the CTA can append before the next sync point
    __device__ __forceinline__ void sync_point(uint32_t slack)
    {
        if (__syncthreads_or(*reinterpret_cast<volatile unsigned *>(cb.cnt) + slack > cap))
            cta_buf_compact(cb.buf, cb.cnt, cb.thr, cap, k);
        thr = *cb.thr;
    }
    // CTA-uniform: buf[0..k) = best k ascending (KEY_EMPTY padded)
    __device__ __forceinline__ void finish() { cta_buf_compact(cb.buf, cb.cnt, cb.thr, cap, k); }
};

__host__ __device__ __forceinline__ uint32_t ctabuf_cap(uint32_t k) { uint32_t p = 1024; while (p < 2 * k + 512) p <<= 1; return p; }

__host__ __device__ __forceinline__ uint32_t pow2_at_least(uint32_t x, uint32_t lo)
{
    uint32_t p = lo;
    while (p < x) p <<= 1;
    return p;
}

}  // namespace csgpu
