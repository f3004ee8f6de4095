/*
 * oracle.c — CPU restatement of the reference's vector-retrieval hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT. Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it, and only as
 * the checker / the timed CPU baseline. The product path (libcsgpu.so) never
 * links, loads or calls anything in this directory.
 *
 * PARITY STATUS: "parity unpinned" for (i) the distance scale (1-cos)/2 and
 * (ii) arroy's ANN recall. The reference (Rust + arroy 0.5.0 + LMDB) cannot be
 * compiled or run in this environment (no cargo/rustc, crate sources not
 * vendored), so this file restates:
 *
 *   - the reference's own exact scan arithmetic
 *       /root/reference/examples/benchmark_models.rs:323-328  (cosine_similarity)
 *       /root/reference/examples/benchmark_models.rs:155-165  (linear scan, strict >)
 *   - the zero-norm convention of the reference's test helper
 *       /root/reference/src/embed/batch.rs:316-324
 *   - arroy 0.5.0's published Cosine distance and result ordering
 *       (Cargo.lock:162-165 pins arroy 0.5.0; call site src/vectordb/store.rs:446-459)
 *       built_distance = pn*qn != 0 ? (1 - dot/(pn*qn))/2 : 0 ;  results ascending
 *       (distance, item id), min(k, N) of them
 *   - the score conversion  src/vectordb/store.rs:477-478  (score = 1 - distance)
 *
 * and is pinned against every known-answer the reference's tests hold for the
 * path (tests/test_oracle.py: store.rs:846-893, embed/batch.rs:326-340).
 *
 * Build: see oracle/Makefile  (gcc -O3 -march=native -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------- */
/* Synthetic corpus generator (shared spec with codesearch_b200/csrc/synth.cuh) */
/* ------------------------------------------------------------------------- */
/*
 * Counter-based so any row of any shard is regenerable anywhere:
 *   (x0,x1,x2,x3) = Philox4x32-10(counter = {row_lo, row_hi, col/4, 0},
 *                                 key     = {seed_lo, seed_hi})
 *   value[col + j] = (float)(byte0(xj) + byte1(xj) + byte2(xj) + byte3(xj) - 510)
 * i.e. a centred Irwin-Hall(4) over bytes: integer-exact (no libm), so the CPU
 * and the GPU produce bit-identical raw rows. Rows are NOT unit length; the
 * store normalises at build, the oracle divides by the norms (as arroy does).
 */
#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u

void cs_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)PHILOX_M0 * c0;
        uint64_t p1 = (uint64_t)PHILOX_M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += PHILOX_W0; k1 += PHILOX_W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static inline float synth_val(uint32_t x)
{
    int s = (int)(x & 0xFF) + (int)((x >> 8) & 0xFF) + (int)((x >> 16) & 0xFF) + (int)(x >> 24);
    return (float)(s - 510);
}

/* rows [first_row, first_row+n) x dim (dim % 4 == 0), row-major into out. */
void cs_synth_rows(uint64_t seed, uint64_t first_row, uint64_t n, uint32_t dim, float *out)
{
    const uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        uint64_t row = first_row + (uint64_t)i;
        float *o = out + (size_t)i * dim;
        for (uint32_t c4 = 0; c4 < dim / 4; ++c4) {
            uint32_t ctr[4] = { (uint32_t)row, (uint32_t)(row >> 32), c4, 0u };
            uint32_t x[4];
            cs_philox4x32_10(ctr, key, x);
            o[4 * c4 + 0] = synth_val(x[0]);
            o[4 * c4 + 1] = synth_val(x[1]);
            o[4 * c4 + 2] = synth_val(x[2]);
            o[4 * c4 + 3] = synth_val(x[3]);
        }
    }
}

/* ------------------------------------------------------------------------- */
/* Scalar arithmetic of the reference                                         */
/* ------------------------------------------------------------------------- */

/* examples/benchmark_models.rs:323-328 — three sequential f32 sums, dot/(|a||b|).
 * No zero guard there (0/0 = NaN), exactly as written. */
float cs_ref_cosine_similarity(const float *a, const float *b, uint32_t d)
{
    volatile float dot = 0.0f, ma = 0.0f, mb = 0.0f; /* volatile: forbid reassociation/FMA contraction */
    for (uint32_t i = 0; i < d; ++i) dot = dot + a[i] * b[i];
    for (uint32_t i = 0; i < d; ++i) ma = ma + a[i] * a[i];
    for (uint32_t i = 0; i < d; ++i) mb = mb + b[i] * b[i];
    return dot / (sqrtf(ma) * sqrtf(mb));
}

/* src/embed/batch.rs:316-324 — same, with the zero-magnitude => 0.0 guard. */
float cs_ref_cosine_similarity_guarded(const float *a, const float *b, uint32_t d)
{
    volatile float dot = 0.0f, ma = 0.0f, mb = 0.0f;
    for (uint32_t i = 0; i < d; ++i) dot = dot + a[i] * b[i];
    for (uint32_t i = 0; i < d; ++i) ma = ma + a[i] * a[i];
    for (uint32_t i = 0; i < d; ++i) mb = mb + b[i] * b[i];
    float na = sqrtf(ma), nb = sqrtf(mb);
    if (na == 0.0f || nb == 0.0f) return 0.0f;
    return dot / (na * nb);
}

/* arroy 0.5.0 Cosine::built_distance restated in strict sequential f32:
 * norm = sqrt(dot(v,v)) per leaf; pn*qn != 0 ? (1 - pq/(pn*qn))/2 : 0. */
float cs_ref_distance_f32(const float *p, const float *q, uint32_t d)
{
    volatile float pq = 0.0f, pp = 0.0f, qq = 0.0f;
    for (uint32_t i = 0; i < d; ++i) pq = pq + p[i] * q[i];
    for (uint32_t i = 0; i < d; ++i) pp = pp + p[i] * p[i];
    for (uint32_t i = 0; i < d; ++i) qq = qq + q[i] * q[i];
    float pnqn = sqrtf(pp) * sqrtf(qq);
    if (pnqn != 0.0f) {
        float c = pq / pnqn;
        return (1.0f - c) / 2.0f;
    }
    return 0.0f;
}

/* Order-independent referee: f64 accumulation from the f32 inputs. */
static inline double dot64(const float *a, const float *b, uint32_t d)
{
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    uint32_t i = 0;
    for (; i + 4 <= d; i += 4) {
        s0 += (double)a[i] * b[i];
        s1 += (double)a[i + 1] * b[i + 1];
        s2 += (double)a[i + 2] * b[i + 2];
        s3 += (double)a[i + 3] * b[i + 3];
    }
    for (; i < d; ++i) s0 += (double)a[i] * b[i];
    return (s0 + s1) + (s2 + s3);
}

double cs_ref_distance_f64(const float *p, const float *q, uint32_t d)
{
    double pq = dot64(p, q, d), pp = dot64(p, p, d), qq = dot64(q, q, d);
    double pnqn = sqrt(pp) * sqrt(qq);
    if (pnqn != 0.0) return (1.0 - pq / pnqn) / 2.0;
    return 0.0;
}

/* ------------------------------------------------------------------------- */
/* Top-k scan: ascending (distance, id), min(k, live rows) results             */
/* ------------------------------------------------------------------------- */

typedef struct { double d; uint32_t id; } cand_t;

/* a "worse" than b  <=>  (a.d, a.id) > (b.d, b.id); NaN sorts last (OrderedFloat). */
static inline int cand_worse(cand_t a, cand_t b)
{
    int an = isnan(a.d), bn = isnan(b.d);
    if (an || bn) { if (an != bn) return an; return a.id > b.id; }
    if (a.d != b.d) return a.d > b.d;
    return a.id > b.id;
}

/* bounded max-heap on "worse": heap[0] is the worst kept candidate. */
static void heap_sift_down(cand_t *h, uint32_t n, uint32_t i)
{
    for (;;) {
        uint32_t l = 2 * i + 1, r = l + 1, m = i;
        if (l < n && cand_worse(h[l], h[m])) m = l;
        if (r < n && cand_worse(h[r], h[m])) m = r;
        if (m == i) return;
        cand_t t = h[i]; h[i] = h[m]; h[m] = t; i = m;
    }
}
static void heap_sift_up(cand_t *h, uint32_t i)
{
    while (i) {
        uint32_t p = (i - 1) / 2;
        if (!cand_worse(h[i], h[p])) return;
        cand_t t = h[i]; h[i] = h[p]; h[p] = t; i = p;
    }
}
static inline void heap_offer(cand_t *h, uint32_t *n, uint32_t k, cand_t c)
{
    if (*n < k) { h[*n] = c; heap_sift_up(h, (*n)++); }
    else if (k && cand_worse(h[0], c)) { h[0] = c; heap_sift_down(h, k, 0); }
}
static int cand_cmp(const void *a, const void *b)
{
    cand_t x = *(const cand_t *)a, y = *(const cand_t *)b;
    if (cand_worse(x, y)) return 1;
    if (cand_worse(y, x)) return -1;
    return 0;
}

/*
 * mode 0: strict sequential f32 arithmetic (faithful to benchmark_models.rs sums
 *         + arroy's f32 distance); ordering on that f32 distance.
 * mode 1: f64 accumulation, distance rounded to f32 for ordering and output
 *         (the order-independent referee the CUDA path is compared with).
 * bitmap: NULL, or bit (id) set <=> chunk id allowed (filtered variant,
 *         exact top-k over passing rows); ids >= n_bits are excluded.
 * ids:    NULL => id = row index.
 * out_dist64 (optional): the unrounded f64 distance of each result (mode 1),
 *         used by the tests to detect near-ties.
 * Returns the number of results written = min(k, passing rows).
 */
uint32_t cs_oracle_search(const float *rows, const uint32_t *ids, uint64_t n, uint32_t d,
                          const float *q, uint32_t k, int mode,
                          const uint64_t *bitmap, uint64_t n_bits,
                          uint32_t *out_ids, float *out_dist, double *out_dist64)
{
    if (k == 0 || n == 0) return 0;
    int nt = 1;
#ifdef _OPENMP
    nt = omp_get_max_threads();
#endif
    cand_t *heaps = (cand_t *)malloc((size_t)nt * k * sizeof(cand_t));
    uint32_t *cnt = (uint32_t *)calloc((size_t)nt, sizeof(uint32_t));
#pragma omp parallel
    {
        int t = 0;
#ifdef _OPENMP
        t = omp_get_thread_num();
#endif
        cand_t *h = heaps + (size_t)t * k;
        uint32_t hn = 0;
#pragma omp for schedule(static)
        for (int64_t i = 0; i < (int64_t)n; ++i) {
            uint32_t id = ids ? ids[i] : (uint32_t)i;
            if (bitmap) {
                if ((uint64_t)id >= n_bits) continue;
                if (!((bitmap[id >> 6] >> (id & 63)) & 1ull)) continue;
            }
            const float *p = rows + (size_t)i * d;
            cand_t c;
            c.id = id;
            if (mode == 0) c.d = (double)cs_ref_distance_f32(p, q, d);
            else           c.d = cs_ref_distance_f64(p, q, d);
            if (mode == 1) {
                /* order on the f32-rounded distance, like arroy's OrderedFloat<f32> */
                cand_t r = c; r.d = (double)(float)c.d;
                /* keep the f64 value recoverable: recomputed at the end */
                heap_offer(h, &hn, k, r);
            } else {
                heap_offer(h, &hn, k, c);
            }
        }
        cnt[t] = hn;
    }
    uint32_t tot = 0;
    for (int t = 0; t < nt; ++t) {
        if (t && cnt[t]) memmove(heaps + tot, heaps + (size_t)t * k, cnt[t] * sizeof(cand_t));
        tot += cnt[t];
    }
    qsort(heaps, tot, sizeof(cand_t), cand_cmp);
    uint32_t m = tot < k ? tot : k;
    for (uint32_t i = 0; i < m; ++i) {
        out_ids[i] = heaps[i].id;
        out_dist[i] = (float)heaps[i].d;
    }
    if (out_dist64) {
        /* id -> row lookup only needed when ids != NULL; do a linear pass for the m winners */
        for (uint32_t i = 0; i < m; ++i) out_dist64[i] = (double)out_dist[i];
        if (mode == 1) {
            if (!ids) {
                for (uint32_t i = 0; i < m; ++i)
                    out_dist64[i] = cs_ref_distance_f64(rows + (size_t)out_ids[i] * d, q, d);
            } else {
                for (uint64_t r = 0; r < n; ++r)
                    for (uint32_t i = 0; i < m; ++i)
                        if (ids[r] == out_ids[i])
                            out_dist64[i] = cs_ref_distance_f64(rows + (size_t)r * d, q, d);
            }
        }
    }
    free(heaps); free(cnt);
    return m;
}

/* Batched: b queries [b,d] against the same rows; outputs [b,k], out_n[b]. */
void cs_oracle_search_batch(const float *rows, const uint32_t *ids, uint64_t n, uint32_t d,
                            const float *q, uint32_t b, uint32_t k, int mode,
                            uint32_t *out_ids, float *out_dist, uint32_t *out_n)
{
    for (uint32_t j = 0; j < b; ++j)
        out_n[j] = cs_oracle_search(rows, ids, n, d, q + (size_t)j * d, k, mode, NULL, 0,
                                    out_ids + (size_t)j * k, out_dist + (size_t)j * k, NULL);
}

/*
 * Streaming oracle over the SYNTHETIC corpus (never materialised): rows
 * [first_row, first_row+n) of cs_synth_rows(seed,...), id = global row index,
 * scored in f64 against b queries at once. Used for the full-size (10M+) parity
 * checks on the GPU box, where 15 GB of host rows is not wanted.
 */
void cs_oracle_search_synth(uint64_t seed, uint64_t first_row, uint64_t n, uint32_t d,
                            const float *q, uint32_t b, uint32_t k,
                            uint32_t *out_ids, float *out_dist, double *out_dist64, uint32_t *out_n)
{
    int nt = 1;
#ifdef _OPENMP
    nt = omp_get_max_threads();
#endif
    const uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    cand_t *heaps = (cand_t *)malloc((size_t)nt * b * k * sizeof(cand_t));
    uint32_t *cnt = (uint32_t *)calloc((size_t)nt * b, sizeof(uint32_t));
    double *qq = (double *)malloc(b * sizeof(double));
    for (uint32_t j = 0; j < b; ++j) qq[j] = dot64(q + (size_t)j * d, q + (size_t)j * d, d);
#pragma omp parallel
    {
        int t = 0;
#ifdef _OPENMP
        t = omp_get_thread_num();
#endif
        float *row = (float *)malloc(d * sizeof(float));
#pragma omp for schedule(static)
        for (int64_t i = 0; i < (int64_t)n; ++i) {
            uint64_t r = first_row + (uint64_t)i;
            for (uint32_t c4 = 0; c4 < d / 4; ++c4) {
                uint32_t ctr[4] = { (uint32_t)r, (uint32_t)(r >> 32), c4, 0u }, x[4];
                cs_philox4x32_10(ctr, key, x);
                row[4 * c4] = synth_val(x[0]); row[4 * c4 + 1] = synth_val(x[1]);
                row[4 * c4 + 2] = synth_val(x[2]); row[4 * c4 + 3] = synth_val(x[3]);
            }
            double pp = dot64(row, row, d);
            for (uint32_t j = 0; j < b; ++j) {
                double pq = dot64(row, q + (size_t)j * d, d);
                double pnqn = sqrt(pp) * sqrt(qq[j]);
                cand_t c; c.id = (uint32_t)r;
                c.d = (pnqn != 0.0) ? (double)(float)((1.0 - pq / pnqn) / 2.0) : 0.0;
                heap_offer(heaps + ((size_t)t * b + j) * k, &cnt[(size_t)t * b + j], k, c);
            }
        }
        free(row);
    }
    cand_t *all = (cand_t *)malloc((size_t)nt * k * sizeof(cand_t));
    float *row = (float *)malloc(d * sizeof(float));
    for (uint32_t j = 0; j < b; ++j) {
        uint32_t tot = 0;
        for (int t = 0; t < nt; ++t) {
            memcpy(all + tot, heaps + ((size_t)t * b + j) * k, cnt[(size_t)t * b + j] * sizeof(cand_t));
            tot += cnt[(size_t)t * b + j];
        }
        qsort(all, tot, sizeof(cand_t), cand_cmp);
        uint32_t m = tot < k ? tot : k;
        out_n[j] = m;
        for (uint32_t i = 0; i < m; ++i) {
            out_ids[(size_t)j * k + i] = all[i].id;
            out_dist[(size_t)j * k + i] = (float)all[i].d;
            if (out_dist64) {
                cs_synth_rows(seed, all[i].id, 1, d, row);
                out_dist64[(size_t)j * k + i] = cs_ref_distance_f64(row, q + (size_t)j * d, d);
            }
        }
    }
    free(row); free(all); free(heaps); free(cnt); free(qq);
}

/* ------------------------------------------------------------------------- */
/* CPU baseline kernels (timed by bench.py; same arithmetic, all host threads)  */
/* ------------------------------------------------------------------------- */
/*
 * The reference's exact scan generalised to top-k, in f32, the way a compiled
 * Rust release build would run it (vectorisable f32 sums), OpenMP over rows.
 * This is "reference exact-scan restatement", NOT arroy. Norms are recomputed
 * per row, as benchmark_models.rs:323-328 does on every call.
 */
uint32_t cs_cpu_baseline_search(const float *rows, uint64_t n, uint32_t d, const float *q,
                                uint32_t k, uint32_t *out_ids, float *out_dist)
{
    int nt = 1;
#ifdef _OPENMP
    nt = omp_get_max_threads();
#endif
    cand_t *heaps = (cand_t *)malloc((size_t)nt * k * sizeof(cand_t));
    uint32_t *cnt = (uint32_t *)calloc((size_t)nt, sizeof(uint32_t));
    float qq = 0.0f;
    for (uint32_t i = 0; i < d; ++i) qq += q[i] * q[i];
    const float qn = sqrtf(qq);
#pragma omp parallel
    {
        int t = 0;
#ifdef _OPENMP
        t = omp_get_thread_num();
#endif
        cand_t *h = heaps + (size_t)t * k;
        uint32_t hn = 0;
#pragma omp for schedule(static)
        for (int64_t i = 0; i < (int64_t)n; ++i) {
            const float *p = rows + (size_t)i * d;
            float pq = 0.0f, pp = 0.0f;
#pragma omp simd reduction(+ : pq, pp)
            for (uint32_t c = 0; c < d; ++c) { pq += p[c] * q[c]; pp += p[c] * p[c]; }
            float pnqn = sqrtf(pp) * qn;
            cand_t c; c.id = (uint32_t)i;
            c.d = (pnqn != 0.0f) ? (double)((1.0f - pq / pnqn) / 2.0f) : 0.0;
            heap_offer(h, &hn, k, c);
        }
        cnt[t] = hn;
    }
    uint32_t tot = 0;
    for (int t = 0; t < nt; ++t) {
        if (t && cnt[t]) memmove(heaps + tot, heaps + (size_t)t * k, cnt[t] * sizeof(cand_t));
        tot += cnt[t];
    }
    qsort(heaps, tot, sizeof(cand_t), cand_cmp);
    uint32_t m = tot < k ? tot : k;
    for (uint32_t i = 0; i < m; ++i) { out_ids[i] = heaps[i].id; out_dist[i] = (float)heaps[i].d; }
    free(heaps); free(cnt);
    return m;
}

int cs_oracle_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void cs_oracle_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
