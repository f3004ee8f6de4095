/* Latency of the C ABI itself (include/csgpu.h), no Python in the loop: what a Rust / C host pays per csgpu_search call with
 * host pointers in and out, and per csgpu_search_batch call of query variants. Plain C on purpose: it is also the proof that
 * the header compiles as C. Prints one JSON line per corpus size. A diagnostic beside bench.py (extras.c_abi_latency).
 *   gcc -O2 -I include tools/bench_c_abi.c -L codesearch_b200 -lcsgpu -Wl,-rpath,... -o build/bench_c_abi
 *   build/bench_c_abi [rows ...]                                                                                        */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "csgpu.h"

static double now_us(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}
static int cmp_d(const void *a, const void *b) { const double x = *(const double *)a, y = *(const double *)b; return x < y ? -1 : x > y; }
#define CK(x) do { int rc_ = (x); if (rc_) { fprintf(stderr, "%s -> %d: %s\n", #x, rc_, csgpu_last_error()); return 1; } } while (0)

int main(int argc, char **argv)
{
    const uint32_t dim = 384, k = 10, NQ = 64;
    uint64_t sizes[8] = {100000, 1000000};
    int n_sizes = 2;
    if (argc > 1) { n_sizes = 0; for (int i = 1; i < argc && n_sizes < 8; ++i) sizes[n_sizes++] = strtoull(argv[i], NULL, 10); }
    float *qs = (float *)malloc((size_t)NQ * dim * sizeof(float));
    uint32_t ids[16 * 256], ns[16];
    float dist[16 * 256];
    for (int s = 0; s < n_sizes; ++s) {
        csgpu_index *ix = NULL;
        CK(csgpu_create(&ix, dim, CSGPU_DTYPE_F32, NULL, 1));
        CK(csgpu_reserve(ix, sizes[s]));
        CK(csgpu_append_synthetic(ix, 1234, 0, sizes[s], 0));
        CK(csgpu_build(ix));
        CK(csgpu_synth_rows_host(ix, 4321, 0, NQ, qs));
        uint32_t n = 0;
        for (int i = 0; i < 50; ++i) CK(csgpu_search(ix, qs + (size_t)(i % NQ) * dim, dim, k, ids, dist, &n));
        enum { REPS = 2000 };
        static double lat[REPS];
        const double t0 = now_us();
        for (int i = 0; i < REPS; ++i) {
            const double a = now_us();
            CK(csgpu_search(ix, qs + (size_t)(i % NQ) * dim, dim, k, ids, dist, &n));
            lat[i] = now_us() - a;
        }
        const double mean = (now_us() - t0) / REPS;
        qsort(lat, REPS, sizeof(double), cmp_d);
        csgpu_stats_t st;
        CK(csgpu_stats(ix, &st));
        printf("{\"rows\": %llu, \"dim\": %u, \"k\": %u, \"csgpu_search_us\": {\"mean\": %.1f, \"p50\": %.1f, \"p99\": %.1f}, \"device_us_last\": %.1f, \"qps_one_thread\": %.0f",
               (unsigned long long)sizes[s], dim, k, mean, lat[REPS / 2], lat[REPS * 99 / 100], st.last_search_us, 1e6 / mean);
        const uint32_t variants[4] = {2, 4, 9, 16};
        printf(", \"csgpu_search_batch_us_by_query_variants\": {");
        for (int v = 0; v < 4; ++v) {
            const uint32_t b = variants[v];
            for (int i = 0; i < 10; ++i) CK(csgpu_search_batch(ix, qs, dim, b, k, ids, dist, ns));
            const double b0 = now_us();
            for (int i = 0; i < 500; ++i) CK(csgpu_search_batch(ix, qs + (size_t)(i % 3) * dim, dim, b, k, ids, dist, ns));
            printf("%s\"%u\": %.1f", v ? ", " : "", b, (now_us() - b0) / 500);
        }
        printf("}, \"top_ids_query0\": [");
        CK(csgpu_search(ix, qs, dim, k, ids, dist, &n));
        for (uint32_t i = 0; i < n; ++i) printf("%s%u", i ? ", " : "", ids[i]);
        printf("]}\n");
        fflush(stdout);
        csgpu_destroy(ix);
    }
    free(qs);
    return 0;
}
