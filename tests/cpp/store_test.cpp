// store_test.cpp — the reference's own VectorStore tests (src/vectordb/store.rs:826-1029) restated against the C++
// host mirror (include/csgpu_store.hpp) over libcsgpu.so, plus the guard messages (store.rs:432-444) and the additive
// methods. Built by __graft_entry__.build() (g++), run by tests/test_cpp_store.py: with a GPU it must print
// "ALL PASSED"; with `--expect-no-gpu` it checks that opening a store fails loudly (there is no CPU fallback).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <string>
#include <unistd.h>
#include <vector>

#include "csgpu_store.hpp"

using namespace codesearch;

static int g_fail = 0;
#define CHECK(cond)                                                                       \
    do {                                                                                  \
        if (!(cond)) { std::printf("  CHECK FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); ++g_fail; } \
    } while (0)

static std::string temp_db(const char *name)
{
    char tmpl[] = "/tmp/csgpu_store_test_XXXXXX";
    const char *d = mkdtemp(tmpl);
    return std::string(d ? d : "/tmp") + "/" + name;
}

static Chunk fn_chunk(const std::string &content, size_t a, size_t b, const std::string &path)
{
    return Chunk(content, a, b, "Function", path);
}

// store.rs:833-844
static void test_vector_store_creation()
{
    VectorStore store = VectorStore::create(temp_db("test.db"), 384);
    CHECK(store.dimensions == 384);
    CHECK(!store.is_indexed());
}

// store.rs:846-893
static void test_insert_and_search()
{
    VectorStore store = VectorStore::create(temp_db("test.db"), 4);
    std::vector<EmbeddedChunk> chunks = {
        EmbeddedChunk(fn_chunk("fn authenticate() {}", 0, 1, "auth.rs"), {1.0f, 0.0f, 0.0f, 0.0f}),   // close to query
        EmbeddedChunk(fn_chunk("fn calculate() {}", 2, 3, "math.rs"), {0.0f, 1.0f, 0.0f, 0.0f}),      // far from query
    };
    CHECK(store.insert_chunks(chunks) == 2);
    store.build_index();
    CHECK(store.is_indexed());
    auto results = store.search({0.9f, 0.1f, 0.0f, 0.0f}, 2);
    CHECK(results.size() == 2);
    CHECK(results[0].content.find("authenticate") != std::string::npos);
    CHECK(results[0].score > results[1].score);
    // the numeric values the reference's arithmetic implies: distance = (1 - cos) / 2, score = 1 - distance
    CHECK(std::fabs(results[0].distance - 0.0030581355f) < 1e-6f && std::fabs(results[1].distance - 0.44478422f) < 1e-6f);
    CHECK(std::fabs(results[0].score - 0.99694186f) < 1e-6f && std::fabs(results[1].score - 0.5552158f) < 1e-6f);
}

// store.rs:895-938
static void test_stats()
{
    VectorStore store = VectorStore::create(temp_db("test.db"), 4);
    store.insert_chunks({EmbeddedChunk(fn_chunk("fn test1() {}", 0, 1, "file1.rs"), {1.0f, 0.0f, 0.0f, 0.0f}),
                         EmbeddedChunk(fn_chunk("fn test2() {}", 0, 1, "file2.rs"), {0.0f, 1.0f, 0.0f, 0.0f})});
    store.build_index();
    StoreStats s = store.stats();
    CHECK(s.total_chunks == 2);
    CHECK(s.total_files == 2);
    CHECK(s.indexed);
    CHECK(s.dimensions == 4);
}

// store.rs:940-970
static void test_clear()
{
    VectorStore store = VectorStore::create(temp_db("test.db"), 4);
    store.insert_chunks({EmbeddedChunk(fn_chunk("fn test() {}", 0, 1, "test.rs"), {1.0f, 0.0f, 0.0f, 0.0f})});
    store.build_index();
    CHECK(store.stats().total_chunks == 1);
    store.clear();
    CHECK(store.stats().total_chunks == 0);
    CHECK(!store.stats().indexed);
}

// store.rs:972-997
static void test_get_chunk()
{
    VectorStore store = VectorStore::create(temp_db("test.db"), 4);
    store.insert_chunks({EmbeddedChunk(fn_chunk("fn test() {}", 0, 1, "test.rs"), {1.0f, 0.0f, 0.0f, 0.0f})});
    auto md = store.get_chunk(0);
    CHECK(md.has_value());
    CHECK(md->content == "fn test() {}");
    CHECK(md->path == "test.rs");
}

// store.rs:999-1028
static void test_persistence()
{
    const std::string db = temp_db("test.db");
    {
        VectorStore store = VectorStore::create(db, 4);
        store.insert_chunks({EmbeddedChunk(fn_chunk("fn test() {}", 0, 1, "test.rs"), {1.0f, 0.0f, 0.0f, 0.0f})});
        store.build_index();
    }
    {
        VectorStore store = VectorStore::create(db, 4);
        CHECK(store.stats().total_chunks == 1);
        CHECK(store.get_chunk(0).has_value());
        CHECK(store.is_indexed());
        auto r = store.search({1.0f, 0.0f, 0.0f, 0.0f}, 1);          // the hydrated device index answers
        CHECK(r.size() == 1 && r[0].id == 0 && r[0].distance == 0.0f);
        auto ids = store.insert_chunks_with_ids({EmbeddedChunk(fn_chunk("fn next() {}", 2, 3, "next.rs"), {0.f, 1.f, 0.f, 0.f})});
        CHECK(ids.size() == 1 && ids[0] == 1);                        // next_id = last key + 1 (store.rs:141-144)
    }
}

// incremental reindex (delete a file's chunks, re-insert, reopen): the persisted file table keeps the file ids the row
// tags were written with (round-1 advisor finding: re-deriving them from the surviving chunks renumbers files)
static void test_file_table_reopen()
{
    const std::string db = temp_db("files.db");
    std::vector<uint32_t> a2, b;
    {
        VectorStore store = VectorStore::create(db, 4);
        auto a = store.insert_chunks_with_ids({EmbeddedChunk(fn_chunk("fn a() {}", 0, 1, "src/a.rs"), {1.f, 0.f, 0.f, 0.f})});
        b = store.insert_chunks_with_ids({EmbeddedChunk(fn_chunk("def b(): pass", 0, 1, "lib/b.py"), {0.f, 1.f, 0.f, 0.f})});
        store.build_index();
        CHECK(store.delete_chunks(a) == 1);
        a2 = store.insert_chunks_with_ids({EmbeddedChunk(fn_chunk("fn a2() {}", 0, 1, "src/a.rs"), {0.9f, 0.1f, 0.f, 0.f})});
        store.build_index();
    }
    {
        VectorStore store = VectorStore::create(db, 4);
        TagFilter fa, fb;
        fa.path_prefix = "src/";
        fb.path_prefix = "lib/";
        auto ra = store.search_tagged({0.5f, 0.5f, 0.f, 0.f}, 5, fa);
        auto rb = store.search_tagged({0.5f, 0.5f, 0.f, 0.f}, 5, fb);
        CHECK(ra.size() == 1 && ra[0].id == a2[0] && ra[0].path == "src/a.rs");
        CHECK(rb.size() == 1 && rb[0].id == b[0] && rb[0].path == "lib/b.py");
        store.delete_chunks(a2);   // every chunk of the first file gone: the second keeps its id
        store.build_index();
    }
    {
        VectorStore store = VectorStore::create(db, 4);
        TagFilter fa, fb;
        fa.path_prefix = "src/";
        fb.path_prefix = "lib/";
        CHECK(store.search_tagged({0.5f, 0.5f, 0.f, 0.f}, 5, fa).empty());
        auto rb = store.search_tagged({0.5f, 0.5f, 0.f, 0.f}, 5, fb);
        CHECK(rb.size() == 1 && rb[0].id == b[0]);
    }
}

// one process, several devices (here: two shards on device 0, which csgpu_create accepts — the one-GPU box covers the whole
// multi-shard path): search() = one fused launch per shard, identical results, snapshot reopens under another device count
static void test_multi_device_store()
{
    const std::string db = temp_db("multi.db");
    std::vector<EmbeddedChunk> chunks;
    for (int i = 0; i < 9000; ++i) {   // >= 4096 per device: one insert is split evenly over the shards
        std::vector<float> e(8);
        for (int c = 0; c < 8; ++c) e[c] = std::sin(0.37f * (float)(i + 1) * (float)(c + 1)) + (c == i % 8 ? 0.5f : 0.f);
        chunks.push_back(EmbeddedChunk(fn_chunk("fn f" + std::to_string(i) + "() {}", (size_t)i, (size_t)i + 1, "src/m" + std::to_string(i / 40) + ".rs"), e));
    }
    std::vector<float> q = {0.3f, -0.2f, 0.9f, 0.1f, 0.0f, 0.4f, -0.7f, 0.2f};
    std::vector<SearchResult> one_r, two_r;
    {
        VectorStore one = VectorStore::create("", 8);
        one.insert_chunks(chunks);
        one.build_index();
        one_r = one.search(q, 50);
    }
    {
        VectorStore two = VectorStore::create(db, 8, {0, 0});
        two.insert_chunks(chunks);
        two.build_index();
        csgpu_stats_t s;
        CHECK(csgpu_stats(two.handle(), &s) == CSGPU_OK && s.n_devices == 2 && s.rows_per_device[0] > 0 && s.rows_per_device[1] > 0);
        const uint64_t l0 = csgpu_kernel_launches();
        two_r = two.search(q, 50);
        CHECK(csgpu_kernel_launches() - l0 == 2);   // one fused scan per shard, no merge launch
    }
    CHECK(one_r.size() == 50 && two_r.size() == 50);
    for (size_t i = 0; i < one_r.size() && i < two_r.size(); ++i)
        CHECK(one_r[i].id == two_r[i].id && one_r[i].distance == two_r[i].distance && one_r[i].path == two_r[i].path);
    {
        VectorStore again = VectorStore::create(db, 8);   // the two-shard snapshot under one device
        auto r = again.search(q, 50);
        CHECK(r.size() == 50);
        for (size_t i = 0; i < r.size() && i < one_r.size(); ++i) CHECK(r[i].id == one_r[i].id && r[i].distance == one_r[i].distance);
    }
}

// guards, store.rs:432-444: literal messages
static void test_guards()
{
    VectorStore store = VectorStore::create("", 4);
    store.insert_chunks({EmbeddedChunk(fn_chunk("fn a() {}", 0, 1, "a.rs"), {1.0f, 0.0f, 0.0f, 0.0f})});
    try {
        store.search({1.0f, 0.0f, 0.0f, 0.0f}, 1);
        CHECK(!"search on an unbuilt index must fail");
    } catch (const Error &e) {
        CHECK(e.code() == CSGPU_ERR_NOT_BUILT);
        CHECK(std::string(e.what()) == "Index not built. Call build_index() after inserting chunks.");
    }
    store.build_index();
    try {
        store.search({1.0f, 0.0f, 0.0f}, 1);
        CHECK(!"dimension mismatch must fail");
    } catch (const Error &e) {
        CHECK(e.code() == CSGPU_ERR_DIM);
        CHECK(std::string(e.what()) == "Query embedding dimension mismatch: expected 4, got 3");
    }
    try {
        store.insert_chunks({EmbeddedChunk(fn_chunk("fn b() {}", 0, 1, "b.rs"), {1.0f, 0.0f})});
        CHECK(!"embedding dimension mismatch must fail");
    } catch (const Error &e) {
        CHECK(std::string(e.what()) == "Embedding dimension mismatch: expected 4, got 2");   // store.rs:666-672
    }
    CHECK(store.delete_chunks({0}) == 1);
    CHECK(!store.is_indexed());                                       // store.rs:605-607
}

// additive methods: filtered / batch / tagged (language + path), deterministic pseudo-random embeddings
static void test_additive_methods()
{
    const size_t d = 64, n = 300;
    VectorStore store = VectorStore::create("", d);
    const char *paths[5] = {"/r/src/auth.rs", "/r/src/math.py", "/r/docs/readme.md", "/r/src/db/store.rs", "/r/Makefile"};
    std::vector<EmbeddedChunk> chunks;
    uint32_t state = 12345u;
    auto rnd = [&]() { state = state * 1664525u + 1013904223u; return ((state >> 8) & 0xFFFF) / 65535.0f - 0.5f; };
    for (size_t i = 0; i < n; ++i) {
        std::vector<float> e(d);
        for (auto &x : e) x = rnd();
        chunks.emplace_back(fn_chunk("chunk " + std::to_string(i), i, i + 1, paths[i % 5]), e);
    }
    store.insert_chunks(chunks);
    store.build_index();
    std::vector<float> q(d);
    for (auto &x : q) x = rnd();
    auto full = store.search(q, n);
    CHECK(full.size() == n);
    for (size_t i = 1; i < full.size(); ++i)                          // ascending (distance, id)
        CHECK(full[i - 1].distance < full[i].distance || (full[i - 1].distance == full[i].distance && full[i - 1].id < full[i].id));
    auto expect_prefix = [&](const std::vector<SearchResult> &got, const std::function<bool(const SearchResult &)> &keep, size_t limit) {
        std::vector<uint32_t> want;
        for (const auto &r : full) if (keep(r) && want.size() < limit) want.push_back(r.id);
        CHECK(got.size() == want.size());
        for (size_t i = 0; i < std::min(got.size(), want.size()); ++i) CHECK(got[i].id == want[i]);
    };
    TagFilter rust; rust.languages = {Language::Rust};
    expect_prefix(store.search_tagged(q, 10, rust), [](const SearchResult &r) { return r.path.size() > 3 && r.path.substr(r.path.size() - 3) == ".rs"; }, 10);
    TagFilter src; src.path_prefix = "src/"; src.project_root = "/r";   // src/search/mod.rs:727-737
    expect_prefix(store.search_tagged(q, 10, src), [](const SearchResult &r) { return r.path.rfind("/r/src/", 0) == 0; }, 10);
    TagFilter db; db.path_contains = "db";                              // src/server/mod.rs:553-559
    expect_prefix(store.search_tagged(q, 10, db), [](const SearchResult &r) { return r.path.find("db") != std::string::npos; }, 10);
    TagFilter go; go.languages = {Language::Go};
    CHECK(store.search_tagged(q, 10, go).empty());
    std::vector<uint32_t> allowed;
    for (uint32_t i = 0; i < n; i += 3) allowed.push_back(i);
    expect_prefix(store.search_filtered(q, 20, RowFilter::from_ids(allowed, n)), [](const SearchResult &r) { return r.id % 3 == 0; }, 20);
    std::vector<std::vector<float>> qs(9, std::vector<float>(d));       // <= 9 query variants, src/search/mod.rs:508-511
    for (auto &v : qs) for (auto &x : v) x = rnd();
    auto batch = store.search_batch(qs, 25);
    CHECK(batch.size() == 9);
    for (size_t j = 0; j < qs.size(); ++j) {
        auto single = store.search(qs[j], 25);
        CHECK(batch[j].size() == single.size());
        for (size_t i = 0; i < std::min(batch[j].size(), single.size()); ++i)
            CHECK(batch[j][i].id == single[i].id && batch[j][i].distance == single[i].distance);
    }
    // query variants -> one list (src/search/mod.rs:508-590 restated on the host: best distance per id, then the best `limit`,
    // ties by id), plain and under a filter: the mirror's one call must equal the dedup of its own per-variant searches
    auto dedup = [&](const std::vector<std::vector<SearchResult>> &lists, size_t limit) {
        std::map<uint32_t, float> best;
        for (const auto &l : lists)
            for (const auto &r : l) {
                auto it = best.find(r.id);
                if (it == best.end() || r.distance < it->second) best[r.id] = r.distance;
            }
        std::vector<std::pair<float, uint32_t>> order;
        for (const auto &kv : best) order.emplace_back(kv.second, kv.first);
        std::sort(order.begin(), order.end());
        if (order.size() > limit) order.resize(limit);
        return order;
    };
    {
        auto got = store.search_variants(qs, 25);
        auto want = dedup(batch, 25);
        CHECK(got.size() == want.size());
        for (size_t i = 0; i < std::min(got.size(), want.size()); ++i) CHECK(got[i].id == want[i].second && got[i].distance == want[i].first);
        std::vector<std::vector<SearchResult>> tagged;
        for (const auto &v : qs) tagged.push_back(store.search_tagged(v, 25, src));
        auto got_t = store.search_variants_tagged(qs, 25, src);
        auto want_t = dedup(tagged, 25);
        CHECK(got_t.size() == want_t.size() && got_t.size() == 25);
        for (size_t i = 0; i < std::min(got_t.size(), want_t.size()); ++i) {
            CHECK(got_t[i].id == want_t[i].second && got_t[i].distance == want_t[i].first);
            CHECK(got_t[i].path.rfind("/r/src/", 0) == 0);
        }
        CHECK(store.search_variants_tagged(qs, 25, go).empty());
    }
    CHECK(language_from_path("main.rs") == Language::Rust);            // src/file/language.rs:145-166
    CHECK(language_from_path("a.pyi") == Language::Python && language_from_path("x.tsx") == Language::TypeScript);
    CHECK(language_from_path("Dockerfile") == Language::Shell && language_from_path(".env") == Language::Shell);
    CHECK(language_from_path("noext") == Language::Unknown && language_from_path("C:\\r\\a.CPP") == Language::Cpp);
}

// opt-in byte prefilter (csgpu_set_byte_prefilter): same results, bit for bit, through the mirror
static void test_byte_prefilter()
{
    const size_t d = 96, n = 6000;
    VectorStore store = VectorStore::create("", d);
    std::vector<EmbeddedChunk> chunks;
    uint32_t state = 777u;
    auto rnd = [&]() { state = state * 1664525u + 1013904223u; return ((state >> 8) & 0xFFFF) / 65535.0f - 0.5f; };
    for (size_t i = 0; i < n; ++i) {
        std::vector<float> e(d);
        for (auto &x : e) x = rnd();
        chunks.emplace_back(fn_chunk("chunk " + std::to_string(i), i, i + 1, "/r/a.rs"), e);
    }
    store.insert_chunks(chunks);
    store.build_index();
    std::vector<float> q(d);
    for (auto &x : q) x = rnd();
    auto plain = store.search(q, 50);
    store.set_byte_prefilter(true);
    auto fast = store.search(q, 50);
    csgpu_stats_t st;
    CHECK(csgpu_stats(store.handle(), &st) == CSGPU_OK);
    CHECK(st.byte_searches == 1 && st.byte_shadow_bytes == n * (128 + 4));
    CHECK(fast.size() == plain.size());
    for (size_t i = 0; i < std::min(fast.size(), plain.size()); ++i)
        CHECK(fast[i].id == plain[i].id && fast[i].distance == plain[i].distance && fast[i].score == plain[i].score);
    store.set_byte_prefilter(false);
    CHECK(csgpu_stats(store.handle(), &st) == CSGPU_OK && st.byte_shadow_bytes == 0);
}

int main(int argc, char **argv)
{
    setenv("CSGPU_I8_MIN_ROWS", "1024", 0);   // read once by the library: lets test_byte_prefilter's small corpus take the int8 route
    if (argc > 1 && std::string(argv[1]) == "--expect-no-gpu") {
        try {
            VectorStore::create("", 4);
            std::printf("FAILED: a store opened without a GPU\n");
            return 1;
        } catch (const Error &e) {
            const bool ok = e.code() == CSGPU_ERR_CUDA && std::string(e.what()).find("no CPU fallback") != std::string::npos;
            std::printf("%s: %s\n", ok ? "NO-GPU OK" : "FAILED: wrong error", e.what());
            return ok ? 0 : 1;
        }
    }
    struct { const char *name; void (*fn)(); } tests[] = {
        {"test_vector_store_creation", test_vector_store_creation}, {"test_insert_and_search", test_insert_and_search},
        {"test_stats", test_stats}, {"test_clear", test_clear}, {"test_get_chunk", test_get_chunk},
        {"test_persistence", test_persistence}, {"test_file_table_reopen", test_file_table_reopen}, {"test_multi_device_store", test_multi_device_store}, {"test_guards", test_guards}, {"test_additive_methods", test_additive_methods},
        {"test_byte_prefilter", test_byte_prefilter},
    };
    for (auto &t : tests) {
        const int before = g_fail;
        try {
            t.fn();
        } catch (const std::exception &e) {
            std::printf("  EXCEPTION in %s: %s\n", t.name, e.what());
            ++g_fail;
        }
        std::printf("%s %s\n", g_fail == before ? "ok  " : "FAIL", t.name);
    }
    std::printf(g_fail ? "%d FAILURE(S)\n" : "ALL PASSED\n", g_fail);
    return g_fail ? 1 : 0;
}
