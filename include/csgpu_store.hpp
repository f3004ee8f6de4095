// csgpu_store.hpp — C++ host-side mirror of the reference's `VectorStore` over the C ABI of csgpu.h.
//
// The reference's host language is Rust (src/vectordb/store.rs); there is no Rust toolchain in this image, so the
// host side above the C ABI is written in C++ (header-only, links against libcsgpu.so) with the reference's method
// names, argument meaning and error behaviour, so that tests/cpp/store_test.cpp reads like the reference's own test
// module (store.rs:826-1029). The Rust binding a maintainer would add is in INTEGRATION.md; codesearch_b200/store.py
// is the same mirror in Python (what the pytest parity tests drive).
//
//   codesearch::VectorStore::create(db_path, dimensions)       VectorStore::new            store.rs:110-176
//   ::open_readonly(db_path, dimensions)                        open_readonly               store.rs:183-250
//   insert_chunks / insert_chunks_with_ids                      store.rs:334, 618-686
//   delete_chunks(ids) -> count                                 store.rs:548-610
//   build_index()                                               store.rs:386-430
//   search(query_embedding, limit) -> vector<SearchResult>      store.rs:431-486   <- the CUDA path (csgpu_search)
//   search_filtered / search_batch / search_tagged              additive
//   search_variants / search_variants_tagged                    additive (§8f N3, N3 x N4): query variants -> one list
//   get_chunk / get_chunk_as_result / get_chunks_by_file / stats / clear / is_indexed
//
// Chunk metadata (the LMDB "chunks" table, store.rs:97) stays on the host: here an ordered map keyed by chunk id,
// persisted as <db>/chunks.bin next to the device snapshot <db>/gpu (csgpu_save / csgpu_load). Only the arroy block
// (store.rs:446-459) runs on the GPU. This header contains no scoring arithmetic and no fallback: every search is
// one call into the C ABI, and every failure surfaces as codesearch::Error carrying the ABI's code and text.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <optional>
#include <set>
#include <stdexcept>
#include <string>
#include <sys/stat.h>
#include <unordered_map>
#include <vector>

#include "csgpu.h"

namespace codesearch {

class Error : public std::runtime_error {
   public:
    Error(int code, const std::string &msg) : std::runtime_error(msg), code_(code) {}
    int code() const { return code_; }

   private:
    int code_;
};

// chunker::Chunk (src/chunker/mod.rs:22-63), the fields the store persists (ChunkMetadata, store.rs:19-39)
struct Chunk {
    std::string content;
    size_t start_line = 0, end_line = 0;
    std::string kind;
    std::string path;
    std::optional<std::string> signature, docstring, context;
    std::string hash;
    std::optional<std::string> context_prev, context_next;
    Chunk() = default;
    Chunk(std::string content_, size_t start, size_t end, std::string kind_, std::string path_)
        : content(std::move(content_)), start_line(start), end_line(end), kind(std::move(kind_)), path(std::move(path_)) {}
};

// embed::EmbeddedChunk (src/embed/batch.rs:47-51)
struct EmbeddedChunk {
    Chunk chunk;
    std::vector<float> embedding;
    EmbeddedChunk(Chunk c, std::vector<float> e) : chunk(std::move(c)), embedding(std::move(e)) {}
};

// store.rs:753-772
struct SearchResult {
    uint32_t id = 0;
    std::string content, path;
    size_t start_line = 0, end_line = 0;
    std::string kind;
    std::optional<std::string> signature, docstring, context;
    std::string hash;
    float distance = 0.f, score = 0.f;
    std::optional<std::string> context_prev, context_next;
};

// store.rs:784-792
struct StoreStats {
    size_t total_chunks = 0, total_files = 0;
    bool indexed = false;
    size_t dimensions = 0;
    uint32_t max_chunk_id = 0;
};

// Allow-set over chunk ids for search_filtered (bit i set <=> chunk id i may be returned)
struct RowFilter {
    std::vector<uint64_t> bitmap;
    uint64_t n_bits = 0;
    static RowFilter from_ids(const std::vector<uint32_t> &ids, uint64_t n_bits)
    {
        RowFilter f;
        f.n_bits = n_bits;
        f.bitmap.assign((n_bits + 63) / 64, 0);
        for (uint32_t id : ids)
            if (id < n_bits) f.bitmap[id >> 6] |= 1ull << (id & 63);
        return f;
    }
};

// ---- row tags: `Language` order of src/file/language.rs:5-29, detection :31-88 -----------------
enum class Language : uint32_t {
    Rust, Python, JavaScript, TypeScript, Go, Java, C, Cpp, CSharp, Ruby, Php, Swift, Kotlin, Shell, Markdown, Json,
    Yaml, Toml, Sql, Html, Css, Xml, Unknown
};

inline std::string normalize_path_str(std::string p)   // src/cache/file_meta.rs:23-25
{
    while (p.rfind("\\\\?\\", 0) == 0) p.erase(0, 4);
    std::replace(p.begin(), p.end(), '\\', '/');
    return p;
}

inline Language language_from_path(const std::string &path)
{
    std::string p = normalize_path_str(path);
    while (!p.empty() && p.back() == '/') p.pop_back();
    const size_t slash = p.rfind('/');
    const std::string name = slash == std::string::npos ? p : p.substr(slash + 1);
    // Path::extension: after the last '.', where a leading dot does not start an extension
    const std::string body = (!name.empty() && name[0] == '.') ? name.substr(1) : name;
    const size_t dot = body.rfind('.');
    std::string ext = dot == std::string::npos ? "" : body.substr(dot + 1);
    std::transform(ext.begin(), ext.end(), ext.begin(), [](unsigned char c) { return (char)std::tolower(c); });
    static const std::unordered_map<std::string, Language> by_ext = {
        {"rs", Language::Rust}, {"py", Language::Python}, {"pyw", Language::Python}, {"pyi", Language::Python},
        {"js", Language::JavaScript}, {"mjs", Language::JavaScript}, {"cjs", Language::JavaScript},
        {"ts", Language::TypeScript}, {"mts", Language::TypeScript}, {"cts", Language::TypeScript},
        {"tsx", Language::TypeScript}, {"jsx", Language::TypeScript}, {"go", Language::Go}, {"java", Language::Java},
        {"c", Language::C}, {"h", Language::C}, {"cpp", Language::Cpp}, {"cc", Language::Cpp}, {"cxx", Language::Cpp},
        {"hpp", Language::Cpp}, {"hxx", Language::Cpp}, {"cs", Language::CSharp}, {"rb", Language::Ruby},
        {"rake", Language::Ruby}, {"php", Language::Php}, {"swift", Language::Swift}, {"kt", Language::Kotlin},
        {"kts", Language::Kotlin}, {"sh", Language::Shell}, {"bash", Language::Shell}, {"zsh", Language::Shell},
        {"md", Language::Markdown}, {"markdown", Language::Markdown}, {"txt", Language::Markdown},
        {"json", Language::Json}, {"yaml", Language::Yaml}, {"yml", Language::Yaml}, {"toml", Language::Toml},
        {"sql", Language::Sql}, {"html", Language::Html}, {"htm", Language::Html}, {"css", Language::Css},
        {"scss", Language::Css}, {"sass", Language::Css}, {"less", Language::Css}, {"xml", Language::Xml},
        {"csproj", Language::Xml}, {"props", Language::Xml}, {"targets", Language::Xml}, {"resx", Language::Xml},
        {"config", Language::Xml}};
    auto it = by_ext.find(ext);
    if (it != by_ext.end()) return it->second;
    static const std::unordered_map<std::string, Language> by_name = {
        {"Dockerfile", Language::Shell}, {"Containerfile", Language::Shell}, {"Makefile", Language::Shell},
        {"GNUmakefile", Language::Shell}, {"makefile", Language::Shell}, {".env", Language::Shell},
        {".envrc", Language::Shell}, {"CMakeLists", Language::Shell}, {"Jenkinsfile", Language::Ruby},
        {"Vagrantfile", Language::Ruby}, {"Fastfile", Language::Ruby}, {"Appfile", Language::Ruby},
        {"Podfile", Language::Ruby}};
    auto jt = by_name.find(name);
    return jt != by_name.end() ? jt->second : Language::Unknown;
}

// What a tagged search may filter on; mirrors the reference's host post-filters (prefix: src/search/mod.rs:727-737,
// substring: src/server/mod.rs:553-559) and adds a language set.
struct TagFilter {
    std::vector<Language> languages;          // empty = any
    std::optional<std::string> path_prefix;   // of the root-relative normalised path
    std::optional<std::string> path_contains;
    std::string project_root;
};

class VectorStore {
   public:
    size_t dimensions = 0;

    // VectorStore::new (store.rs:110-176): creates the directory, hydrates the device index from <db>/gpu when a
    // snapshot exists (next_id = last key + 1, indexed = snapshot present). db_path may be empty (in-memory store).
    static VectorStore create(const std::string &db_path, size_t dimensions) { return VectorStore(db_path, dimensions, false); }
    // The same store over several GPUs of this process (additive; knob CODESEARCH_GPU_DEVICES in the Rust patch): rows are
    // sharded row-wise, search() is one fused launch per device (include/csgpu.h, csgpu_create).
    static VectorStore create(const std::string &db_path, size_t dimensions, const std::vector<int32_t> &devices)
    {
        return VectorStore(db_path, dimensions, false, devices);
    }
    // open_readonly (store.rs:183-250): searches while another process writes; mutations are refused.
    static VectorStore open_readonly(const std::string &db_path, size_t dimensions) { return VectorStore(db_path, dimensions, true); }

    VectorStore(VectorStore &&) = default;
    VectorStore &operator=(VectorStore &&) = default;

    bool is_indexed() const   // store.rs:747
    {
        csgpu_stats_t s;
        check(csgpu_stats(ix_.get(), &s));
        return s.built != 0;
    }

    std::vector<uint32_t> insert_chunks_with_ids(const std::vector<EmbeddedChunk> &chunks)   // store.rs:618-686
    {
        check_writable();
        if (chunks.empty()) return {};
        for (const auto &c : chunks)
            if (c.embedding.size() != dimensions)   // store.rs:666-672
                throw Error(CSGPU_ERR_DIM, "Embedding dimension mismatch: expected " + std::to_string(dimensions) + ", got " +
                                               std::to_string(c.embedding.size()));
        std::vector<float> rows(chunks.size() * dimensions);
        std::vector<uint32_t> ids(chunks.size()), tags(chunks.size());
        for (size_t i = 0; i < chunks.size(); ++i) {
            std::memcpy(rows.data() + i * dimensions, chunks[i].embedding.data(), dimensions * sizeof(float));
            ids[i] = next_id_ + (uint32_t)i;   // dense, monotonically assigned ids (store.rs:659-685)
            tags[i] = tag_of(chunks[i].chunk.path);
        }
        check(csgpu_append_tagged(ix_.get(), rows.data(), ids.data(), tags.data(), chunks.size()));
        for (size_t i = 0; i < chunks.size(); ++i) chunks_[ids[i]] = chunks[i].chunk;
        next_id_ += (uint32_t)chunks.size();
        return ids;
    }
    size_t insert_chunks(const std::vector<EmbeddedChunk> &chunks) { return insert_chunks_with_ids(chunks).size(); }   // store.rs:334

    size_t delete_chunks(const std::vector<uint32_t> &chunk_ids)   // store.rs:548-610
    {
        check_writable();
        if (chunk_ids.empty()) return 0;
        uint64_t removed = 0;
        check(csgpu_remove(ix_.get(), chunk_ids.data(), chunk_ids.size(), &removed));
        for (uint32_t id : chunk_ids) chunks_.erase(id);
        return (size_t)removed;
    }

    void build_index()   // store.rs:386-430
    {
        check_writable();
        check(csgpu_build(ix_.get()));
        if (!db_path_.empty()) save_snapshot();
    }

    void clear()   // store.rs:690-706
    {
        check_writable();
        check(csgpu_clear(ix_.get()));
        chunks_.clear();
        files_.clear();
        file_paths_.clear();
        next_id_ = 0;
        if (!db_path_.empty())
            for (const char *nm : {"gpu/meta.json", "gpu/ids.u32", "gpu/rows.f32", "gpu/rows.bf16", "gpu/tags.u32", "gpu/zero.u32", "chunks.bin", "files.bin"})
                std::remove((db_path_ + "/" + nm).c_str());
    }

    // store.rs:431-486 — guards (:432-444) are enforced by the library with the reference's literal messages;
    // the metadata join (:464-483) and score = 1 - distance (:478) happen here.
    std::vector<SearchResult> search(const std::vector<float> &query_embedding, size_t limit) const
    {
        std::vector<uint32_t> ids(std::max<size_t>(limit, 1));
        std::vector<float> dist(std::max<size_t>(limit, 1));
        uint32_t n = 0;
        check(csgpu_search(ix_.get(), query_embedding.data(), (uint32_t)query_embedding.size(), (uint32_t)limit, ids.data(),
                           dist.data(), &n));
        return join(ids.data(), dist.data(), n);
    }

    std::vector<SearchResult> search_filtered(const std::vector<float> &q, size_t limit, const RowFilter &filter) const
    {
        std::vector<uint32_t> ids(std::max<size_t>(limit, 1));
        std::vector<float> dist(std::max<size_t>(limit, 1));
        uint32_t n = 0;
        check(csgpu_search_filtered(ix_.get(), q.data(), (uint32_t)q.size(), (uint32_t)limit, filter.bitmap.data(), filter.n_bits,
                                    ids.data(), dist.data(), &n));
        return join(ids.data(), dist.data(), n);
    }

    // the <= 9 query variants of src/search/mod.rs:508-511 in one call
    std::vector<std::vector<SearchResult>> search_batch(const std::vector<std::vector<float>> &queries, size_t limit) const
    {
        const size_t b = queries.size(), k = std::max<size_t>(limit, 1);
        std::vector<float> q(b * dimensions);
        for (size_t j = 0; j < b; ++j) {
            if (queries[j].size() != dimensions)
                throw Error(CSGPU_ERR_DIM, "Query embedding dimension mismatch: expected " + std::to_string(dimensions) + ", got " +
                                               std::to_string(queries[j].size()));
            std::memcpy(q.data() + j * dimensions, queries[j].data(), dimensions * sizeof(float));
        }
        std::vector<uint32_t> ids(b * k), n(b, 0);
        std::vector<float> dist(b * k);
        check(csgpu_search_batch(ix_.get(), q.data(), (uint32_t)dimensions, (uint32_t)b, (uint32_t)limit, ids.data(), dist.data(), n.data()));
        std::vector<std::vector<SearchResult>> out(b);
        for (size_t j = 0; j < b; ++j) out[j] = join(ids.data() + j * limit, dist.data() + j * limit, n[j]);
        return out;
    }

    // [b][dimensions] row-major, with the reference's guard message on a wrong length (store.rs:432-438)
    std::vector<float> flatten(const std::vector<std::vector<float>> &queries) const
    {
        std::vector<float> q(queries.size() * dimensions);
        for (size_t j = 0; j < queries.size(); ++j) {
            if (queries[j].size() != dimensions)
                throw Error(CSGPU_ERR_DIM, "Query embedding dimension mismatch: expected " + std::to_string(dimensions) + ", got " +
                                               std::to_string(queries[j].size()));
            std::memcpy(q.data() + j * dimensions, queries[j].data(), dimensions * sizeof(float));
        }
        return q;
    }
    // TagFilter -> csgpu_predicate_t; `bm` receives the per-file bitmap the predicate points into (keep it alive for the call)
    csgpu_predicate_t predicate_of(const TagFilter &f, std::vector<uint64_t> &bm) const
    {
        csgpu_predicate_t p;
        std::memset(&p, 0, sizeof p);
        p.lang_mask = 0xFFFFFFFFu;
        p.file_lo = 0;
        p.file_hi = 0xFFFFFFFFu;
        if (!f.languages.empty()) {
            p.lang_mask = 0;
            for (Language l : f.languages) p.lang_mask |= 1u << (uint32_t)l;
        }
        if (f.path_prefix || f.path_contains) {
            const std::string root = normalize_path_str(f.project_root);
            const std::string prefix = f.path_prefix ? normalize_path_str(*f.path_prefix) : std::string();
            bm.assign(std::max<size_t>((file_paths_.size() + 63) / 64, 1), 0);
            for (size_t fid = 0; fid < file_paths_.size(); ++fid) {
                const std::string &path = file_paths_[fid];
                std::string rel = (!root.empty() && path.rfind(root, 0) == 0) ? path.substr(root.size()) : path;
                while (!rel.empty() && rel[0] == '/') rel.erase(0, 1);
                while (rel.rfind("./", 0) == 0) rel.erase(0, 2);
                bool ok = true;
                if (f.path_prefix) ok = rel.rfind(prefix, 0) == 0;
                if (ok && f.path_contains) ok = path.find(*f.path_contains) != std::string::npos;
                if (ok) bm[fid >> 6] |= 1ull << (fid & 63);
            }
            p.file_bitmap = bm.data();
            p.n_file_bits = file_paths_.size();
        }
        return p;
    }

    // search() restricted by language / path BEFORE scoring (device-side row-tag predicate), so `limit` results survive
    std::vector<SearchResult> search_tagged(const std::vector<float> &q, size_t limit, const TagFilter &f) const
    {
        std::vector<uint64_t> bm;
        const csgpu_predicate_t p = predicate_of(f, bm);
        std::vector<uint32_t> ids(std::max<size_t>(limit, 1));
        std::vector<float> dist(std::max<size_t>(limit, 1));
        uint32_t n = 0;
        check(csgpu_search_tagged(ix_.get(), q.data(), (uint32_t)q.size(), (uint32_t)limit, &p, ids.data(), dist.data(), &n));
        return join(ids.data(), dist.data(), n);
    }

    // The <= 16 query variants of ONE user query -> one list: per chunk id the best distance, then the best `limit`
    // (src/search/mod.rs:508-590: par_iter of searches + HashMap / BinaryHeap dedup) in one call.
    std::vector<SearchResult> search_variants(const std::vector<std::vector<float>> &queries, size_t limit) const
    {
        const std::vector<float> q = flatten(queries);
        std::vector<uint32_t> ids(std::max<size_t>(limit, 1));
        std::vector<float> dist(std::max<size_t>(limit, 1));
        uint32_t n = 0;
        check(csgpu_search_variants(ix_.get(), q.data(), (uint32_t)dimensions, (uint32_t)queries.size(), (uint32_t)limit, ids.data(), dist.data(), &n));
        return join(ids.data(), dist.data(), n);
    }
    // ... under a language / path filter: the hybrid search of src/search/mod.rs:508-590 with the post-filters of :727-737 applied
    // BEFORE scoring, one call
    std::vector<SearchResult> search_variants_tagged(const std::vector<std::vector<float>> &queries, size_t limit, const TagFilter &f) const
    {
        const std::vector<float> q = flatten(queries);
        std::vector<uint64_t> bm;
        const csgpu_predicate_t p = predicate_of(f, bm);
        std::vector<uint32_t> ids(std::max<size_t>(limit, 1));
        std::vector<float> dist(std::max<size_t>(limit, 1));
        uint32_t n = 0;
        check(csgpu_search_variants_tagged(ix_.get(), q.data(), (uint32_t)dimensions, (uint32_t)queries.size(), (uint32_t)limit, &p,
                                           ids.data(), dist.data(), &n));
        return join(ids.data(), dist.data(), n);
    }

    std::optional<Chunk> get_chunk(uint32_t id) const   // store.rs:709
    {
        auto it = chunks_.find(id);
        if (it == chunks_.end()) return std::nullopt;
        return it->second;
    }
    std::optional<SearchResult> get_chunk_as_result(uint32_t id) const   // store.rs:715
    {
        auto it = chunks_.find(id);
        if (it == chunks_.end()) return std::nullopt;
        return to_result(id, it->second, 0.f);
    }
    std::map<std::string, std::vector<uint32_t>> get_chunks_by_file() const   // store.rs:529
    {
        std::map<std::string, std::vector<uint32_t>> out;
        for (const auto &kv : chunks_) out[kv.second.path].push_back(kv.first);
        return out;
    }
    StoreStats stats() const   // store.rs:501
    {
        StoreStats s;
        std::set<std::string> files;
        for (const auto &kv : chunks_) files.insert(kv.second.path);
        s.total_chunks = chunks_.size();
        s.total_files = files.size();
        s.indexed = is_indexed();
        s.dimensions = dimensions;
        s.max_chunk_id = chunks_.empty() ? 0 : chunks_.rbegin()->first;
        return s;
    }
    // opt-in: batches at tensor-core speed with exact fp32 results (csgpu_set_tensor_prefilter)
    void set_tensor_prefilter(bool on) { check(csgpu_set_tensor_prefilter(ix_.get(), on ? 1 : 0)); }
    // opt-in: single queries from a 1-byte shadow + exact fp32 rescoring, bit-identical results (csgpu_set_byte_prefilter)
    void set_byte_prefilter(bool on) { check(csgpu_set_byte_prefilter(ix_.get(), on ? 1 : 0)); }
    csgpu_index *handle() const { return ix_.get(); }

   private:
    struct Destroy {
        void operator()(csgpu_index *p) const { csgpu_destroy(p); }
    };
    std::unique_ptr<csgpu_index, Destroy> ix_;
    std::string db_path_;
    bool read_only_ = false;
    uint32_t next_id_ = 0;
    std::map<uint32_t, Chunk> chunks_;
    std::unordered_map<std::string, uint32_t> files_;   // normalised path -> file_id (first-seen order)
    std::vector<std::string> file_paths_;

    static void check(int rc)
    {
        if (rc != CSGPU_OK) throw Error(rc, csgpu_last_error());
    }
    void check_writable() const
    {
        if (read_only_) throw Error(CSGPU_ERR_ARG, "VectorStore is open read-only");
    }
    uint32_t tag_of(const std::string &path)
    {
        const std::string key = normalize_path_str(path);
        auto it = files_.find(key);
        uint32_t fid;
        if (it == files_.end()) {
            fid = (uint32_t)file_paths_.size();
            if (fid > CSGPU_TAG_FILE_MASK) throw Error(CSGPU_ERR_ARG, "more than 2^27 files: file_id does not fit the tag");
            files_.emplace(key, fid);
            file_paths_.push_back(key);
        } else {
            fid = it->second;
        }
        return CSGPU_TAG((uint32_t)language_from_path(key), fid);
    }
    static SearchResult to_result(uint32_t id, const Chunk &c, float distance)
    {
        SearchResult r;
        r.id = id; r.content = c.content; r.path = c.path; r.start_line = c.start_line; r.end_line = c.end_line;
        r.kind = c.kind; r.signature = c.signature; r.docstring = c.docstring; r.context = c.context; r.hash = c.hash;
        r.distance = distance;
        r.score = 1.0f - distance;   // store.rs:478
        r.context_prev = c.context_prev; r.context_next = c.context_next;
        return r;
    }
    std::vector<SearchResult> join(const uint32_t *ids, const float *dist, uint32_t n) const
    {
        std::vector<SearchResult> out;
        out.reserve(n);
        for (uint32_t i = 0; i < n; ++i) {
            auto it = chunks_.find(ids[i]);
            if (it == chunks_.end()) continue;   // store.rs:465 silently drops ids without metadata
            out.push_back(to_result(ids[i], it->second, dist[i]));
        }
        return out;
    }

    // ---- persistence: <db>/gpu (csgpu_save) + <db>/chunks.bin (length-prefixed fields) ----
    static void put_str(FILE *f, const std::string &s)
    {
        const uint64_t n = s.size();
        fwrite(&n, sizeof n, 1, f);
        fwrite(s.data(), 1, s.size(), f);
    }
    static void put_opt(FILE *f, const std::optional<std::string> &s)
    {
        const uint8_t has = s.has_value();
        fwrite(&has, 1, 1, f);
        if (has) put_str(f, *s);
    }
    static bool get_str(FILE *f, std::string *s)
    {
        uint64_t n = 0;
        if (fread(&n, sizeof n, 1, f) != 1 || n > (1ull << 32)) return false;
        s->resize(n);
        return n == 0 || fread(&(*s)[0], 1, n, f) == n;
    }
    static bool get_opt(FILE *f, std::optional<std::string> *s)
    {
        uint8_t has = 0;
        if (fread(&has, 1, 1, f) != 1) return false;
        if (!has) { s->reset(); return true; }
        std::string v;
        if (!get_str(f, &v)) return false;
        *s = std::move(v);
        return true;
    }
    void save_snapshot() const
    {
        const std::string tmp = db_path_ + "/chunks.bin.tmp";
        FILE *f = fopen(tmp.c_str(), "wb");
        if (!f) throw Error(CSGPU_ERR_ARG, "cannot write " + tmp);
        const uint64_t n = chunks_.size();
        fwrite(&n, sizeof n, 1, f);
        for (const auto &kv : chunks_) {
            const Chunk &c = kv.second;
            const uint64_t id = kv.first, a = c.start_line, b = c.end_line;
            fwrite(&id, sizeof id, 1, f); fwrite(&a, sizeof a, 1, f); fwrite(&b, sizeof b, 1, f);
            put_str(f, c.content); put_str(f, c.kind); put_str(f, c.path); put_str(f, c.hash);
            put_opt(f, c.signature); put_opt(f, c.docstring); put_opt(f, c.context); put_opt(f, c.context_prev); put_opt(f, c.context_next);
        }
        fclose(f);
        if (std::rename(tmp.c_str(), (db_path_ + "/chunks.bin").c_str()) != 0) throw Error(CSGPU_ERR_ARG, "cannot publish chunks.bin");
        // the file table, verbatim in file-id order: the row tags (HBM, gpu/tags.u32) index into it, and after delete +
        // re-insert cycles it cannot be re-derived from the surviving chunks
        const std::string ftmp = db_path_ + "/files.bin.tmp";
        FILE *ff = fopen(ftmp.c_str(), "wb");
        if (!ff) throw Error(CSGPU_ERR_ARG, "cannot write " + ftmp);
        const uint64_t nf = file_paths_.size();
        fwrite(&nf, sizeof nf, 1, ff);
        for (const std::string &p : file_paths_) put_str(ff, p);
        fclose(ff);
        if (std::rename(ftmp.c_str(), (db_path_ + "/files.bin").c_str()) != 0) throw Error(CSGPU_ERR_ARG, "cannot publish files.bin");
        check(csgpu_save(ix_.get(), (db_path_ + "/gpu").c_str()));
    }
    void load_chunks()
    {
        FILE *f = fopen((db_path_ + "/chunks.bin").c_str(), "rb");
        if (!f) return;
        uint64_t n = 0;
        bool ok = fread(&n, sizeof n, 1, f) == 1;
        for (uint64_t i = 0; ok && i < n; ++i) {
            uint64_t id = 0, a = 0, b = 0;
            Chunk c;
            ok = fread(&id, sizeof id, 1, f) == 1 && fread(&a, sizeof a, 1, f) == 1 && fread(&b, sizeof b, 1, f) == 1 &&
                 get_str(f, &c.content) && get_str(f, &c.kind) && get_str(f, &c.path) && get_str(f, &c.hash) &&
                 get_opt(f, &c.signature) && get_opt(f, &c.docstring) && get_opt(f, &c.context) && get_opt(f, &c.context_prev) &&
                 get_opt(f, &c.context_next);
            if (!ok) break;
            c.start_line = a; c.end_line = b;
            chunks_[(uint32_t)id] = std::move(c);
        }
        fclose(f);
        if (!ok) throw Error(CSGPU_ERR_ARG, "chunks.bin is truncated or corrupt");
        if (FILE *ff = fopen((db_path_ + "/files.bin").c_str(), "rb")) {
            uint64_t nf = 0;
            bool fok = fread(&nf, sizeof nf, 1, ff) == 1 && nf <= (uint64_t)CSGPU_TAG_FILE_MASK + 1;
            for (uint64_t i = 0; fok && i < nf; ++i) {
                std::string p;
                fok = get_str(ff, &p);
                if (fok) { files_.emplace(p, (uint32_t)file_paths_.size()); file_paths_.push_back(std::move(p)); }
            }
            fclose(ff);
            if (!fok) throw Error(CSGPU_ERR_ARG, "files.bin is truncated or corrupt");
        } else {
            for (const auto &kv : chunks_) tag_of(kv.second.path);   // snapshot older than files.bin: first-seen order over surviving chunks
        }
        if (!chunks_.empty()) next_id_ = chunks_.rbegin()->first + 1;   // store.rs:141-144
    }

    VectorStore(const std::string &db_path, size_t dims, bool read_only, const std::vector<int32_t> &devices = {})
        : dimensions(dims), db_path_(db_path), read_only_(read_only)
    {
        csgpu_index *raw = nullptr;
        check(csgpu_create(&raw, (uint32_t)dims, CSGPU_DTYPE_F32, devices.empty() ? nullptr : devices.data(),
                           devices.empty() ? 1u : (uint32_t)devices.size()));
        ix_.reset(raw);
        if (!db_path_.empty()) {
            mkdir(db_path_.c_str(), 0755);   // store.rs:116 create_dir_all
            struct stat st;
            if (stat((db_path_ + "/gpu/meta.json").c_str(), &st) == 0) {
                check(csgpu_load(ix_.get(), (db_path_ + "/gpu").c_str()));
                load_chunks();
            }
        }
    }
};

}  // namespace codesearch
