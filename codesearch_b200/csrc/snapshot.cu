// snapshot.cu — sidecar snapshot of the device index.
//
// HBM is volatile; the reference's truth lives in LMDB (`.codesearch.db/data.mdb`, written by arroy through
// /root/reference/src/vectordb/store.rs:618-686 and committed at build_index :422-430). Re-hydrating the GPU from
// LMDB would go item by item through arroy, so `build_index` also writes a flat sidecar next to it,
//     <db>/gpu/meta.json   {format, version, dim, dim_pad, dtype, rows, zero_ids, tags, checksum}
//     <db>/gpu/ids.u32     [rows]            chunk ids in row order
//     <db>/gpu/rows.f32    [rows, dim_pad]   unit-normalised rows exactly as they sit in HBM   (rows.bf16: [rows, dim])
//     <db>/gpu/tags.u32    [rows]            packed row tags (lang_id << 27 | file_id)
//     <db>/gpu/zero.u32    [zero_ids] ids of zero-norm rows (distance 0.0, kept off the matrix), then [zero_ids] tags
// and `VectorStore::new / open_readonly` (store.rs:110-176,183-250) load it with large sequential reads +
// cudaMemcpyAsync from pinned staging. A loaded index is already built and returns bit-identical results
// (the rows are not re-normalised). db_discovery's validity rule (src/db_discovery/mod.rs:8-15) extends naturally:
// a snapshot is valid iff meta.json parses, sizes match and the checksum agrees.
#include <cerrno>
#include <cinttypes>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <sys/stat.h>

#include "index.h"

namespace csgpu {

constexpr size_t SNAP_CHUNK = 64u << 20;   // staging buffer size (pinned, two of them)

static uint64_t mix_words(uint64_t h, const void *data, size_t bytes)
{
    const uint64_t *w = static_cast<const uint64_t *>(data);
    const size_t nw = bytes / 8;
    for (size_t i = 0; i < nw; ++i) h = (h ^ w[i]) * 0x100000001B3ull;
    const unsigned char *t = static_cast<const unsigned char *>(data) + nw * 8;
    for (size_t i = 0; i < bytes % 8; ++i) h = (h ^ t[i]) * 0x100000001B3ull;
    return h;
}

// Version-2 checksum: one running hash per FILE over its whole byte stream (chunk- and shard-boundary independent: a
// partial word is carried into the next update), combined at the end. How the rows were split over devices when the
// snapshot was written — uneven shards after small appends / deletes, or a different n_devices at load — therefore does
// not change it (round-1 advisor finding: the version-1 sum folded ids/rows/tags shard by shard).
struct StreamHash {
    uint64_t h = 0xCBF29CE484222325ull;
    unsigned char carry[8];
    size_t nc = 0;
    void word(uint64_t w) { h = (h ^ w) * 0x100000001B3ull; }
    void update(const void *data, size_t bytes)
    {
        const unsigned char *p = static_cast<const unsigned char *>(data);
        while (nc && nc < 8 && bytes) { carry[nc++] = *p++; --bytes; }
        if (nc == 8) { uint64_t w; memcpy(&w, carry, 8); word(w); nc = 0; }
        for (; bytes >= 8; bytes -= 8, p += 8) { uint64_t w; memcpy(&w, p, 8); word(w); }
        while (bytes) { carry[nc++] = *p++; --bytes; }
    }
    uint64_t final() const
    {
        uint64_t r = h;
        for (size_t i = 0; i < nc; ++i) r = (r ^ carry[i]) * 0x100000001B3ull;
        return r;
    }
};
static uint64_t combine_hashes(const StreamHash &ids, const StreamHash &rows, const StreamHash &tags, const StreamHash &zero)
{
    uint64_t r = 0xCBF29CE484222325ull;
    for (uint64_t v : {ids.final(), rows.final(), tags.final(), zero.final()}) r = (r ^ v) * 0x100000001B3ull;
    return r;
}

struct Stager {
    void *buf[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    ~Stager()
    {
        for (int i = 0; i < 2; ++i) { if (buf[i]) cudaFreeHost(buf[i]); if (ev[i]) cudaEventDestroy(ev[i]); }
    }
    int init()
    {
        for (int i = 0; i < 2; ++i) {
            CS_CUDA(cudaHostAlloc(&buf[i], SNAP_CHUNK, cudaHostAllocDefault));
            CS_CUDA(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
        }
        return CSGPU_OK;
    }
};

// device [bytes] -> file, double-buffered; folds the bytes into *sum
static int dump_device(FILE *f, const void *dev, size_t bytes, cudaStream_t st, Stager &sg, StreamHash *sum)
{
    const char *src = static_cast<const char *>(dev);
    size_t off = 0;
    int cur = 0;
    size_t pend_bytes[2] = {0, 0};
    if (bytes) {
        pend_bytes[0] = std::min(SNAP_CHUNK, bytes);
        CS_CUDA(cudaMemcpyAsync(sg.buf[0], src, pend_bytes[0], cudaMemcpyDeviceToHost, st));
        CS_CUDA(cudaEventRecord(sg.ev[0], st));
    }
    while (off < bytes) {
        const size_t nb = pend_bytes[cur];
        const size_t next_off = off + nb;
        if (next_off < bytes) {   // start the next copy before touching this one
            pend_bytes[cur ^ 1] = std::min(SNAP_CHUNK, bytes - next_off);
            CS_CUDA(cudaMemcpyAsync(sg.buf[cur ^ 1], src + next_off, pend_bytes[cur ^ 1], cudaMemcpyDeviceToHost, st));
            CS_CUDA(cudaEventRecord(sg.ev[cur ^ 1], st));
        }
        CS_CUDA(cudaEventSynchronize(sg.ev[cur]));
        sum->update(sg.buf[cur], nb);
        if (fwrite(sg.buf[cur], 1, nb, f) != nb) return fail(CSGPU_ERR_ARG, std::string("snapshot write failed: ") + strerror(errno));
        off = next_off;
        cur ^= 1;
    }
    return CSGPU_OK;
}

// file -> device [bytes]
static int fill_device(FILE *f, void *dev, size_t bytes, cudaStream_t st, Stager &sg, StreamHash *sum, uint64_t *legacy)
{
    char *dst = static_cast<char *>(dev);
    size_t off = 0;
    int cur = 0;
    bool busy[2] = {false, false};
    while (off < bytes) {
        const size_t nb = std::min(SNAP_CHUNK, bytes - off);
        if (busy[cur]) { CS_CUDA(cudaEventSynchronize(sg.ev[cur])); busy[cur] = false; }
        if (fread(sg.buf[cur], 1, nb, f) != nb) return fail(CSGPU_ERR_ARG, "snapshot is truncated");
        sum->update(sg.buf[cur], nb);
        if (legacy) *legacy = mix_words(*legacy, sg.buf[cur], nb);
        CS_CUDA(cudaMemcpyAsync(dst + off, sg.buf[cur], nb, cudaMemcpyHostToDevice, st));
        CS_CUDA(cudaEventRecord(sg.ev[cur], st));
        busy[cur] = true;
        off += nb;
        cur ^= 1;
    }
    CS_CUDA(cudaStreamSynchronize(st));
    return CSGPU_OK;
}

static FILE *open_in(const std::string &dir, const char *name, const char *mode)
{
    return fopen((dir + "/" + name).c_str(), mode);
}

struct FileCloser {
    FILE *f;
    ~FileCloser() { if (f) fclose(f); }
};

int snapshot_save(const csgpu_index *ix, const char *dir_c)
{
    if (!ix || !dir_c) return fail(CSGPU_ERR_ARG, "null argument");
    if (!ix->built) return fail(CSGPU_ERR_NOT_BUILT, "Index not built. Call build_index() after inserting chunks.");
    const std::string dir(dir_c);
    if (mkdir(dir.c_str(), 0755) != 0 && errno != EEXIST) return fail(CSGPU_ERR_ARG, "cannot create " + dir + ": " + strerror(errno));
    const bool bf16 = ix->dtype == CSGPU_DTYPE_BF16;
    const size_t row_bytes = bf16 ? (size_t)ix->dim * 2 : (size_t)ix->dim_pad * sizeof(float);
    Stager sg;
    {
        DeviceGuard dg(ix->shards[0]->device);
        int rc = sg.init();
        if (rc) return rc;
    }
    uint64_t rows = 0;
    StreamHash h_ids, h_rows, h_tags, h_zero;
    {
        FileCloser fi{open_in(dir, "ids.u32.tmp", "wb")}, fr{open_in(dir, bf16 ? "rows.bf16.tmp" : "rows.f32.tmp", "wb")};
        FileCloser ft{open_in(dir, "tags.u32.tmp", "wb")};
        if (!fi.f || !fr.f || !ft.f) return fail(CSGPU_ERR_ARG, "cannot open snapshot files in " + dir + ": " + strerror(errno));
        for (const Shard *sh : ix->shards) {   // shards are concatenated in order: row order is (shard, local row)
            DeviceGuard dg(sh->device);
            int rc = dump_device(fi.f, sh->ids, sh->n_built * sizeof(uint32_t), sh->stream, sg, &h_ids);
            if (!rc) rc = dump_device(fr.f, bf16 ? sh->rows_bf16 : (const void *)sh->rows, sh->n_built * row_bytes, sh->stream, sg, &h_rows);
            if (!rc) rc = dump_device(ft.f, sh->tags, sh->n_built * sizeof(uint32_t), sh->stream, sg, &h_tags);
            if (rc) return rc;
            rows += sh->n_built;
        }
    }
    {
        FileCloser fz{open_in(dir, "zero.u32.tmp", "wb")};
        if (!fz.f) return fail(CSGPU_ERR_ARG, "cannot open zero.u32 in " + dir);
        if (!ix->zero_ids.empty()) {
            h_zero.update(ix->zero_ids.data(), ix->zero_ids.size() * sizeof(uint32_t));
            if (fwrite(ix->zero_ids.data(), sizeof(uint32_t), ix->zero_ids.size(), fz.f) != ix->zero_ids.size())
                return fail(CSGPU_ERR_ARG, "snapshot write failed");
            std::vector<uint32_t> zt(ix->zero_tags);
            zt.resize(ix->zero_ids.size(), CSGPU_TAG_NONE);
            h_zero.update(zt.data(), zt.size() * sizeof(uint32_t));
            if (fwrite(zt.data(), sizeof(uint32_t), zt.size(), fz.f) != zt.size()) return fail(CSGPU_ERR_ARG, "snapshot write failed");
        }
    }
    {
        FileCloser fm{open_in(dir, "meta.json.tmp", "wb")};
        if (!fm.f) return fail(CSGPU_ERR_ARG, "cannot open meta.json in " + dir);
        fprintf(fm.f,
                "{\"format\": \"csgpu-snapshot\", \"version\": 2, \"dim\": %u, \"dim_pad\": %u, \"dtype\": \"%s\", "
                "\"rows\": %" PRIu64 ", \"zero_ids\": %zu, \"tags\": 1, \"checksum\": \"%016" PRIx64 "\"}\n",
                ix->dim, ix->dim_pad, bf16 ? "bf16" : "f32", rows, ix->zero_ids.size(), combine_hashes(h_ids, h_rows, h_tags, h_zero));
    }
    // publish: data files first, meta.json last (a reader that finds meta.json finds complete data)
    const char *names[5] = {"ids.u32", bf16 ? "rows.bf16" : "rows.f32", "tags.u32", "zero.u32", "meta.json"};
    for (const char *nm : names)
        if (rename((dir + "/" + nm + ".tmp").c_str(), (dir + "/" + nm).c_str()) != 0)
            return fail(CSGPU_ERR_ARG, std::string("rename failed for ") + nm + ": " + strerror(errno));
    return CSGPU_OK;
}

static bool json_u64(const std::string &js, const char *key, uint64_t *out)
{
    const std::string pat = std::string("\"") + key + "\":";
    size_t p = js.find(pat);
    if (p == std::string::npos) return false;
    p += pat.size();
    while (p < js.size() && js[p] == ' ') ++p;
    char *end = nullptr;
    *out = strtoull(js.c_str() + p, &end, 10);
    return end != js.c_str() + p;
}
static bool json_str(const std::string &js, const char *key, std::string *out)
{
    const std::string pat = std::string("\"") + key + "\":";
    size_t p = js.find(pat);
    if (p == std::string::npos) return false;
    p = js.find('"', p + pat.size());
    if (p == std::string::npos) return false;
    const size_t q = js.find('"', p + 1);
    if (q == std::string::npos) return false;
    *out = js.substr(p + 1, q - p - 1);
    return true;
}

int snapshot_load(csgpu_index *ix, const char *dir_c, int (*reserve)(csgpu_index *, uint64_t), int (*finish)(csgpu_index *))
{
    if (!ix || !dir_c) return fail(CSGPU_ERR_ARG, "null argument");
    for (const Shard *sh : ix->shards)
        if (sh->n_total) return fail(CSGPU_ERR_ARG, "csgpu_load needs an empty index (csgpu_clear first)");
    if (!ix->zero_ids.empty()) return fail(CSGPU_ERR_ARG, "csgpu_load needs an empty index (csgpu_clear first)");
    const std::string dir(dir_c);
    std::string js;
    {
        FileCloser fm{open_in(dir, "meta.json", "rb")};
        if (!fm.f) return fail(CSGPU_ERR_ARG, "no snapshot at " + dir + " (meta.json missing)");
        char buf[1024];
        size_t n;
        while ((n = fread(buf, 1, sizeof buf, fm.f)) > 0) js.append(buf, n);
    }
    uint64_t version = 0, dim = 0, dim_pad = 0, rows = 0, nzero = 0, has_tags = 0;
    std::string fmt, dtype, checksum;
    if (!json_str(js, "format", &fmt) || fmt != "csgpu-snapshot" || !json_u64(js, "version", &version) || (version != 1 && version != 2) ||
        !json_u64(js, "dim", &dim) || !json_u64(js, "dim_pad", &dim_pad) || !json_u64(js, "rows", &rows) ||
        !json_u64(js, "zero_ids", &nzero) || !json_str(js, "dtype", &dtype) || !json_str(js, "checksum", &checksum))
        return fail(CSGPU_ERR_ARG, "snapshot meta.json is malformed");
    json_u64(js, "tags", &has_tags);   // absent in snapshots written before the tag column existed: rows load untagged
    const bool bf16 = ix->dtype == CSGPU_DTYPE_BF16;
    if (dim != ix->dim || dim_pad != ix->dim_pad) {
        char b[160];
        snprintf(b, sizeof b, "Snapshot dimension mismatch: expected %u, got %" PRIu64, ix->dim, dim);
        return fail(CSGPU_ERR_DIM, b);
    }
    if (dtype != (bf16 ? "bf16" : "f32")) return fail(CSGPU_ERR_ARG, "snapshot dtype does not match the index dtype");
    const size_t row_bytes = bf16 ? (size_t)ix->dim * 2 : (size_t)ix->dim_pad * sizeof(float);
    int rc = reserve(ix, rows);
    if (rc) return rc;
    Stager sg;
    {
        DeviceGuard dg(ix->shards[0]->device);
        if ((rc = sg.init())) return rc;
    }
    // version 1 folded one sum over (ids, rows, tags) per shard of an EVEN split at the writer's device count; it is only
    // reproducible when this index splits the same way (single device always does). Version 2 hashes per file.
    uint64_t sum = 0xCBF29CE484222325ull;
    uint64_t *legacy = version == 1 ? &sum : nullptr;
    StreamHash h_ids, h_rows, h_tags, h_zero;
    {
        FileCloser fi{open_in(dir, "ids.u32", "rb")}, fr{open_in(dir, bf16 ? "rows.bf16" : "rows.f32", "rb")};
        FileCloser ft{has_tags ? open_in(dir, "tags.u32", "rb") : nullptr};
        if (!fi.f || !fr.f || (has_tags && !ft.f)) return fail(CSGPU_ERR_ARG, "snapshot data files are missing in " + dir);
        const uint64_t G = ix->shards.size();
        for (uint64_t g = 0; g < G; ++g) {
            Shard *sh = ix->shards[g];
            const uint64_t a = rows * g / G, b = rows * (g + 1) / G, m = b - a;
            DeviceGuard dg(sh->device);
            rc = fill_device(fi.f, sh->ids, m * sizeof(uint32_t), sh->stream, sg, &h_ids, legacy);
            if (!rc) rc = fill_device(fr.f, bf16 ? sh->rows_bf16 : (void *)sh->rows, m * row_bytes, sh->stream, sg, &h_rows, legacy);
            if (!rc && has_tags) rc = fill_device(ft.f, sh->tags, m * sizeof(uint32_t), sh->stream, sg, &h_tags, legacy);
            if (rc) return rc;
            if (!has_tags && sh->cap) CS_CUDA(cudaMemsetAsync(sh->tags, 0xFF, sh->cap * sizeof(uint32_t), sh->stream));
            CS_CUDA(cudaMemsetAsync(sh->status, 0, sh->cap, sh->stream));
            CS_CUDA(cudaStreamSynchronize(sh->stream));
            sh->n_total = sh->n_built = m;
        }
    }
    if (nzero) {
        FileCloser fz{open_in(dir, "zero.u32", "rb")};
        ix->zero_ids.resize(nzero);
        if (!fz.f || fread(ix->zero_ids.data(), sizeof(uint32_t), nzero, fz.f) != nzero) {
            ix->zero_ids.clear();
            return fail(CSGPU_ERR_ARG, "snapshot zero.u32 is missing or truncated");
        }
        sum = mix_words(sum, ix->zero_ids.data(), nzero * sizeof(uint32_t));
        h_zero.update(ix->zero_ids.data(), nzero * sizeof(uint32_t));
        ix->zero_tags.assign(nzero, CSGPU_TAG_NONE);
        if (has_tags) {
            if (fread(ix->zero_tags.data(), sizeof(uint32_t), nzero, fz.f) != nzero) {
                ix->zero_ids.clear(); ix->zero_tags.clear();
                return fail(CSGPU_ERR_ARG, "snapshot zero.u32 is missing or truncated");
            }
            sum = mix_words(sum, ix->zero_tags.data(), nzero * sizeof(uint32_t));
            h_zero.update(ix->zero_tags.data(), nzero * sizeof(uint32_t));
        }
    }
    char hex[32];
    snprintf(hex, sizeof hex, "%016" PRIx64, version == 1 ? sum : combine_hashes(h_ids, h_rows, h_tags, h_zero));
    if (checksum != hex) {
        for (Shard *sh : ix->shards) sh->n_total = sh->n_built = 0;
        ix->zero_ids.clear(); ix->zero_tags.clear();
        return fail(CSGPU_ERR_ARG, "snapshot checksum mismatch (corrupt or partially written snapshot)");
    }
    return finish(ix);
}

}  // namespace csgpu
