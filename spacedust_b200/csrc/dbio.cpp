// MMseqs2 flat-file DB access (see dbio.h).
#include "dbio.h"

#include <dlfcn.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <sys/stat.h>

namespace sddb {

bool fileExists(const std::string &path) {
    struct stat st;
    return stat(path.c_str(), &st) == 0 && S_ISREG(st.st_mode);
}

bool readFile(const std::string &path, std::string &out) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    size_t base = out.size();
    out.resize(base + (size_t)n);
    size_t got = n ? fread(&out[base], 1, (size_t)n, f) : 0;
    fclose(f);
    return got == (size_t)n;
}

static inline uint64_t parseU64(const char *&p) {
    uint64_t v = 0;
    while (*p >= '0' && *p <= '9') v = v * 10 + (uint64_t)(*p++ - '0');
    return v;
}

// zstd is resolved at run time (libzstd.so.1; no header in this image): only compressed DBs need it
namespace {
struct ZstdApi {
    void *h = nullptr;
    size_t (*decompress)(void *, size_t, const void *, size_t) = nullptr;
    unsigned (*isError)(size_t) = nullptr;
    const char *(*errorName)(size_t) = nullptr;
    size_t (*compress)(void *, size_t, const void *, size_t, int) = nullptr;
    size_t (*compressBound)(size_t) = nullptr;
};
bool g_compressOutput = false;
ZstdApi &zstd() {
    static ZstdApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        for (const char *name : {"libzstd.so.1", "libzstd.so"}) {
            api.h = dlopen(name, RTLD_NOW);
            if (api.h) break;
        }
        if (api.h) {
            api.decompress = reinterpret_cast<decltype(api.decompress)>(dlsym(api.h, "ZSTD_decompress"));
            api.isError = reinterpret_cast<decltype(api.isError)>(dlsym(api.h, "ZSTD_isError"));
            api.errorName = reinterpret_cast<decltype(api.errorName)>(dlsym(api.h, "ZSTD_getErrorName"));
            api.compress = reinterpret_cast<decltype(api.compress)>(dlsym(api.h, "ZSTD_compress"));
            api.compressBound = reinterpret_cast<decltype(api.compressBound)>(dlsym(api.h, "ZSTD_compressBound"));
        }
    }
    return api;
}
}  // namespace

// A DB written with --compressed 1 (DBWriter::writeEnd, DBWriter.cpp:326-395): every entry is [u32 n][n bytes][marker]; marker 0 = the
// n bytes are one zstd frame, 0xFF = stored as they are; the index keeps the ORIGINAL length (+ terminator).  The entries are
// expanded once into the layout of a plain DB (what DBReader::getDataCompressed does per access, DBReader.cpp:560-587).
bool Db::inflate(const std::string &path, std::string &err) {
    ZstdApi &z = zstd();
    std::string plain;
    uint64_t total = 0;
    for (const Entry &e : index) total += e.len;
    plain.reserve(total);
    // in ascending offset, so that the expanded entries keep their relative order (DBReader ids of LINEAR_ACCCESS follow the offsets)
    std::vector<size_t> order(index.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return index[a].off < index[b].off; });
    for (size_t oi : order) {
        Entry &e = index[oi];
        if (e.off + 5 > data.size()) { err = "index entry beyond the data file in " + path; return false; }
        uint32_t n;
        memcpy(&n, data.data() + e.off, 4);
        const char *payload = data.data() + e.off + 4;
        if (e.off + 4 + (uint64_t)n + 1 > data.size()) { err = "compressed entry beyond the data file in " + path; return false; }
        const bool isFrame = payload[n] == 0;
        const uint64_t newOff = plain.size();
        if (isFrame) {
            if (!z.decompress || !z.isError) { err = path + " is a compressed DB (--compressed 1) and libzstd.so.1 could not be loaded"; return false; }
            const size_t want = e.len ? (size_t)e.len - 1 : 0;
            plain.resize(newOff + want + 1);
            const size_t got = z.decompress(&plain[newOff], want, payload, n);
            if (z.isError(got) || got != want) {
                err = path + ": zstd entry " + std::to_string(e.key) + " does not decompress to its indexed length (" + (z.isError(got) && z.errorName ? z.errorName(got) : "size mismatch") + ")";
                return false;
            }
            plain[newOff + want] = '\0';
        } else {
            plain.append(payload, n);
            plain.push_back('\0');
            e.len = (uint64_t)n + 1;
        }
        e.off = newOff;
    }
    data.swap(plain);
    return true;
}

bool Db::open(const std::string &path, std::string &err) {
    data.clear();
    index.clear();
    if (fileExists(path)) {
        if (!readFile(path, data)) { err = "cannot read " + path; return false; }
    } else {
        int i = 0;
        while (fileExists(path + "." + std::to_string(i))) {
            if (!readFile(path + "." + std::to_string(i), data)) { err = "cannot read " + path + "." + std::to_string(i); return false; }
            i++;
        }
        if (i == 0 && !fileExists(path + ".index")) { err = "database " + path + " does not exist"; return false; }
    }
    std::string t;
    dbtype = -1;
    bool compressed = false;
    if (readFile(path + ".dbtype", t) && t.size() >= 4) {
        uint32_t v;
        memcpy(&v, t.data(), 4);
        dbtype = (int)(v & 0xffff);
        compressed = (v & 0x80000000u) != 0;      // bit 31: entries are zstd frames (Parameters.h:90-94, DBWriter.cpp:326-395)
    }
    std::string idx;
    if (!readFile(path + ".index", idx)) { err = "cannot read " + path + ".index"; return false; }
    const char *p = idx.c_str();
    while (*p) {
        Entry e;
        e.key = (uint32_t)parseU64(p);
        if (*p != '\t') { err = "malformed index " + path + ".index"; return false; }
        p++;
        e.off = parseU64(p);
        if (*p != '\t') { err = "malformed index " + path + ".index"; return false; }
        p++;
        e.len = parseU64(p);
        while (*p && *p != '\n') p++;
        if (*p == '\n') p++;
        if (!compressed && e.off + e.len > data.size()) { err = "index entry beyond the data file in " + path; return false; }
        index.push_back(e);
    }
    if (compressed && !inflate(path, err)) return false;
    return true;
}

static void splitUnderscore(const std::string &s, std::vector<std::string> &out) {
    out.clear();
    size_t b = 0;
    while (b <= s.size()) {
        size_t e = s.find('_', b);
        if (e == std::string::npos) e = s.size();
        if (e > b) out.push_back(s.substr(b, e - b));     // Util::split drops empty fields
        b = e + 1;
    }
}

static uint64_t fastAtoi(const std::string &s) {
    const char *p = s.c_str();
    return parseU64(p);
}

bool SetDb::open(const std::string &p, std::string &err) {
    path = p;
    Db db;
    if (!db.open(p, err)) return false;
    if (db.dbtype != DBTYPE_AMINO_ACIDS) { err = p + ": only amino-acid set DBs are supported (dbtype " + std::to_string(db.dbtype) + ")"; return false; }
    std::vector<Entry> ord = db.index;
    std::stable_sort(ord.begin(), ord.end(), [](const Entry &a, const Entry &b) { return a.off < b.off; });
    const size_t n = ord.size();
    keys.resize(n);
    off.assign(n + 1, 0);
    uint64_t total = 0;
    for (size_t i = 0; i < n; i++) {
        if (ord[i].len < 2) { err = p + ": sequence entry shorter than its terminator"; return false; }
        total += ord[i].len - 2;
    }
    residues.resize(total);
    posOfKey.reserve(n * 2);
    for (size_t i = 0; i < n; i++) {
        keys[i] = ord[i].key;
        if (i && keys[i] <= keys[i - 1]) keysAscend = false;
        posOfKey[ord[i].key] = (uint32_t)i;
        const uint64_t l = ord[i].len - 2;
        memcpy(&residues[off[i]], db.data.data() + ord[i].off, l);
        off[i + 1] = off[i] + l;
    }
    // <db>.lookup: key \t name \t set ; clusterhits takes gene index and strand from the last three '_' fields of the name
    // (src/util/ClusterHits.cpp:328-350)
    std::string lk;
    if (!readFile(p + ".lookup", lk)) { err = "cannot read " + p + ".lookup (was the DB made by createsetdb?)"; return false; }
    names.assign(n, std::string());
    setOf.assign(n, 0);
    genePos.assign(n, 0);
    strand.assign(n, 0);
    std::vector<uint32_t> rawSet(n, 0);
    std::vector<char> seen(n, 0);
    std::map<uint32_t, uint32_t> setCount;
    std::vector<std::string> cols;
    const char *q = lk.c_str();
    while (*q) {
        const char *ls = q;
        while (*q && *q != '\n') q++;
        std::string line(ls, q);
        if (*q == '\n') q++;
        if (line.empty()) continue;
        size_t t1 = line.find('\t'), t2 = t1 == std::string::npos ? t1 : line.find('\t', t1 + 1);
        if (t2 == std::string::npos) { err = "malformed lookup line in " + p + ".lookup: " + line; return false; }
        const uint32_t key = (uint32_t)strtoul(line.c_str(), nullptr, 10);
        auto it = posOfKey.find(key);
        if (it == posOfKey.end()) continue;
        const uint32_t pos = it->second;
        names[pos] = line.substr(t1 + 1, t2 - t1 - 1);
        rawSet[pos] = (uint32_t)strtoul(line.c_str() + t2 + 1, nullptr, 10);
        seen[pos] = 1;
        setCount[rawSet[pos]]++;
        splitUnderscore(names[pos], cols);
        if (cols.size() < 4) { err = "Invalid lookup record (need <id>_<geneIdx>_<start>_<end>): " + names[pos]; return false; }
        const uint64_t start = fastAtoi(cols[cols.size() - 2]), end = fastAtoi(cols.back());
        genePos[pos] = (uint32_t)fastAtoi(cols[cols.size() - 3]);
        strand[pos] = (int)start < (int)end ? 1 : 0;
    }
    for (size_t i = 0; i < n; i++)
        if (!seen[i]) { err = p + ".lookup has no record for key " + std::to_string(keys[i]); return false; }
    std::unordered_map<uint32_t, uint32_t> dense;
    setKey.clear();
    setSize.clear();
    for (auto &kv : setCount) {     // std::map: ascending set key
        dense[kv.first] = (uint32_t)setKey.size();
        setKey.push_back(kv.first);
        setSize.push_back(kv.second);
    }
    for (size_t i = 0; i < n; i++) setOf[i] = dense[rawSet[i]];
    // <db>_set_size (result2stats) wins over the lookup count when present
    if (fileExists(p + "_set_size.index")) {
        Db ss;
        std::string e2;
        if (ss.open(p + "_set_size", e2))
            for (const Entry &e : ss.index) {
                auto it = dense.find(e.key);
                if (it != dense.end()) setSize[it->second] = (uint32_t)strtoul(ss.ptr(e), nullptr, 10);
            }
    }
    setSource.assign(setKey.size(), std::string());
    std::string src;
    if (readFile(p + ".source", src)) {
        const char *s = src.c_str();
        while (*s) {
            const char *ls = s;
            while (*s && *s != '\n') s++;
            std::string line(ls, s);
            if (*s == '\n') s++;
            size_t t = line.find('\t');
            if (t == std::string::npos) continue;
            auto it = dense.find((uint32_t)strtoul(line.c_str(), nullptr, 10));
            if (it != dense.end()) setSource[it->second] = line.substr(t + 1);
        }
    }
    return true;
}

void DbWriter::add(uint32_t key, const char *p, size_t n) {
    Entry e;
    e.key = key;
    e.off = data_.size();
    e.len = n + 1;
    data_.append(p, n);
    data_.push_back('\0');
    index_.push_back(e);
}

void setCompressedOutput(bool on) { g_compressOutput = on; }

bool DbWriter::close(std::string &err) {
    if (g_compressOutput) {
        // every entry becomes [u32 n][zstd frame of n bytes][0]; the index keeps the data offset of the record and the ORIGINAL length
        ZstdApi &z = zstd();
        if (!z.compress || !z.compressBound || !z.isError) { err = "--compressed 1: libzstd.so.1 could not be loaded"; return false; }
        std::string packed;
        std::string frame;
        for (Entry &e : index_) {
            const size_t n = (size_t)e.len - 1;
            frame.resize(z.compressBound(n));
            const size_t got = z.compress(&frame[0], frame.size(), data_.data() + e.off, n, 3);      // level 3 (Parameters: --compress-level... default)
            if (z.isError(got)) { err = "zstd compression failed for " + path_; return false; }
            const uint32_t n32 = (uint32_t)got;
            e.off = packed.size();
            packed.append(reinterpret_cast<const char *>(&n32), 4);
            packed.append(frame.data(), got);
            packed.push_back('\0');
        }
        data_.swap(packed);
        dbtype_ = (int)((uint32_t)dbtype_ | 0x80000000u);
    }
    FILE *f = fopen(path_.c_str(), "wb");
    if (!f) { err = "cannot write " + path_; return false; }
    if (!data_.empty() && fwrite(data_.data(), 1, data_.size(), f) != data_.size()) { fclose(f); err = "short write on " + path_; return false; }
    fclose(f);
    std::stable_sort(index_.begin(), index_.end(), [](const Entry &a, const Entry &b) { return a.key < b.key; });
    std::string idx;
    idx.reserve(index_.size() * 24);
    char buf[96];
    for (const Entry &e : index_) {
        int n = snprintf(buf, sizeof buf, "%u\t%llu\t%llu\n", e.key, (unsigned long long)e.off, (unsigned long long)e.len);
        idx.append(buf, (size_t)n);
    }
    f = fopen((path_ + ".index").c_str(), "wb");
    if (!f) { err = "cannot write " + path_ + ".index"; return false; }
    if (!idx.empty()) fwrite(idx.data(), 1, idx.size(), f);
    fclose(f);
    f = fopen((path_ + ".dbtype").c_str(), "wb");
    if (!f) { err = "cannot write " + path_ + ".dbtype"; return false; }
    const uint32_t t = (uint32_t)dbtype_;
    fwrite(&t, 4, 1, f);
    fclose(f);
    return true;
}

}  // namespace sddb
