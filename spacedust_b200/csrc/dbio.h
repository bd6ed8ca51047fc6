// MMseqs2 flat-file DB access for the sdb200 front-end (the on-disk contract of SURVEY.md section 8b):
//   <db>        entries, each terminated by '\0' (sequence entries: residues + "\n\0", lib/mmseqs/src/commons/DBReader.h:229-231)
//   <db>.index  ASCII "key\toffset\tlength\n", length includes the '\0'
//   <db>.dbtype 4-byte little-endian int (lib/mmseqs/src/commons/Parameters.h:68-94)
//   <db>.lookup "key\tname\tfileNo\n", <db>.source "fileNo\tfilename\n"
// Readers accept merged DBs and per-thread split files <db>.0 .. <db>.N-1 (DBWriter.cpp:531-623, FileUtil::findDatafiles).
#ifndef SD_DBIO_H
#define SD_DBIO_H
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

namespace sddb {

enum { DBTYPE_AMINO_ACIDS = 0, DBTYPE_ALIGNMENT_RES = 5, DBTYPE_PREFILTER_RES = 7, DBTYPE_GENERIC_DB = 12 };

struct Entry {
    uint32_t key;
    uint64_t off;
    uint64_t len;   // includes the terminating '\0'
};

struct Db {
    std::string data;
    std::vector<Entry> index;   // file order of <db>.index
    int dbtype = -1;
    bool open(const std::string &path, std::string &err);
    bool inflate(const std::string &path, std::string &err);     // compressed DB -> plain layout (called by open)
    // entry payload without the trailing '\0'
    const char *ptr(const Entry &e) const { return data.data() + e.off; }
    size_t size(const Entry &e) const { return e.len ? (size_t)e.len - 1 : 0; }
};

// A set DB as `createsetdb` leaves it, laid out for the C ABI: proteins in ascending data offset (the order
// DBReader's LINEAR_ACCCESS ids follow), residues concatenated without the "\n\0" terminators.
struct SetDb {
    std::string path;
    std::vector<uint32_t> keys;             // position -> DB key
    bool keysAscend = true;                 // keys ascend with the data offset (what createsetdb / createdb write)
    std::unordered_map<uint32_t, uint32_t> posOfKey;
    std::string residues;
    std::vector<uint64_t> off;              // n+1
    std::vector<std::string> names;         // lookup name per position
    std::vector<uint32_t> setOf, genePos;   // dense set index, gene index inside the set
    std::vector<uint8_t> strand;            // 1: start < end
    std::vector<uint32_t> setKey;           // dense set index -> set key (ascending)
    std::vector<uint32_t> setSize;
    std::vector<std::string> setSource;     // <db>.source name per dense set
    uint32_t n() const { return (uint32_t)keys.size(); }
    bool open(const std::string &path, std::string &err);
};

// --compressed 1 of the calling module: every DbWriter closed afterwards writes zstd frames ([u32 n][frame][0], index lengths =
// original sizes, dbtype bit 31: DBWriter.cpp:326-395).  The frames need not equal the reference's byte for byte, only expand to the same.
void setCompressedOutput(bool on);

// Collects entries and writes <path>, <path>.index (sorted by key, stable) and <path>.dbtype like a merged DBWriter result.
class DbWriter {
public:
    DbWriter(const std::string &path, int dbtype) : path_(path), dbtype_(dbtype) {}
    void add(uint32_t key, const char *p, size_t n);
    void add(uint32_t key, const std::string &s) { add(key, s.data(), s.size()); }
    bool close(std::string &err);
private:
    std::string path_;
    int dbtype_;
    std::string data_;
    std::vector<Entry> index_;
};

bool readFile(const std::string &path, std::string &out);
bool fileExists(const std::string &path);

}  // namespace sddb
#endif
