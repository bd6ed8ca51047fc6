// sdb200 createsetdb <a.faa> [<b.faa> ...] <setDB> <tmpDir>
//
// Native single-pass replacement of the protein-FASTA branch of data/createsetdb.sh (createdb --shuffle 0
// --write-lookup 1, prefixid --tsv, the awk/join/sort pipeline of lines 121-135, tsv2db x2, result2stats): every file
// the script leaves behind is written with the same bytes --
//   setDB{,.index,.dbtype,.lookup,.source}  setDB_h{,.index,.dbtype}
//   setDB_member_to_set{,.index,.dbtype,.tsv}  setDB_set_to_member{,.index,.dbtype,.tsv}  setDB_set_size{,.index,.dbtype}
// (tests/test_frontend_cpu.py diffs them against the reference binary's output).  One set = one input file; genes keep
// file order; lookup names are `<id>_<geneIdxInSet>_<start>_<end>` from Prodigal headers `>id # start # end # strand # ...`,
// start and end swapped on the minus strand.  Nucleotide input / --gff-dir is not implemented and fails loudly.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <dlfcn.h>

#include "dbio.h"

namespace {

[[noreturn]] void fail(const std::string &m) {
    fprintf(stderr, "Error: %s\n", m.c_str());
    exit(EXIT_FAILURE);
}

std::string baseName(const std::string &p) {
    size_t s = p.find_last_of('/');
    return s == std::string::npos ? p : p.substr(s + 1);
}

void writeText(const std::string &path, const std::string &s) {
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) fail("cannot write " + path);
    if (!s.empty()) fwrite(s.data(), 1, s.size(), f);
    fclose(f);
}

// the awk stage: strip blanks, split at '#', drop the last field, swap start/end when the strand field is "-1"
bool prodigalFields(const std::string &header, std::string &id, std::string &start, std::string &end) {
    std::string h;
    for (char c : header) if (c != ' ') h.push_back(c);
    std::vector<std::string> f;
    size_t b = 0;
    while (true) {
        size_t e = h.find_first_of("\t#", b);
        if (e == std::string::npos) { f.push_back(h.substr(b)); break; }
        f.push_back(h.substr(b, e - b));
        b = e + 1;
    }
    if (!f.empty()) f.pop_back();
    if (f.size() < 3) return false;
    id = f[0];
    start = f[1];
    end = f[2];
    if (f.size() >= 4 && f[3] == "-1") std::swap(start, end);
    return true;
}

}  // namespace

int modCreatesetdb(int argc, char **argv) {
    std::vector<std::string> pos;
    for (int i = 2; i < argc; i++) {
        const char *s = argv[i];
        if (s[0] == '-' && s[1] != '\0') {
            if (!strcmp(s, "--gff-dir") || !strcmp(s, "--gff-type")) fail("nucleotide input with --gff-dir is not implemented (protein FASTA only)");
            if (i + 1 >= argc) fail(std::string("flag ") + s + " needs a value");
            if (!strcmp(s, "--compressed") && atoi(argv[i + 1]) != 0) sddb::setCompressedOutput(true);      // zstd frames, like DBWriter
            i++;        // reference flags (--threads, -v, --remove-tmp-files ...) do not change the files written here
        } else pos.push_back(s);
    }
    if (pos.size() < 3) fail("usage: sdb200 createsetdb <in1.faa> [<in2.faa> ...] <setDB> <tmpDir>");
    const std::string out = pos[pos.size() - 2];
    std::vector<std::string> inputs(pos.begin(), pos.end() - 2);

    sddb::DbWriter seqW(out, sddb::DBTYPE_AMINO_ACIDS), hdrW(out + "_h", sddb::DBTYPE_GENERIC_DB);
    sddb::DbWriter m2sW(out + "_member_to_set", sddb::DBTYPE_ALIGNMENT_RES), s2mW(out + "_set_to_member", sddb::DBTYPE_ALIGNMENT_RES),
        sizeW(out + "_set_size", sddb::DBTYPE_GENERIC_DB);
    std::string lookup, source, m2sTsv, s2mTsv;
    uint32_t key = 0;
    for (size_t fi = 0; fi < inputs.size(); fi++) {
        std::string text;
        if (!sddb::readFile(inputs[fi], text)) fail("cannot read " + inputs[fi]);
        if (text.size() >= 2 && (unsigned char)text[0] == 0x1f && (unsigned char)text[1] == 0x8b) {
            // gzip input (createdb reads it through kseq + zlib, createdb.cpp): zlib resolved at run time
            static void *hz = dlopen("libz.so.1", RTLD_NOW);
            using GzOpen = void *(*)(const char *, const char *);
            using GzRead = int (*)(void *, void *, unsigned);
            using GzClose = int (*)(void *);
            GzOpen gzopenFn = hz ? reinterpret_cast<GzOpen>(dlsym(hz, "gzopen")) : nullptr;
            GzRead gzreadFn = hz ? reinterpret_cast<GzRead>(dlsym(hz, "gzread")) : nullptr;
            GzClose gzcloseFn = hz ? reinterpret_cast<GzClose>(dlsym(hz, "gzclose")) : nullptr;
            if (!gzopenFn || !gzreadFn || !gzcloseFn) fail(inputs[fi] + ": gzip input needs libz.so.1, which could not be loaded");
            void *gz = gzopenFn(inputs[fi].c_str(), "rb");
            if (!gz) fail("cannot read " + inputs[fi]);
            text.clear();
            std::vector<char> chunk(1 << 20);
            int got;
            while ((got = gzreadFn(gz, chunk.data(), (unsigned)chunk.size())) > 0) text.append(chunk.data(), (size_t)got);
            gzcloseFn(gz);
            if (got < 0) fail(inputs[fi] + ": corrupt gzip stream");
        }
        source += std::to_string(fi) + "\t" + baseName(inputs[fi]) + "\n";
        std::string members;
        uint32_t geneIdx = 0;
        size_t p = 0;
        const size_t n = text.size();
        while (p < n && text[p] != '>') p++;       // kseq skips to the first record marker
        while (p < n) {
            // header line: name up to the first blank, rest of the line is the comment (kseq); createdb stores "name comment\n"
            size_t le = text.find('\n', p);
            if (le == std::string::npos) le = n;
            size_t he = le;
            if (he > p && text[he - 1] == '\r') he--;
            std::string line = text.substr(p + 1, he - p - 1);
            size_t sp = line.find_first_of(" \t");
            std::string header = sp == std::string::npos ? line : line.substr(0, sp) + " " + line.substr(sp + 1);
            // sequence lines until the next '>' at a line start
            std::string seq;
            size_t q = le < n ? le + 1 : n;
            while (q < n && text[q] != '>') {
                size_t e = text.find('\n', q);
                if (e == std::string::npos) e = n;
                for (size_t k = q; k < e; k++) {
                    const unsigned char c = (unsigned char)text[k];
                    if (c > ' ' && c < 127) seq.push_back((char)c);     // isgraph
                }
                q = e < n ? e + 1 : n;
            }
            p = q;
            if (seq.empty()) fail(inputs[fi] + ": record '" + header + "' has no sequence");
            if (seq.find_first_not_of("ACGTUNacgtun*") == std::string::npos && seq.size() > 20)
                fail(inputs[fi] + ": looks like nucleotides; nucleotide input with --gff-dir is not implemented");
            std::string id, start, end;
            if (!prodigalFields(header, id, start, end))
                fail(inputs[fi] + ": header '" + header + "' is not in Prodigal form `id # start # end # strand # ...`");
            seq.push_back('\n');
            header.push_back('\n');
            seqW.add(key, seq);
            hdrW.add(key, header);
            lookup += std::to_string(key) + "\t" + id + "_" + std::to_string(geneIdx) + "_" + start + "_" + end + "\t" + std::to_string(fi) + "\n";
            const std::string ks = std::to_string(key), fs = std::to_string(fi);
            m2sTsv += ks + "\t" + fs + "\n";
            s2mTsv += fs + "\t" + ks + "\n";
            m2sW.add(key, fs + "\n");
            members += ks + "\n";
            key++;
            geneIdx++;
        }
        if (geneIdx == 0) fail(inputs[fi] + ": no FASTA record");
        s2mW.add((uint32_t)fi, members);
        sizeW.add((uint32_t)fi, std::to_string(geneIdx) + "\n");
    }
    std::string err;
    if (!seqW.close(err) || !hdrW.close(err) || !m2sW.close(err) || !s2mW.close(err) || !sizeW.close(err)) fail(err);
    writeText(out + ".lookup", lookup);
    writeText(out + ".source", source);
    writeText(out + "_member_to_set.tsv", m2sTsv);
    writeText(out + "_set_to_member.tsv", s2mTsv);
    return EXIT_SUCCESS;
}
