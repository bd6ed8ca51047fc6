// Host-only modules that follow the clustering step in data/clustersearch.sh:151-165 -- pure text work on the reference's DB
// files, no GPU involved:
//   sdb200 summarizeresults <qSetDB> <tSetDB> <clustersDB> <out.tsv>     src/util/SummarizeResults.cpp:14-130
//   sdb200 filterdb <inDB> <outDB> --trim-to-one-column                   first column of every line (clusters -> clu_to_seq)
//   sdb200 swapdb <inDB> <outDB>                                          invert a one-column DB (clu_to_seq -> seq_to_clu)
// and the two text modules of the iterative search glue (lib/mmseqs/data/workflow/blastpgp.sh:83-124; the profile search
// itself is not built yet):
//   sdb200 subtractdbs <leftDB> <rightDB> <outDB> [--e-profile X]         lib/mmseqs/src/util/subtractdbs.cpp:20-110
//   sdb200 mergedbs <queryDB> <outDB> <db1> <db2> ...                     lib/mmseqs/src/util/mergedbs.cpp:8-69
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <string>
#include <vector>

#include "dbio.h"

namespace {

[[noreturn]] void fail(const std::string &m) {
    fprintf(stderr, "Error: %s\n", m.c_str());
    exit(EXIT_FAILURE);
}

std::vector<std::string> positional(int argc, char **argv, std::map<std::string, std::string> *flags = nullptr) {
    std::vector<std::string> pos;
    for (int i = 2; i < argc; i++) {
        const char *s = argv[i];
        if (s[0] == '-' && s[1] != '\0') {
            if (!strcmp(s, "--trim-to-one-column")) {           // the one flag of this family without a value
                if (flags) (*flags)[s] = "1";
                continue;
            }
            if (i + 1 >= argc) fail(std::string("flag ") + s + " needs a value");
            if (flags) (*flags)[s] = argv[i + 1];
            i++;
        } else pos.push_back(s);
    }
    return pos;
}

// <db>.lookup: key -> name ; <db>.source: set key -> file name
std::map<uint32_t, std::string> readKeyed(const std::string &path, int column) {
    std::string text;
    if (!sddb::readFile(path, text)) fail("cannot read " + path);
    std::map<uint32_t, std::string> out;
    size_t p = 0;
    while (p < text.size()) {
        size_t e = text.find('\n', p);
        if (e == std::string::npos) e = text.size();
        const std::string line = text.substr(p, e - p);
        p = e + 1;
        if (line.empty()) continue;
        std::vector<std::string> f;
        size_t b = 0;
        while (true) {
            size_t t = line.find('\t', b);
            if (t == std::string::npos) { f.push_back(line.substr(b)); break; }
            f.push_back(line.substr(b, t - b));
            b = t + 1;
        }
        if ((int)f.size() > column) out[(uint32_t)strtoul(f[0].c_str(), nullptr, 10)] = f[column];
    }
    return out;
}

}  // namespace

// first column of every line of every entry
int filterdbTrimToOneColumn(const std::string &in, const std::string &out) {
    sddb::Db db;
    std::string err;
    if (!db.open(in, err)) fail(err);
    sddb::DbWriter w(out, db.dbtype >= 0 ? db.dbtype : sddb::DBTYPE_GENERIC_DB);
    std::string buf;
    for (const sddb::Entry &e : db.index) {
        buf.clear();
        const char *p = db.ptr(e), *end = p + db.size(e);
        while (p < end) {
            const char *ls = p;
            while (p < end && *p != '\n') p++;
            const char *c = ls;
            while (c < p && *c != '\t' && *c != ' ') c++;
            buf.append(ls, (size_t)(c - ls));
            buf.push_back('\n');
            if (p < end) p++;
        }
        w.add(e.key, buf);
    }
    if (!w.close(err)) fail(err);
    return EXIT_SUCCESS;
}

// every line value v of entry k becomes a line k of entry v (ascending k inside an entry)
int swapOneColumnDb(const std::string &in, const std::string &out) {
    sddb::Db db;
    std::string err;
    if (!db.open(in, err)) fail(err);
    std::vector<std::pair<uint32_t, uint32_t>> pairs;       // (value, key)
    for (const sddb::Entry &e : db.index) {
        const char *p = db.ptr(e), *end = p + db.size(e);
        while (p < end) {
            const char *ls = p;
            while (p < end && *p != '\n') p++;
            if (p > ls) pairs.emplace_back((uint32_t)strtoul(ls, nullptr, 10), e.key);
            if (p < end) p++;
        }
    }
    std::sort(pairs.begin(), pairs.end());
    sddb::DbWriter w(out, db.dbtype >= 0 ? db.dbtype : sddb::DBTYPE_GENERIC_DB);
    std::string buf;
    for (size_t i = 0; i < pairs.size();) {
        size_t j = i;
        buf.clear();
        while (j < pairs.size() && pairs[j].first == pairs[i].first) { buf += std::to_string(pairs[j].second) + "\n"; j++; }
        w.add(pairs[i].first, buf);
        i = j;
    }
    if (!w.close(err)) fail(err);
    return EXIT_SUCCESS;
}

int modFilterdb(int argc, char **argv) {
    std::map<std::string, std::string> flags;
    std::vector<std::string> pos = positional(argc, argv, &flags);
    if (pos.size() != 2) fail("usage: sdb200 filterdb <inDB> <outDB> --trim-to-one-column");
    if (!flags.count("--trim-to-one-column")) fail("filterdb: only --trim-to-one-column is implemented");
    return filterdbTrimToOneColumn(pos[0], pos[1]);
}

int modSwapdb(int argc, char **argv) {
    std::vector<std::string> pos = positional(argc, argv);
    if (pos.size() != 2) fail("usage: sdb200 swapdb <inDB> <outDB>");
    return swapOneColumnDb(pos[0], pos[1]);
}

int modSummarizeresults(int argc, char **argv) {
    std::vector<std::string> pos = positional(argc, argv);
    if (pos.size() != 4) fail("usage: sdb200 summarizeresults <querySetDB> <targetSetDB> <clustersDB> <output.tsv>");
    const std::map<uint32_t, std::string> qName = readKeyed(pos[0] + ".lookup", 1), tName = readKeyed(pos[1] + ".lookup", 1);
    const std::map<uint32_t, std::string> qSrc = readKeyed(pos[0] + ".source", 1), tSrc = readKeyed(pos[1] + ".source", 1);
    sddb::Db clu, hdr;
    std::string err;
    if (!clu.open(pos[2], err)) fail(err);
    if (!hdr.open(pos[2] + "_h", err)) fail(err);
    // the alignment entries are looked up by key (clusterhits gives header and body the same key)
    std::map<uint32_t, const sddb::Entry *> byKey;
    for (const sddb::Entry &e : clu.index) byKey[e.key] = &e;
    FILE *f = fopen(pos[3].c_str(), "wb");
    if (!f) fail("cannot write " + pos[3]);
    auto fields = [](const std::string &line) {
        std::vector<std::string> v;
        size_t b = 0;
        while (true) {
            size_t t = line.find('\t', b);
            if (t == std::string::npos) { v.push_back(line.substr(b)); break; }
            v.push_back(line.substr(b, t - b));
            b = t + 1;
        }
        return v;
    };
    auto nameOf = [](const std::map<uint32_t, std::string> &m, uint32_t k) {
        auto it = m.find(k);
        return it == m.end() ? std::string() : it->second;
    };
    std::string out;
    for (const sddb::Entry &he : hdr.index) {
        auto it = byKey.find(he.key);
        if (it == byKey.end()) fail("clusters DB has no entry for header key " + std::to_string(he.key));
        const std::string head(hdr.ptr(he), hdr.size(he));
        size_t hp = 0;
        while (hp < head.size()) {
            size_t e = head.find('\n', hp);
            if (e == std::string::npos) e = head.size();
            const std::vector<std::string> h = fields(head.substr(hp, e - hp));
            hp = e + 1;
            if (h.size() < 5) fail("Invalid alignment result record");
            out += "#" + std::to_string(he.key) + "\t" + nameOf(qSrc, (uint32_t)strtoul(h[0].c_str(), nullptr, 10)) + "\t" +
                   nameOf(tSrc, (uint32_t)strtoul(h[1].c_str(), nullptr, 10)) + "\t" + h[2] + "\t" + h[3] + "\t" + h[4] + "\n";
            const std::string body(clu.ptr(*it->second), clu.size(*it->second));
            size_t bp = 0;
            while (bp < body.size()) {
                size_t be = body.find('\n', bp);
                if (be == std::string::npos) be = body.size();
                const std::string line = body.substr(bp, be - bp);
                bp = be + 1;
                if (line.empty()) continue;
                const std::vector<std::string> a = fields(line);
                if (a.size() < 10) fail("Invalid alignment result record");
                out += ">" + nameOf(qName, (uint32_t)strtoul(a[0].c_str(), nullptr, 10)) + "\t" + nameOf(tName, (uint32_t)strtoul(a[1].c_str(), nullptr, 10));
                for (size_t c = 2; c < a.size(); c++) out += "\t" + a[c];
                out += "\n";
            }
        }
        if (out.size() > (1u << 22)) { fwrite(out.data(), 1, out.size(), f); out.clear(); }
    }
    if (!out.empty()) fwrite(out.data(), 1, out.size(), f);
    fclose(f);
    return EXIT_SUCCESS;
}


namespace {
// key (first column) and, for alignment lines (>= 10 columns), the E-value (4th column) of one result line
void lineKeyEvalue(const char *ls, const char *le, uint32_t &key, double &evalue) {
    key = (uint32_t)strtoul(ls, nullptr, 10);
    evalue = 0.0;
    int columns = 0;
    const char *col4 = nullptr;
    const char *p = ls;
    while (p < le) {
        while (p < le && (*p == '\t' || *p == ' ')) p++;
        if (p >= le) break;
        columns++;
        if (columns == 4) col4 = p;
        while (p < le && *p != '\t' && *p != ' ') p++;
    }
    if (columns >= 10 && col4) evalue = strtod(col4, nullptr);
}
}  // namespace

// lines of the left entry whose key does not occur in the right entry of the same DB key (both sides only count lines with
// E-value <= threshold; prefilter lines have no E-value and always count)
int modSubtractdbs(int argc, char **argv) {
    std::map<std::string, std::string> flags;
    std::vector<std::string> pos = positional(argc, argv, &flags);
    if (pos.size() != 3) fail("usage: sdb200 subtractdbs <leftDB> <rightDB> <outDB> [-e X] [--e-profile X]");
    // subtractdbs.cpp:17: the threshold is the smaller of -e (default 0.001) and --e-profile (default 0.001, Parameters.cpp)
    const double evalThr = flags.count("-e") ? strtod(flags["-e"].c_str(), nullptr) : 0.001;
    const double evalProfile = flags.count("--e-profile") ? strtod(flags["--e-profile"].c_str(), nullptr) : 0.001;
    const double thr = std::min(evalThr, evalProfile);
    sddb::setCompressedOutput(flags.count("--compressed") && atoi(flags["--compressed"].c_str()) != 0);
    sddb::Db left, right;
    std::string err;
    if (!left.open(pos[0], err)) fail(err);
    if (!right.open(pos[1], err)) fail(err);
    std::map<uint32_t, const sddb::Entry *> rightByKey;
    for (const sddb::Entry &e : right.index) rightByKey[e.key] = &e;
    sddb::DbWriter w(pos[2], left.dbtype >= 0 ? left.dbtype : sddb::DBTYPE_GENERIC_DB);
    std::string buf;
    auto forLines = [](const char *p, const char *end, const std::function<void(const char *, const char *)> &fn) {
        while (p < end) {
            const char *ls = p;
            while (p < end && *p != '\n') p++;
            const char *le = p;
            if (p < end) p++;
            fn(ls, le < end ? le + 1 : le);
        }
    };
    for (const sddb::Entry &e : left.index) {
        std::map<uint32_t, bool> keep;
        const char *lp = left.ptr(e), *lend = lp + left.size(e);
        forLines(lp, lend, [&](const char *ls, const char *le) {
            uint32_t k; double ev;
            lineKeyEvalue(ls, le, k, ev);
            if (ev <= thr) keep[k] = true;
        });
        auto it = rightByKey.find(e.key);
        if (it != rightByKey.end()) {
            const char *rp = right.ptr(*it->second), *rend = rp + right.size(*it->second);
            forLines(rp, rend, [&](const char *ls, const char *le) {
                uint32_t k; double ev;
                lineKeyEvalue(ls, le, k, ev);
                if (ev <= thr) keep[k] = false;
            });
        }
        buf.clear();
        forLines(lp, lend, [&](const char *ls, const char *le) {
            const uint32_t k = (uint32_t)strtoul(ls, nullptr, 10);
            auto kt = keep.find(k);
            if (kt != keep.end() && kt->second) buf.append(ls, (size_t)(le - ls));
        });
        w.add(e.key, buf);
    }
    if (!w.close(err)) fail(err);
    return EXIT_SUCCESS;
}

// per key of the query DB: the entries of db1, db2, ... concatenated in that order
int modMergedbs(int argc, char **argv) {
    std::vector<std::string> pos = positional(argc, argv);
    if (pos.size() < 4) fail("usage: sdb200 mergedbs <queryDB> <outDB> <db1> <db2> [...]");
    sddb::Db q;
    std::string err;
    if (!q.open(pos[0], err)) fail(err);
    const size_t n = pos.size() - 2;
    std::vector<sddb::Db> dbs(n);
    std::vector<std::map<uint32_t, const sddb::Entry *>> byKey(n);
    for (size_t i = 0; i < n; i++) {
        if (!dbs[i].open(pos[i + 2], err)) fail(err);
        for (const sddb::Entry &e : dbs[i].index) byKey[i][e.key] = &e;
    }
    sddb::DbWriter w(pos[1], dbs[0].dbtype >= 0 ? dbs[0].dbtype : sddb::DBTYPE_GENERIC_DB);
    std::string buf;
    for (const sddb::Entry &qe : q.index) {
        buf.clear();
        for (size_t i = 0; i < n; i++) {
            auto it = byKey[i].find(qe.key);
            if (it == byKey[i].end()) continue;
            buf.append(dbs[i].ptr(*it->second), dbs[i].size(*it->second));
        }
        w.add(qe.key, buf);
    }
    if (!w.close(err)) fail(err);
    return EXIT_SUCCESS;
}
