// The aggregation modules between `align` and `summarizeresults` in data/clustersearch.sh:119-146 as stand-alone sdb200 modules on
// the reference's DB files, so that a tmp directory can be resumed or mixed module by module with the reference binary:
//   sdb200 prefixid <resultDB> <result_prefixed>                                    lib/mmseqs/src/util/prefixid.cpp:10-84
//   sdb200 besthitbyset <qSetDB> <tSetDB> <result_prefixed> <aggregate>             src/util/besthitbyset.cpp:41-144, Aggregation.cpp:24-91
//   sdb200 mergeresultsbyset <qSetDB>_set_to_member <aggregate> <aggregate_merged>  lib/mmseqs/src/util/mergeresultsbyset.cpp:11-71
//   sdb200 combinehits <qSetDB> <tSetDB> <aggregate_merged> <matches> <tmp>         src/util/combinehits.cpp:74-234, Aggregation.cpp:93-157
//   sdb200 clusterhits <qSetDB> <tSetDB> <matches> <clusters>                       src/util/ClusterHits.cpp:214-490
// The first four are text work on the host (the fused `sdb200 clustersearch` does the same arithmetic on binary records inside
// sd_cluster_run); clusterhits parses the `matches` / `matches_h` records and runs the merge loop on the GPU (sd_clusterhits_run).
// Entry keys of `matches` and `clusters` are running numbers in (query set, target set) order -- what the reference writes with
// one thread (with more threads its counters are per thread, Aggregation.cpp:113, ClusterHits.cpp:284).
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/sdgpu.h"
#include "dbio.h"

namespace {

[[noreturn]] void fail(const std::string &m) {
    fprintf(stderr, "Error: %s\n", m.c_str());
    exit(EXIT_FAILURE);
}

struct Cli {
    std::vector<std::string> pos;
    std::map<std::string, std::string> flags;
    double num(const char *f, double dflt) const {
        auto it = flags.find(f);
        return it == flags.end() ? dflt : strtod(it->second.c_str(), nullptr);
    }
};

Cli parseCli(int argc, char **argv) {
    Cli c;
    for (int i = 2; i < argc; i++) {
        const char *s = argv[i];
        if (s[0] == '-' && s[1] != '\0' && !(s[1] >= '0' && s[1] <= '9')) {
            if (i + 1 >= argc) fail(std::string("flag ") + s + " needs a value");
            c.flags[s] = argv[i + 1];
            i++;
        } else c.pos.push_back(s);
    }
    return c;
}

// Flags of these modules that are implemented, inert here (host-runtime knobs), or accepted only with the value the workflow passes.
void checkFlags(const Cli &c, std::initializer_list<const char *> implemented) {
    static const char *inert[] = {"--threads", "-v", "--compressed", "--db-load-mode", "--tsv", nullptr};
    for (const auto &kv : c.flags) {
        bool ok = false;
        for (const char *f : implemented) ok = ok || kv.first == f;
        for (int i = 0; inert[i] && !ok; i++) ok = kv.first == inert[i];
        if (kv.first == "--compressed") sddb::setCompressedOutput(kv.second != "0");
        if (kv.first == "--tsv" && kv.second != "0") fail("--tsv 1 is not implemented");
        if (kv.first == "--simple-best-hit" && kv.second != "1") fail("--simple-best-hit 0 is not implemented (only 1, the value the clustersearch workflow passes)");
        if (kv.first == "--suboptimal-hits" && kv.second != "0") fail("--suboptimal-hits > 0 is not implemented");
        if (kv.first == "--aggregation-mode" && kv.second != "0") fail("--aggregation-mode != 0 is not implemented");
        if (kv.first == "--cluster-use-weight" && kv.second != "0") fail("--cluster-use-weight 1 is not implemented");
        if (kv.first == "--cluster-use-weight") ok = true;
        if (kv.first == "--db-output" && kv.second != "1") fail("--db-output 0 is not implemented (the modules write DBs, as the clustersearch workflow asks)");
        if (kv.first == "--simple-best-hit" || kv.first == "--suboptimal-hits" || kv.first == "--aggregation-mode" || kv.first == "--db-output") ok = true;
        if (!ok) fail("unknown or unsupported flag " + kv.first + " (it would be ignored; refusing to run)");
    }
}

sddb::Db openDb(const std::string &path) {
    sddb::Db db;
    std::string err;
    if (!db.open(path, err)) fail(err);
    return db;
}

// entries in the order DBReader's LINEAR_ACCCESS ids follow: ascending data offset
std::vector<sddb::Entry> byOffset(const sddb::Db &db) {
    std::vector<sddb::Entry> e = db.index;
    std::stable_sort(e.begin(), e.end(), [](const sddb::Entry &a, const sddb::Entry &b) { return a.off < b.off; });
    return e;
}

// key -> unsigned value of a one-number-per-entry DB (_member_to_set, _set_size)
std::unordered_map<uint32_t, uint32_t> numberDb(const std::string &path) {
    const sddb::Db db = openDb(path);
    std::unordered_map<uint32_t, uint32_t> m;
    m.reserve(db.index.size() * 2);
    for (const sddb::Entry &e : db.index) m[e.key] = (uint32_t)strtoul(std::string(db.ptr(e), db.size(e)).c_str(), nullptr, 10);
    return m;
}

struct Line {
    const char *p;
    size_t n;      // without the '\n'
};

// the lines of an entry (a last line without '\n' counts, as std::getline / Util::skipLine see it)
std::vector<Line> linesOf(const char *p, size_t n) {
    std::vector<Line> out;
    size_t b = 0;
    while (b < n) {
        const char *nl = (const char *)memchr(p + b, '\n', n - b);
        const size_t e = nl ? (size_t)(nl - p) : n;
        out.push_back(Line{p + b, e - b});
        b = e + 1;
    }
    return out;
}

std::vector<std::string> splitTabs(const Line &l) {
    std::vector<std::string> f;
    size_t b = 0;
    while (true) {
        const char *t = (const char *)memchr(l.p + b, '\t', l.n - b);
        if (!t) { f.emplace_back(l.p + b, l.n - b); break; }
        f.emplace_back(l.p + b, (size_t)(t - (l.p + b)));
        b = (size_t)(t - l.p) + 1;
    }
    return f;
}

std::string fmt3E(double x) {     // SSTR(double) = fmt "{:.3E}" (Util.cpp:657-660) = printf "%.3E"
    char b[48];
    snprintf(b, sizeof b, "%.3E", x);
    return b;
}

// Aggregation::buildMap (Aggregation.cpp:24-47): the non-empty lines of an entry grouped by the SET of the target in column 2
std::map<uint32_t, std::vector<std::vector<std::string>>> groupByTargetSet(const char *p, size_t n, const std::unordered_map<uint32_t, uint32_t> &memberToSet) {
    std::map<uint32_t, std::vector<std::vector<std::string>>> g;
    for (const Line &l : linesOf(p, n)) {
        if (l.n == 0) continue;
        std::vector<std::string> cols = splitTabs(l);
        if (cols.size() < 2) fail("invalid result record: " + std::string(l.p, l.n));
        const uint32_t tKey = (uint32_t)strtoul(cols[1].c_str(), nullptr, 10);
        auto it = memberToSet.find(tKey);
        if (it == memberToSet.end()) fail("Invalid target database key " + cols[1] + ".");
        g[it->second].push_back(std::move(cols));
    }
    return g;
}

double computeLogPval(double eval) {      // besthitbyset.cpp:10-20 with logCalibration = log(1)
    const double logCalibration = std::log(1);
    if (eval == 0) return std::log(DBL_MIN) - logCalibration;
    if (eval > 0 && eval < 10e-4) return std::log(eval) - logCalibration;
    return std::log(1 - std::exp(-eval)) - logCalibration;
}

}  // namespace

int modPrefixid(int argc, char **argv) {
    const Cli c = parseCli(argc, argv);
    checkFlags(c, {});
    if (c.pos.size() != 2) fail("usage: sdb200 prefixid <resultDB> <outDB>");
    const sddb::Db in = openDb(c.pos[0]);
    sddb::DbWriter w(c.pos[1], in.dbtype >= 0 ? in.dbtype : sddb::DBTYPE_ALIGNMENT_RES);
    std::string buf;
    for (const sddb::Entry &e : byOffset(in)) {
        buf.clear();
        const std::string key = std::to_string(e.key);
        for (const Line &l : linesOf(in.ptr(e), in.size(e))) {
            buf += key;
            buf += '\t';
            buf.append(l.p, l.n);
            buf += '\n';
        }
        w.add(e.key, buf);
    }
    std::string err;
    if (!w.close(err)) fail(err);
    return EXIT_SUCCESS;
}

int modBesthitbyset(int argc, char **argv) {
    const Cli c = parseCli(argc, argv);
    checkFlags(c, {});
    if (c.pos.size() != 4) fail("usage: sdb200 besthitbyset <querySetDB> <targetSetDB> <result_prefixed> <aggregateDB>");
    const auto memberToSet = numberDb(c.pos[1] + "_member_to_set");
    const auto setSize = numberDb(c.pos[1] + "_set_size");
    const sddb::Db in = openDb(c.pos[2]);
    sddb::DbWriter w(c.pos[3], sddb::DBTYPE_ALIGNMENT_RES);
    std::string buf;
    for (const sddb::Entry &e : byOffset(in)) {
        buf.clear();
        auto groups = groupByTargetSet(in.ptr(e), in.size(e), memberToSet);
        for (auto &kv : groups) {
            if (!setSize.count(kv.first)) fail("Invalid target size database key " + std::to_string(kv.first) + ".");
            // simple best hit (besthitbyset.cpp:69-75,109-111): the first line with the smallest E-value; its column 3 becomes the
            // log P-value of that E-value
            double bestEval = DBL_MAX;
            const std::vector<std::string> *best = nullptr;
            for (const auto &cols : kv.second) {
                if (cols.size() < 5) fail("invalid result record in entry " + std::to_string(e.key));
                const double eval = strtod(cols[4].c_str(), nullptr);
                if (eval < bestEval) { bestEval = eval; best = &cols; }
            }
            if (best) {
                const std::string logP = fmt3E(computeLogPval(bestEval));
                for (size_t i = 0; i < best->size(); i++) {
                    buf += (i == 2) ? logP : (*best)[i];
                    if (i + 1 != best->size()) buf += '\t';
                }
            }
            buf += '\n';
        }
        w.add(e.key, buf);
    }
    std::string err;
    if (!w.close(err)) fail(err);
    return EXIT_SUCCESS;
}

int modMergeresultsbyset(int argc, char **argv) {
    const Cli c = parseCli(argc, argv);
    checkFlags(c, {});
    if (c.pos.size() != 3) fail("usage: sdb200 mergeresultsbyset <setDB>_set_to_member <resultDB> <outDB>");
    const sddb::Db sets = openDb(c.pos[0]);
    const sddb::Db res = openDb(c.pos[1]);
    std::unordered_map<uint32_t, const sddb::Entry *> byKey;
    byKey.reserve(res.index.size() * 2);
    for (const sddb::Entry &e : res.index) byKey[e.key] = &e;
    // dbtype of the result DB with the "needs source" extension bit (mergeresultsbyset.cpp:34-35, DBReader.h:375-377)
    const int dbtype = (res.dbtype >= 0 ? res.dbtype : sddb::DBTYPE_ALIGNMENT_RES) | (2 << 16);
    sddb::DbWriter w(c.pos[2], dbtype);
    std::string buf;
    for (const sddb::Entry &s : byOffset(sets)) {
        buf.clear();
        for (const Line &l : linesOf(sets.ptr(s), sets.size(s))) {
            size_t k = 0;
            while (k < l.n && l.p[k] != '\t' && l.p[k] != ' ' && l.p[k] != '\n') k++;       // Util::parseKey
            const uint32_t key = (uint32_t)strtoul(std::string(l.p, k).c_str(), nullptr, 10);
            auto it = byKey.find(key);
            if (it == byKey.end()) fail("Invalid key " + std::to_string(key) + " in entry " + std::to_string(s.key) + ".");
            buf.append(res.ptr(*it->second), res.size(*it->second));
        }
        w.add(s.key, buf);
    }
    std::string err;
    if (!w.close(err)) fail(err);
    return EXIT_SUCCESS;
}

int modCombinehits(int argc, char **argv) {
    const Cli c = parseCli(argc, argv);
    checkFlags(c, {"--alpha", "--filter-self-match"});
    if (c.pos.size() != 5 && c.pos.size() != 4) fail("usage: sdb200 combinehits <querySetDB> <targetSetDB> <aggregate_merged> <matchesDB> <tmpDir>");
    const double alpha = (double)(float)c.num("--alpha", 1.0);
    const bool filterSelf = c.num("--filter-self-match", 0) != 0;
    const auto memberToSet = numberDb(c.pos[1] + "_member_to_set");
    const auto qSize = numberDb(c.pos[0] + "_set_size");
    const auto tSize = numberDb(c.pos[1] + "_set_size");
    const size_t numTargetSets = tSize.size();
    uint32_t maxOrf = 0;
    for (const auto &kv : qSize) maxOrf = std::max(maxOrf, kv.second);
    std::vector<double> lgam(maxOrf + 2);
    for (size_t i = 0; i < lgam.size(); i++) lgam[i] = lgamma((double)i);            // combinehits.cpp:41-44
    std::vector<double> logB(std::max<uint32_t>(maxOrf, 1));
    const sddb::Db in = openDb(c.pos[2]);
    sddb::DbWriter w(c.pos[3], sddb::DBTYPE_ALIGNMENT_RES), wh(c.pos[3] + "_h", sddb::DBTYPE_GENERIC_DB);
    uint32_t matchIdx = 0;
    std::string header, body;
    for (const sddb::Entry &e : byOffset(in)) {
        const uint32_t qSet = e.key;
        auto groups = groupByTargetSet(in.ptr(e), in.size(e), memberToSet);
        auto qs = qSize.find(qSet);
        if (qs == qSize.end()) fail("Invalid query size database key " + std::to_string(qSet) + ".");
        const uint32_t orfCount = qs->second;
        if (orfCount == 0) fail("query set " + std::to_string(qSet) + " has no genes");
        {   // precomputeLogB (lib/mmseqs/src/multihit/combinepvalperset.cpp:17-26) with the threshold alpha / (orfCount + 1)
            const double thr = alpha / (orfCount + 1);
            const double logThr = std::log(thr), log1m = std::log(1 - thr);
            logB[orfCount - 1] = orfCount * logThr;
            for (int i = (int)orfCount - 2; i >= 0; i--) {
                const int k = i + 1;
                const double newTerm = (lgam[orfCount + 1] - lgam[orfCount - k + 1] - lgam[k + 1]) + k * logThr + (orfCount - k) * log1m;
                logB[i] = logB[i + 1] + std::log(1 + std::exp(newTerm - logB[i + 1]));
            }
        }
        for (auto &kv : groups) {
            const uint32_t tSet = kv.first;
            if (filterSelf && qSet == tSet) continue;
            auto ts = tSize.find(tSet);
            if (ts == tSize.end()) fail("Invalid target size database key " + std::to_string(tSet) + ".");
            // multihit mode (combinehits.cpp:98-150): hits below the hard-coded P-value threshold 10e-7
            const double logPvalThr = std::log(10e-7);
            size_t k = 0;
            double r = 0;
            std::vector<const std::vector<std::string> *> entries;
            for (const auto &cols : kv.second) {
                if (cols.size() < 3) fail("invalid aggregate record in entry " + std::to_string(qSet));
                const double logP = strtod(cols[2].c_str(), nullptr);
                if (logP < logPvalThr) { k++; r -= logP - logPvalThr; entries.push_back(&cols); }
            }
            if (r == 0 || k == 0) continue;
            header = std::to_string(qSet) + "\t" + std::to_string(tSet) + "\t" + std::to_string(orfCount) + "\t" + std::to_string(ts->second) + "\t" +
                     std::to_string(k) + "\t";
            const double expMinusR = std::exp(-r);
            if (std::isinf(r) || expMinusR == 0) header += fmt3E(0.0);
            else {
                double fisher = 0;
                const double logR = std::log(r);
                for (size_t i = 0; i < orfCount; i++) fisher += std::exp(i * logR - lgam[i + 1] + logB[i]);
                header += fmt3E(expMinusR * fisher * (double)numTargetSets);
            }
            header += '\n';
            body.clear();
            for (size_t j = 0; j < entries.size(); j++) {
                const std::vector<std::string> &cols = *entries[j];
                for (size_t i = 0; i < cols.size(); i++) {
                    body += (i == 2) ? fmt3E(std::exp(strtod(cols[i].c_str(), nullptr))) : cols[i];
                    if (i + 1 != cols.size()) body += '\t';
                }
                body += '\n';
            }
            // (the reference passes "header,lines" through one string and splits it at the commas: a comma inside a line would cut it)
            const size_t comma = body.find(',');
            if (comma != std::string::npos) { body.resize(comma); body += '\n'; }
            wh.add(matchIdx, header);
            w.add(matchIdx, body);
            matchIdx++;
        }
    }
    std::string err;
    if (!w.close(err) || !wh.close(err)) fail(err);
    return EXIT_SUCCESS;
}

int modClusterhits(int argc, char **argv) {
    const Cli c = parseCli(argc, argv);
    checkFlags(c, {"--alpha", "--max-gene-gap", "--cluster-size", "--cluster-pval", "--multihit-pval", "--filter-self-match", "--gpu"});
    if (c.pos.size() != 4) fail("usage: sdb200 clusterhits <querySetDB> <targetSetDB> <matchesDB> <clustersDB>");
    sddb::SetDb qdb, tdb;
    std::string err;
    if (!qdb.open(c.pos[0], err)) fail(err);
    const bool same = c.pos[0] == c.pos[1];
    if (!same && !tdb.open(c.pos[1], err)) fail(err);
    const sddb::SetDb &tref = same ? qdb : tdb;
    std::unordered_map<uint32_t, uint32_t> qSetOfKey;
    for (uint32_t s = 0; s < qdb.setKey.size(); s++) qSetOfKey[qdb.setKey[s]] = s;
    const sddb::Db matches = openDb(c.pos[2]), hdr = openDb(c.pos[2] + "_h");
    if (matches.index.size() != hdr.index.size()) fail("matches and matches_h differ in their number of entries");
    const std::vector<sddb::Entry> mE = byOffset(matches), hE = byOffset(hdr);

    // one record per match: header columns qSet tSet Nq ...; per hit the whole line (passed through), P-value (column 3),
    // gene position and strand of both genes from the lookup names (ClusterHits.cpp:296-345)
    std::vector<uint64_t> first(1, 0);
    std::vector<uint32_t> mQSet, mTSet, mNq, qPos, tPos;
    std::vector<uint8_t> sameStrand;
    std::vector<double> pval;
    std::vector<Line> hitLine;
    uint32_t maxSet = 0;
    for (uint32_t s : qdb.setSize) maxSet = std::max(maxSet, s);
    for (uint32_t s : tref.setSize) maxSet = std::max(maxSet, s);
    for (size_t i = 0; i < mE.size(); i++) {
        const std::vector<Line> hl = linesOf(hdr.ptr(hE[i]), hdr.size(hE[i]));
        if (hl.empty()) fail("Invalid header record");
        const std::vector<std::string> hc = splitTabs(hl[0]);
        if (hc.size() < 6) fail("Invalid header record");
        mQSet.push_back((uint32_t)strtoul(hc[0].c_str(), nullptr, 10));
        mTSet.push_back((uint32_t)strtoul(hc[1].c_str(), nullptr, 10));
        mNq.push_back((uint32_t)strtoul(hc[2].c_str(), nullptr, 10));
        for (const Line &l : linesOf(matches.ptr(mE[i]), matches.size(mE[i]))) {
            if (l.n == 0) continue;
            const std::vector<std::string> cols = splitTabs(l);
            if (cols.size() < 3) fail("Invalid alignment result record");
            const uint32_t qKey = (uint32_t)strtoul(cols[0].c_str(), nullptr, 10), tKey = (uint32_t)strtoul(cols[1].c_str(), nullptr, 10);
            auto qi = qdb.posOfKey.find(qKey);
            auto ti = tref.posOfKey.find(tKey);
            if (qi == qdb.posOfKey.end()) fail("Invalid query lookup record");
            if (ti == tref.posOfKey.end()) fail("Invalid target lookup record");
            qPos.push_back(qdb.genePos[qi->second]);
            tPos.push_back(tref.genePos[ti->second]);
            sameStrand.push_back((qdb.strand[qi->second] != 0) == (tref.strand[ti->second] != 0) ? 1 : 0);
            pval.push_back(strtod(cols[2].c_str(), nullptr));
            hitLine.push_back(l);
        }
        first.push_back(qPos.size());
    }

    sd_cluster_params cp;
    sd_cluster_params_default(&cp);
    cp.alpha = (float)c.num("--alpha", cp.alpha);
    cp.max_gene_gap = (int)c.num("--max-gene-gap", cp.max_gene_gap);
    cp.cluster_size = (int)c.num("--cluster-size", cp.cluster_size);
    cp.cluster_pval = c.num("--cluster-pval", cp.cluster_pval);
    cp.multihit_pval = c.num("--multihit-pval", cp.multihit_pval);
    sd_context *ctx = nullptr;
    if (sd_context_create((int)c.num("--gpu", -1), &ctx) != SD_OK) fail(sd_last_error());
    std::vector<sd_cluster> clusters(qPos.size() / 2 + 1);
    std::vector<sd_cluster_member> members(qPos.size() + 1);
    uint64_t nC = 0, nM = 0;
    sd_cluster_stats st;
    if (sd_clusterhits_run(ctx, (uint32_t)mQSet.size(), first.data(), mQSet.data(), mTSet.data(), mNq.data(), qPos.data(), tPos.data(),
                           sameStrand.data(), pval.data(), maxSet, &cp, clusters.data(), clusters.size(), &nC, members.data(), members.size(), &nM,
                           &st) != SD_OK)
        fail(sd_last_error());
    sd_context_destroy(ctx);

    // clusters: the member lines as they stand in `matches`; clusters_h: qSet tSet pCO pMH size (ClusterHits.cpp:455-483)
    sddb::DbWriter wc(c.pos[3], sddb::DBTYPE_ALIGNMENT_RES), wh(c.pos[3] + "_h", sddb::DBTYPE_GENERIC_DB);
    std::string body, header;
    for (uint64_t k = 0; k < nC; k++) {
        const sd_cluster &cl = clusters[k];
        header = std::to_string(cl.q_set) + "\t" + std::to_string(cl.t_set) + "\t" + fmt3E(cl.p_co) + "\t" + fmt3E(cl.p_mh) + "\t" + std::to_string(cl.n) + "\n";
        body.clear();
        for (uint32_t m = 0; m < cl.n; m++) {
            const Line &l = hitLine[members[cl.first + m].aln];
            body.append(l.p, l.n);
            body += '\n';
        }
        wc.add((uint32_t)k, body);
        wh.add((uint32_t)k, header);
    }
    if (!wc.close(err) || !wh.close(err)) fail(err);
    fprintf(stderr, "clusterhits: %zu matches, %llu with at least two hits, %llu clusters (kernel %.2f ms)\n", mQSet.size(),
            (unsigned long long)st.pairs_clustered, (unsigned long long)nC, st.kernel_ms);
    return EXIT_SUCCESS;
}
