// sdb200 -- stand-alone front-end of libsdgpu.so on the reference's on-disk contract.
//
// The reference runs the clustersearch hot path as modules of the `spacedust` binary that exchange MMseqs2 flat-file
// DBs (data/clustersearch.sh:112-151, lib/mmseqs/data/workflow/blastp.sh:70,85).  This tool offers the same modules on
// the same files, each one a thin host wrapper around one C-ABI call of include/sdgpu.h:
//   sdb200 prefilter     <qDB> <tDB> <prefDB>            [prefilter flags]   Prefiltering::runSplit   -> sd_prefilter_run
//   sdb200 align         <qDB> <tDB> <prefDB> <resultDB> [align flags]       Alignment::run           -> sd_align_run
//   sdb200 clustersearch <qSetDB> <tSetDB> <out.tsv> <tmp> [all flags]       clustersearch.sh:112-151 -> sd_clustersearch_run
//                        (writes <tmp>/clusters, <tmp>/clusters_h and, with --write-dbs 1, <tmp>/result, result_prefixed, aggregate, aggregate_merged, matches(_h); then the TSV of
//                         summarizeresults, src/util/SummarizeResults.cpp:77-114, <tmp>/clu_to_seq and <out>_seq_to_clu)
//   sdb200 summarizeresults / filterdb --trim-to-one-column / swapdb      host-only text modules (postmods.cpp)
// Flags keep the reference's names and meaning; reference flags that do not change the result of this path are
// accepted and ignored; flags that select a path this implementation does not have make the tool fail
// (Debug(Debug::ERROR) + EXIT(EXIT_FAILURE) in the reference, "Error: ..." + exit 1 here).  No CPU fallback.
#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits.h>
#include <map>
#include <string>
#include <sys/stat.h>
#include <sys/wait.h>
#include <unistd.h>
#include <vector>

#include "../../include/sdgpu.h"
#include "dbio.h"

// postmods.cpp
int filterdbTrimToOneColumn(const std::string &in, const std::string &out);
int swapOneColumnDb(const std::string &in, const std::string &out);

int modCreatesetdb(int argc, char **argv);
int modSummarizeresults(int argc, char **argv);
int modFilterdb(int argc, char **argv);
int modSwapdb(int argc, char **argv);
int modSubtractdbs(int argc, char **argv);
int modMergedbs(int argc, char **argv);
int modPrefixid(int argc, char **argv);
int modBesthitbyset(int argc, char **argv);
int modMergeresultsbyset(int argc, char **argv);
int modCombinehits(int argc, char **argv);
int modClusterhits(int argc, char **argv);

namespace {

int g_verbosity = 3;

[[noreturn]] void die(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    fprintf(stderr, "Error: ");
    vfprintf(stderr, fmt, ap);
    fprintf(stderr, "\n");
    va_end(ap);
    exit(EXIT_FAILURE);
}

void info(const char *fmt, ...) {
    if (g_verbosity < 3) return;
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stdout, fmt, ap);
    va_end(ap);
    fflush(stdout);
}

double nowMs() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

void sdCheck(int rc, const char *what) {
    if (rc != SD_OK) die("%s failed (%d): %s", what, rc, sd_last_error());
}

struct Args {
    std::vector<std::string> pos;
    std::map<std::string, std::string> flags;
    bool has(const char *f) const { return flags.count(f) != 0; }
    std::string str(const char *f, const char *d) const { auto it = flags.find(f); return it == flags.end() ? std::string(d) : it->second; }
    double num(const char *f, double d) const {
        auto it = flags.find(f);
        if (it == flags.end()) return d;
        // "aa:11,nucl:5" style values: take the aa part
        std::string v = it->second;
        size_t a = v.find("aa:");
        if (a != std::string::npos) v = v.substr(a + 3);
        char *end = nullptr;
        double x = strtod(v.c_str(), &end);
        if (end == v.c_str()) die("flag %s: cannot parse value '%s'", f, it->second.c_str());
        return x;
    }
};

Args parseArgs(int argc, char **argv, int first) {
    Args a;
    for (int i = first; i < argc; i++) {
        const char *s = argv[i];
        if (s[0] == '-' && s[1] != '\0' && !(s[1] >= '0' && s[1] <= '9')) {
            if (i + 1 >= argc) die("flag %s needs a value", s);
            a.flags[s] = argv[++i];       // every flag of these modules takes one value
        } else {
            a.pos.push_back(s);
        }
    }
    return a;
}

// Every flag is classified: implemented (read somewhere below), inert for this path (host-runtime knobs and parameters of code
// that `--search-mode 0` sequence search never reaches: nucleotide / ORF extraction, profile building, GPU server of the
// reference, ...), or accepted only with the value the clustersearch workflow passes (data/clustersearch.sh, blastp.sh) because any
// other value selects code that is not built here.  Anything else stops the program: a silently ignored flag would return a
// different result with exit code 0.
void checkKnownFlags(const Args &a) {
    static const char *implemented[] = {
        "-s", "-c", "-e", "-k", "-v", "-a", "--max-seqs", "--min-ungapped-score", "--cov-mode", "--mask", "--mask-prob", "--comp-bias-corr",
        "--min-seq-id", "--min-aln-len", "--score-bias", "--gap-open", "--gap-extend", "--alignment-mode", "--filter-self-match",
        "--max-gene-gap", "--cluster-size", "--cluster-pval", "--multihit-pval", "--alpha", "--gpu", "--gpus", "--rank", "--nccl-id-file", "--write-dbs", "--ignore-index", "--search-mode",
        "--num-iterations", "--exact-kmer-matching", "--diag-score", "--spaced-kmer-mode", "--realign", "--alt-ali", "--compressed",
        "--split", "--add-self-matches", "--cluster-use-weight", "--aggregation-mode", "--sub-mat", "--seed-sub-mat", nullptr};
    static const char *inert[] = {
        "--threads", "--remove-tmp-files", "--force-reuse", "--db-load-mode", "--split-mode", "--split-memory-limit", "--disk-space-limit",
        "--local-tmp", "--mpi-runner", "--gpu-server", "--gpu-server-wait-timeout", "--zdrop", "--pca", "--pcb", "--mask-profile", "--e-profile",
        "--wg", "--allow-deletion", "--filter-msa", "--filter-min-enable", "--max-seq-id", "--qid", "--qsc", "--cov", "--diff", "--pseudo-cnt-mode",
        "--profile-output-mode", "--min-length", "--max-length", "--max-gaps", "--contig-start-mode", "--contig-end-mode", "--orf-start-mode",
        "--forward-frames", "--reverse-frames", "--translation-table", "--translate", "--use-all-table-starts", "--id-offset", "--create-lookup",
        "--sequence-overlap", "--sequence-split-mode", "--headers-split-mode", "--translation-mode", "--strand", "--start-sens", "--sens-steps",
        "--realign-score-bias", "--realign-max-seqs", "--corr-score-weight", "--chain-alignments", "--merge-query", "--db-output",
        "--tar-include", "--tar-exclude", "--file-include", "--file-exclude", "--gff-type", "--dbtype", "--shuffle", "--write-lookup",
        "--createdb-mode", "--stat", "--tsv", "--output-mode", "--mask-n-repeat", nullptr};
    struct Only { const char *flag; const char *values[3]; };
    static const Only only[] = {
        {"--suboptimal-hits", {"0"}}, {"--simple-best-hit", {"1"}}, {"--profile-cluster-search", {"0"}},
        {"--max-accept", {"2147483647"}}, {"--max-rejected", {"2147483647"}}, {"--seq-id-mode", {"0"}},
        {"--alignment-output-mode", {"0"}}, {"--wrapped-scoring", {"0"}}, {"--exhaustive-search", {"0"}},
        {"--exhaustive-search-filter", {"0"}}, {"--target-search-mode", {"0"}}, {"--k-score", {"seq:2147483647,prof:2147483647", "2147483647"}},
        {"--alph-size", {"aa:21,nucl:5", "21"}}, {"--comp-bias-corr-scale", {"1", "1.0", "1.000"}}, {"--mask-lower-case", {"0"}},
        {"--max-seq-len", {"65535"}}, {"--prefilter-mode", {"0"}}, {"--rescore-mode", {"0"}}, {"--filter-hits", {"0"}},
        {"--sort-results", {"0"}}, {"--search-type", {"0"}}, {"--lca-search", {"0"}}, {"--taxon-list", {""}},
    };
    for (const auto &kv : a.flags) {
        bool ok = false;
        for (int i = 0; implemented[i] && !ok; i++) ok = kv.first == implemented[i];
        for (int i = 0; inert[i] && !ok; i++) ok = kv.first == inert[i];
        if (ok) continue;
        bool known = false;
        for (const Only &o : only) {
            if (kv.first != o.flag) continue;
            known = true;
            for (const char *v : o.values) ok = ok || (v && kv.second == v);
            if (!ok) die("%s %s is not implemented (only %s %s, the value the clustersearch workflow passes)", o.flag, kv.second.c_str(), o.flag, o.values[0]);
        }
        if (!known) die("unknown or unsupported flag %s (it would be ignored; refusing to run)", kv.first.c_str());
    }
}

// Flags whose non-default values select code this path does not implement.
void rejectUnsupported(const Args &a) {
    checkKnownFlags(a);
    if (a.num("--search-mode", 0) != 0) die("--search-mode %s is not implemented (only 0: MMseqs2 protein search)", a.str("--search-mode", "").c_str());
    if (a.num("--num-iterations", 1) != 1) die("--num-iterations > 1 (profile search) is not implemented");
    if (a.num("-k", 0) != 0 && a.num("-k", 0) != 6) die("-k %s is not implemented (k = 6, as the reference selects for these DB sizes)", a.str("-k", "").c_str());
    if (a.num("--alignment-mode", 2) != 2 && a.has("--alignment-mode") && a.num("--alignment-mode", 2) != 0 && a.num("--alignment-mode", 2) != 3)
        die("--alignment-mode %s is not implemented (2/3: score, coordinates, backtrace)", a.str("--alignment-mode", "").c_str());
    if (a.num("--exact-kmer-matching", 0) != 0) die("--exact-kmer-matching 1 is not implemented");
    if (a.num("--diag-score", 1) != 1) die("--diag-score 0 is not implemented");
    if (a.num("--spaced-kmer-mode", 1) != 1) die("--spaced-kmer-mode 0 is not implemented");
    if (a.num("--realign", 0) != 0) die("--realign 1 is not implemented");
    if (a.num("--alt-ali", 0) != 0) die("--alt-ali > 0 is not implemented");
    sddb::setCompressedOutput(a.num("--compressed", 0) != 0);      // result DBs as zstd frames, like DBWriter in compressed mode
    if (a.num("--split", 0) > 1) die("--split > 1 is not implemented (the whole target index fits one B200)");
    if (a.num("--add-self-matches", 0) != 0) die("--add-self-matches 1 is not implemented");
    if (a.num("--cluster-use-weight", 0) != 0) die("--cluster-use-weight 1 is not implemented");
    if (a.num("--aggregation-mode", 0) != 0) die("--aggregation-mode != 0 is not implemented");
    if (a.has("--sub-mat") && a.str("--sub-mat", "").find("blosum62") == std::string::npos) die("--sub-mat: only blosum62.out is built in");
    if (a.has("--seed-sub-mat") && a.str("--seed-sub-mat", "").find("VTML80") == std::string::npos) die("--seed-sub-mat: only VTML80.out is built in");
}

std::string realPath(const std::string &p) {
    char buf[PATH_MAX];
    if (realpath(p.c_str(), buf)) return buf;
    return p;
}

// workflow = true: defaults of `clustersearch` (data/clustersearch.sh passes -s 5.7 -c 0.8 --cov-mode 2 -e 10 --min-aln-len 30);
// workflow = false: the stand-alone `prefilter` / `align` modules start from the reference's module defaults
// (Parameters.cpp:2364-2425: -s 4, -c 0, --cov-mode 0, -e 0.001, --min-aln-len 0), so that the same command line gives the same DB.
sd_prefilter_params prefParams(const Args &a, bool sameDb, bool workflow = true) {
    sd_prefilter_params p;
    sd_prefilter_params_default(&p);
    if (!workflow) { p.sensitivity = 4.0f; p.cov_thr = 0.0f; p.cov_mode = 0; }
    p.sensitivity = (float)a.num("-s", p.sensitivity);
    p.max_seqs = (int)a.num("--max-seqs", p.max_seqs);
    p.min_ungapped_score = (int)a.num("--min-ungapped-score", p.min_ungapped_score);
    p.cov_thr = (float)a.num("-c", p.cov_thr);
    p.cov_mode = (int)a.num("--cov-mode", p.cov_mode);
    p.mask = (int)a.num("--mask", p.mask);
    p.mask_prob = (float)a.num("--mask-prob", p.mask_prob);
    p.comp_bias_corr = (int)a.num("--comp-bias-corr", p.comp_bias_corr);
    p.same_db = sameDb ? 1 : 0;
    return p;
}

sd_align_params alignParams(const Args &a, bool sameDb, bool workflow = true) {
    sd_align_params p;
    sd_align_params_default(&p);
    if (!workflow) { p.evalue_thr = 0.001; p.cov_thr = 0.0f; p.cov_mode = 0; p.min_aln_len = 0; }
    p.evalue_thr = a.num("-e", p.evalue_thr);
    p.min_seq_id = a.num("--min-seq-id", p.min_seq_id);
    p.cov_thr = (float)a.num("-c", p.cov_thr);
    p.cov_mode = (int)a.num("--cov-mode", p.cov_mode);
    p.min_aln_len = (int)a.num("--min-aln-len", p.min_aln_len);
    p.comp_bias_corr = (int)a.num("--comp-bias-corr", p.comp_bias_corr);
    p.gap_open = (int)a.num("--gap-open", p.gap_open);
    p.gap_extend = (int)a.num("--gap-extend", p.gap_extend);
    p.score_bias = (float)a.num("--score-bias", p.score_bias);
    p.same_db = sameDb ? 1 : 0;
    return p;
}

sd_cluster_params clusterParams(const Args &a) {
    sd_cluster_params p;
    sd_cluster_params_default(&p);
    p.alpha = (float)a.num("--alpha", p.alpha);
    p.filter_self_match = (int)a.num("--filter-self-match", p.filter_self_match);
    p.max_gene_gap = (int)a.num("--max-gene-gap", p.max_gene_gap);
    p.cluster_size = (int)a.num("--cluster-size", p.cluster_size);
    p.cluster_pval = a.num("--cluster-pval", p.cluster_pval);
    p.multihit_pval = a.num("--multihit-pval", p.multihit_pval);
    return p;
}

// Sequence DBs for the stand-alone prefilter/align modules need no lookup; fall back to a bare reader.
struct SeqView {
    std::vector<uint32_t> keys;
    std::unordered_map<uint32_t, uint32_t> posOfKey;
    std::string residues;
    std::vector<uint64_t> off;
    bool keysAscend = true;
    void open(const std::string &path) {
        sddb::Db db;
        std::string err;
        if (!db.open(path, err)) die("%s", err.c_str());
        std::vector<sddb::Entry> ord = db.index;
        std::stable_sort(ord.begin(), ord.end(), [](const sddb::Entry &x, const sddb::Entry &y) { return x.off < y.off; });
        off.assign(ord.size() + 1, 0);
        keys.resize(ord.size());
        uint64_t total = 0;
        for (auto &e : ord) {
            if (e.len < 2) die("%s: sequence entry shorter than its terminator", path.c_str());
            total += e.len - 2;
        }
        residues.resize(total);
        for (size_t i = 0; i < ord.size(); i++) {
            keys[i] = ord[i].key;
            if (i && keys[i] <= keys[i - 1]) keysAscend = false;
            posOfKey[ord[i].key] = (uint32_t)i;
            memcpy(&residues[off[i]], db.data.data() + ord[i].off, ord[i].len - 2);
            off[i + 1] = off[i] + ord[i].len - 2;
        }
    }
};

inline char *putU32(char *p, uint32_t v) {
    char tmp[12];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *p++ = tmp[--n];
    return p;
}

inline char *putI32(char *p, int32_t v) {
    if (v < 0) { *p++ = '-'; return putU32(p, (uint32_t)(-(int64_t)v)); }
    return putU32(p, (uint32_t)v);
}

// ---------------------------------------------------------------------------------------------------------- prefilter
int modPrefilter(int argc, char **argv) {
    Args a = parseArgs(argc, argv, 2);
    if (a.pos.size() != 3) die("usage: sdb200 prefilter <queryDB> <targetDB> <prefDB> [flags]");
    g_verbosity = (int)a.num("-v", 3);
    rejectUnsupported(a);
    const bool same = realPath(a.pos[0]) == realPath(a.pos[1]);
    SeqView t, qv;
    t.open(a.pos[1]);
    if (!same) qv.open(a.pos[0]);
    SeqView &q = same ? t : qv;
    sd_context *ctx = nullptr;
    sdCheck(sd_context_create(-1, &ctx), "sd_context_create");
    sd_prefilter_params p = prefParams(a, same, false);
    info("prefilter: -s %g --max-seqs %d -c %g --cov-mode %d --min-ungapped-score %d (reference module defaults unless given)\n", p.sensitivity,
         p.max_seqs, p.cov_thr, p.cov_mode, p.min_ungapped_score);
    sd_target_index *idx = nullptr;
    double t0 = nowMs();
    sdCheck(sd_target_index_build(ctx, t.residues.data(), t.off.data(), (uint32_t)t.keys.size(), &p, &idx), "sd_target_index_build");
    info("Index table: %llu entries, %.0f ms\n", (unsigned long long)sd_target_index_entries(idx), nowMs() - t0);
    const uint64_t cap = (uint64_t)q.keys.size() * (uint64_t)std::max(1, std::min<int>(p.max_seqs, (int)t.keys.size())) + 1;
    std::vector<sd_prefilter_hit> hits(cap);
    uint64_t n = 0;
    sd_prefilter_stats st;
    t0 = nowMs();
    sdCheck(sd_prefilter_run(ctx, idx, q.residues.data(), q.off.data(), (uint32_t)q.keys.size(), 0, &p, hits.data(), cap, &n, &st), "sd_prefilter_run");
    info("prefilter: %llu hits for %zu queries, kernel %.1f ms, call %.1f ms\n", (unsigned long long)n, q.keys.size(), st.kernel_ms, nowMs() - t0);
    // QueryMatcher::prefilterHitToBuffer (QueryMatcher.h:118-130): "tKey\tscore\tdiag\n", diag printed as signed short
    sddb::DbWriter w(a.pos[2], sddb::DBTYPE_PREFILTER_RES);
    std::string buf;
    uint64_t i = 0;
    for (uint32_t qi = 0; qi < q.keys.size(); qi++) {
        buf.clear();
        char line[48];
        while (i < n && hits[i].query == qi) {
            char *e = putU32(line, t.keys[hits[i].target]);
            *e++ = '\t';
            e = putI32(e, hits[i].score);
            *e++ = '\t';
            e = putI32(e, hits[i].diag);
            *e++ = '\n';
            buf.append(line, (size_t)(e - line));
            i++;
        }
        w.add(q.keys[qi], buf);
    }
    std::string err;
    if (!w.close(err)) die("%s", err.c_str());
    sd_target_index_destroy(idx);
    sd_context_destroy(ctx);
    return EXIT_SUCCESS;
}

// -------------------------------------------------------------------------------------------------------------- align
int modAlign(int argc, char **argv) {
    Args a = parseArgs(argc, argv, 2);
    if (a.pos.size() != 4) die("usage: sdb200 align <queryDB> <targetDB> <prefDB> <resultDB> [flags]");
    g_verbosity = (int)a.num("-v", 3);
    rejectUnsupported(a);
    const bool same = realPath(a.pos[0]) == realPath(a.pos[1]);
    SeqView t, qv;
    t.open(a.pos[1]);
    if (!same) qv.open(a.pos[0]);
    SeqView &q = same ? t : qv;
    if (!t.keysAscend)
        die("%s: the DB keys do not ascend with the data offset; Matcher::compareHits breaks ties of equal (E-value, score, target length) by "
            "the target KEY, this implementation by the target's position -- rewrite the DB with createdb / createsetdb", a.pos[1].c_str());
    sddb::Db pref;
    std::string err;
    if (!pref.open(a.pos[2], err)) die("%s", err.c_str());
    // pairs grouped by query in ascending query position
    std::vector<std::vector<uint32_t>> perQ(q.keys.size());
    for (const sddb::Entry &e : pref.index) {
        auto it = q.posOfKey.find(e.key);
        if (it == q.posOfKey.end()) die("prefilter DB key %u is not in the query DB", e.key);
        const char *p = pref.ptr(e), *end = p + pref.size(e);
        while (p < end) {
            uint32_t tk = (uint32_t)strtoul(p, nullptr, 10);
            auto jt = t.posOfKey.find(tk);
            if (jt == t.posOfKey.end()) die("prefilter DB target key %u is not in the target DB", tk);
            perQ[it->second].push_back(jt->second);
            while (p < end && *p != '\n') p++;
            p++;
        }
    }
    std::vector<uint32_t> pq, pt;
    for (uint32_t qi = 0; qi < perQ.size(); qi++)
        for (uint32_t tj : perQ[qi]) { pq.push_back(qi); pt.push_back(tj); }
    sd_context *ctx = nullptr;
    sdCheck(sd_context_create(-1, &ctx), "sd_context_create");
    sd_prefilter_params pp = prefParams(a, same);
    pp.mask = 0;       // the alignment reads unmasked residues; no k-mer index is probed here
    sd_align_params ap = alignParams(a, same, false);
    info("align: -e %g -c %g --cov-mode %d --min-aln-len %d --min-seq-id %g (reference module defaults unless given)\n", ap.evalue_thr, ap.cov_thr,
         ap.cov_mode, ap.min_aln_len, ap.min_seq_id);
    sd_target_index *idx = nullptr;
    sdCheck(sd_target_index_build(ctx, t.residues.data(), t.off.data(), (uint32_t)t.keys.size(), &pp, &idx), "sd_target_index_build");
    uint64_t cap = std::max<uint64_t>(1024, pq.size() / 4), btCap = cap * 512, n = 0, used = 0;
    std::vector<sd_alignment> alns;
    std::vector<char> pool;
    sd_align_stats st;
    for (int attempt = 0;; attempt++) {
        alns.resize(cap);
        pool.resize(btCap);
        int rc = sd_align_run(ctx, idx, q.residues.data(), q.off.data(), (uint32_t)q.keys.size(), 0, pq.data(), pt.data(), pq.size(), 0, &ap,
                              alns.data(), cap, &n, pool.data(), btCap, &used, &st);
        if (rc == SD_ESMALL && attempt == 0) { cap = n + 16; btCap = used + 64; continue; }
        sdCheck(rc, "sd_align_run");
        break;
    }
    info("align: %llu pairs, %llu accepted, SW kernels %.1f ms, traceback %.1f ms\n", (unsigned long long)st.alignments,
         (unsigned long long)st.passed, st.sw_ms, st.traceback_ms);
    sddb::DbWriter w(a.pos[3], sddb::DBTYPE_ALIGNMENT_RES);
    std::string buf;
    std::vector<char> line(1 << 20);
    uint64_t i = 0;
    for (uint32_t qi = 0; qi < q.keys.size(); qi++) {
        buf.clear();
        while (i < n && alns[i].query == qi) {
            int len = sd_format_alignment(&alns[i], t.keys[alns[i].target], pool.data(), line.data(), (int)line.size());
            if (len <= 0) die("alignment line does not fit the format buffer");
            buf.append(line.data(), (size_t)len);
            i++;
        }
        w.add(q.keys[qi], buf);
    }
    if (!w.close(err)) die("%s", err.c_str());
    sd_target_index_destroy(idx);
    sd_context_destroy(ctx);
    return EXIT_SUCCESS;
}

// ------------------------------------------------------------------------------------------------------ clustersearch
// `createindex <setDB> <tmpDir>`: build the target index once and keep it next to the DB (<setDB>.sdidx) -- the reference's
// createindex writes <db>.idx, which later searches pick up instead of rebuilding the table (PrefilteringIndexReader).
// `search <queryDB> <targetDB> <resultDB> <tmpDir> [flags]`: the sequence search of lib/mmseqs/data/workflow/blastp.sh:70-95 (one
// sensitivity step, --num-iterations 1): prefilter into <tmpDir>/pref_0, then align.  With it, and the aggregation modules, the
// reference's own data/clustersearch.sh runs with MMSEQS pointing at this binary.
int modSearch(int argc, char **argv) {
    Args a = parseArgs(argc, argv, 2);
    if (a.pos.size() != 4) die("usage: sdb200 search <queryDB> <targetDB> <resultDB> <tmpDir> [flags]");
    if (a.num("--start-sens", 4) != 4 && a.has("--sens-steps") && a.num("--sens-steps", 1) > 1) die("--sens-steps > 1 is not implemented");
    for (size_t i = 1; i <= a.pos[3].size(); i++)      // mkdir -p: the search workflow creates its tmp directory itself
        if (i == a.pos[3].size() || a.pos[3][i] == '/') mkdir(a.pos[3].substr(0, i).c_str(), 0777);
    const std::string pref = a.pos[3] + "/pref_0";
    std::vector<std::string> flagTokens;          // every flag with its value, in the order given
    {
        size_t seenPos = 0;
        for (int i = 2; i < argc; i++) {
            if (argv[i][0] == '-' && argv[i][1] != '\0' && !(argv[i][1] >= '0' && argv[i][1] <= '9')) {
                flagTokens.push_back(argv[i]);
                if (i + 1 < argc) flagTokens.push_back(argv[++i]);
            } else seenPos++;
        }
        (void)seenPos;
    }
    auto run = [&](int (*fn)(int, char **), std::vector<std::string> args) {
        args.insert(args.end(), flagTokens.begin(), flagTokens.end());
        std::vector<char *> av;
        for (std::string &x : args) av.push_back(&x[0]);
        return fn((int)av.size(), av.data());
    };
    int rc = run(modPrefilter, {argv[0], "prefilter", a.pos[0], a.pos[1], pref});
    if (rc != EXIT_SUCCESS) return rc;
    return run(modAlign, {argv[0], "align", a.pos[0], a.pos[1], pref, a.pos[2]});
}

// `rmdb <db>`: remove a DB and its side files (lib/mmseqs/src/util/rmdb.cpp -> DBReader::removeDb)
int modRmdb(int argc, char **argv) {
    Args a = parseArgs(argc, argv, 2);
    if (a.pos.size() != 1) die("usage: sdb200 rmdb <DB>");
    const std::string &db = a.pos[0];
    for (const char *suffix : {"", ".index", ".dbtype", ".lookup", ".source", "_h", "_h.index", "_h.dbtype"}) remove((db + suffix).c_str());
    for (int i = 0;; i++) {
        const std::string part = db + "." + std::to_string(i);
        struct stat stPart;
        if (stat(part.c_str(), &stPart) != 0) break;
        remove(part.c_str());
    }
    return EXIT_SUCCESS;
}

int modCreateindex(int argc, char **argv) {
    Args a = parseArgs(argc, argv, 2);
    if (a.pos.size() < 1 || a.pos.size() > 2) die("usage: sdb200 createindex <setDB> [tmpDir] [-s 5.7] [--mask 1] [--mask-prob 0.9]");
    g_verbosity = (int)a.num("-v", 3);
    rejectUnsupported(a);
    sddb::SetDb tdb;
    std::string err;
    if (!tdb.open(a.pos[0], err)) die("%s", err.c_str());
    sd_context *ctx = nullptr;
    sdCheck(sd_context_create((int)a.num("--gpu", -1), &ctx), "sd_context_create");
    sd_prefilter_params pp = prefParams(a, true);
    sd_target_index *idx = nullptr;
    double t0 = nowMs();
    sdCheck(sd_target_index_build(ctx, tdb.residues.data(), tdb.off.data(), tdb.n(), &pp, &idx), "sd_target_index_build");
    const std::string idxFile = a.pos[0] + ".sdidx";
    sdCheck(sd_target_index_save(idx, &pp, idxFile.c_str()), "sd_target_index_save");
    info("Index table: %llu entries for %u targets (-s %g), %.0f ms -> %s\n", (unsigned long long)sd_target_index_entries(idx), tdb.n(), pp.sensitivity,
         nowMs() - t0, idxFile.c_str());
    sd_target_index_destroy(idx);
    sd_context_destroy(ctx);
    return EXIT_SUCCESS;
}

int modClustersearch(int argc, char **argv) {
    Args a = parseArgs(argc, argv, 2);
    if (a.pos.size() != 4) die("usage: sdb200 clustersearch <querySetDB> <targetSetDB> <output.tsv> <tmpDir> [flags]");
    g_verbosity = (int)a.num("-v", 3);
    rejectUnsupported(a);
    const bool same = realPath(a.pos[0]) == realPath(a.pos[1]);
    const std::string out = a.pos[2], tmp = a.pos[3];
    mkdir(tmp.c_str(), 0777);
    double t0 = nowMs(), tAll = t0;
    sddb::SetDb tdb, qdbOwn;
    std::string err;
    if (!tdb.open(a.pos[1], err)) die("%s", err.c_str());
    if (!same && !qdbOwn.open(a.pos[0], err)) die("%s", err.c_str());
    sddb::SetDb &qdb = same ? tdb : qdbOwn;
    if (!tdb.keysAscend)
        die("%s: the DB keys do not ascend with the data offset; Matcher::compareHits breaks ties of equal (E-value, score, target length) by "
            "the target KEY, this implementation by the target's position -- rewrite the DB with createdb / createsetdb", a.pos[1].c_str());
    info("Read set DBs: %u query genes in %zu sets, %u target genes in %zu sets, %.0f ms\n", qdb.n(), qdb.setKey.size(), tdb.n(),
         tdb.setKey.size(), nowMs() - t0);

    // ---- multi-GPU (--gpus N): one process per GPU.  The parent starts N copies of itself (--rank r); rank 0 creates the NCCL id and
    // hands it over through a file in the tmp dir; every rank searches its block of query genes against the whole target index, the
    // ncclAllGather exchanges inside libsdgpu.so (sd_shard_clustersearch_run) give every rank the whole result; rank 0 writes it.
    const int nGpus = std::max(1, (int)a.num("--gpus", 1));
    const int rank = (int)a.num("--rank", -1);
    if (nGpus > 1 && rank < 0) {
        const std::string idFile = tmp + "/nccl_id." + std::to_string((long)getpid());
        unlink(idFile.c_str());
        std::vector<pid_t> kids;
        for (int r = 0; r < nGpus; r++) {
            pid_t pid = fork();
            if (pid < 0) die("fork failed");
            if (pid == 0) {
                std::vector<std::string> args(argv, argv + argc);
                args.push_back("--rank"); args.push_back(std::to_string(r));
                args.push_back("--nccl-id-file"); args.push_back(idFile);
                std::vector<char *> av;
                for (std::string &x : args) av.push_back(&x[0]);
                av.push_back(nullptr);
                execv("/proc/self/exe", av.data());
                _exit(127);
            }
            kids.push_back(pid);
        }
        int bad = 0;
        for (pid_t pid : kids) {
            int status = 0;
            waitpid(pid, &status, 0);
            if (!WIFEXITED(status) || WEXITSTATUS(status) != 0) bad++;
        }
        unlink(idFile.c_str());
        if (bad) die("%d of %d ranks failed", bad, nGpus);
        return EXIT_SUCCESS;
    }
    const int world = rank >= 0 ? nGpus : 1;
    const int myRank = rank >= 0 ? rank : 0;
    if (myRank != 0) g_verbosity = std::min(g_verbosity, 1);

    sd_context *ctx = nullptr;
    sdCheck(sd_context_create(world > 1 ? (int)std::max(0.0, a.num("--gpu", 0)) + myRank : (int)a.num("--gpu", -1), &ctx), "sd_context_create");
    uint32_t qLo = 0, qHi = qdb.n();
    if (world > 1) {
        const std::string idFile = a.str("--nccl-id-file", "");
        uint8_t id[SD_NCCL_ID_BYTES];
        if (myRank == 0) {
            sdCheck(sd_shard_unique_id(id), "sd_shard_unique_id");
            const std::string part = idFile + ".part";
            FILE *f = fopen(part.c_str(), "wb");
            if (!f || fwrite(id, 1, sizeof id, f) != sizeof id) die("cannot write %s", part.c_str());
            fclose(f);
            if (rename(part.c_str(), idFile.c_str()) != 0) die("cannot publish %s", idFile.c_str());
        } else {
            const double tWait = nowMs();
            for (;;) {
                FILE *f = fopen(idFile.c_str(), "rb");
                if (f) {
                    const size_t got = fread(id, 1, sizeof id, f);
                    fclose(f);
                    if (got == sizeof id) break;
                }
                if (nowMs() - tWait > 300000.0) die("rank %d: no NCCL id from rank 0 after 300 s", myRank);
                usleep(20000);
            }
        }
        sdCheck(sd_shard_init(ctx, myRank, world, id), "sd_shard_init");
        // this rank's block of query genes (prefilter + align); the set-wise stages are re-partitioned inside the library after the
        // first exchange (sd_shard_clustersearch_run)
        for (uint32_t g = 1; g < qdb.n(); g++)
            if (qdb.setOf[g] < qdb.setOf[g - 1]) die("query sets are not contiguous ascending blocks in the set DB");
        std::vector<uint32_t> first((size_t)world + 1);
        sdCheck(sd_shard_partition_genes(qdb.off.data(), qdb.n(), world, first.data()), "sd_shard_partition_genes");
        qLo = first[myRank]; qHi = first[myRank + 1];
        info("rank %d of %d: query genes %u..%u\n", myRank, world, qLo, qHi);
    }
    sd_prefilter_params pp = prefParams(a, same);
    sd_align_params ap = alignParams(a, same);
    sd_cluster_params cp = clusterParams(a);
    t0 = nowMs();
    sd_target_index *idx = nullptr;
    {   // a persisted index next to the target DB (`sdb200 createindex`, the counterpart of the reference's <db>.idx) is used when it
        // was built with the same sensitivity / masking; otherwise the index is built now
        const std::string idxFile = a.pos[1] + ".sdidx";
        struct stat stIdx;
        if ((int)a.num("--ignore-index", 0) == 0 && stat(idxFile.c_str(), &stIdx) == 0) {
            if (sd_target_index_load(ctx, idxFile.c_str(), &pp, &idx) == SD_OK && sd_target_index_targets(idx) == tdb.n())
                info("Loaded the persisted index %s\n", idxFile.c_str());
            else {
                info("Not using %s: %s\n", idxFile.c_str(), idx ? "it indexes another DB" : sd_last_error());
                if (idx) { sd_target_index_destroy(idx); idx = nullptr; }
            }
        }
    }
    if (!idx) sdCheck(sd_target_index_build(ctx, tdb.residues.data(), tdb.off.data(), tdb.n(), &pp, &idx), "sd_target_index_build");
    sd_query_batch *batch = nullptr;
    std::vector<uint64_t> qOffBlock(qdb.off.begin() + qLo, qdb.off.begin() + qHi + 1);
    for (uint64_t &o : qOffBlock) o -= qdb.off[qLo];
    sdCheck(sd_query_batch_create(ctx, qdb.residues.data() + qdb.off[qLo], qOffBlock.data(), qHi - qLo, &batch), "sd_query_batch_create");
    info("Index table: %llu entries, %.0f ms\n", (unsigned long long)sd_target_index_entries(idx), nowMs() - t0);
    t0 = nowMs();
    sd_search_result *res = nullptr;
    sd_prefilter_stats pst;
    sd_align_stats ast;
    sd_cluster_stats cst;
    memset(&pst, 0, sizeof pst);
    // same DB: the identity hit of query gene g is target g = qLo + index in the block
    if (world > 1) {
        uint64_t gathered = 0;
        sdCheck(sd_shard_clustersearch_run(ctx, idx, batch, qLo, &pp, &ap, &cp, qdb.setOf.data(), qdb.genePos.data(), qdb.strand.data(), qdb.n(),
                                           tdb.setOf.data(), tdb.genePos.data(), tdb.strand.data(), tdb.n(), qdb.setSize.data(),
                                           (uint32_t)qdb.setSize.size(), tdb.setSize.data(), (uint32_t)tdb.setSize.size(), 0, &res, &pst, &ast, &cst,
                                           &gathered),
                "sd_shard_clustersearch_run");
        info("all-gathers of the alignment and cluster lists: %.1f MB over NCCL\n", gathered / 1e6);
        if (myRank != 0) {      // every rank holds the whole result now; rank 0 writes it
            sd_search_result_destroy(res);
            sd_query_batch_destroy(batch);
            sd_target_index_destroy(idx);
            sd_context_destroy(ctx);
            return EXIT_SUCCESS;
        }
    } else {
        sdCheck(sd_clustersearch_run(ctx, idx, batch, qLo, &pp, &ap, &cp, qdb.setOf.data() + qLo, qdb.genePos.data() + qLo, qdb.strand.data() + qLo,
                                     tdb.setOf.data(), tdb.genePos.data(), tdb.strand.data(), tdb.n(), qdb.setSize.data(),
                                     (uint32_t)qdb.setSize.size(), tdb.setSize.data(), (uint32_t)tdb.setSize.size(), 0, &res, &pst, &ast, &cst),
                "sd_clustersearch_run");
    }
    uint64_t nA = 0, nPool = 0, nC = 0, nM = 0;
    sd_search_result_sizes(res, &nA, &nPool, &nC, &nM);
    std::vector<sd_alignment> alns(nA + 1);
    std::vector<char> pool(nPool + 1);
    std::vector<sd_cluster> clusters(nC + 1);
    std::vector<sd_cluster_member> members(nM + 1);
    sdCheck(sd_search_result_copy(res, alns.data(), pool.data(), clusters.data(), members.data()), "sd_search_result_copy");
    sd_search_result_destroy(res);
    info("prefilter %.0f ms (kernel %.0f), align %.0f ms (SW %.0f, traceback %.0f), cluster %.0f ms: %llu alignments, %llu clusters, "
         "search %.0f ms\n", pst.total_ms, pst.kernel_ms, ast.total_ms, ast.sw_ms, ast.traceback_ms, cst.total_ms,
         (unsigned long long)nA, (unsigned long long)nC, nowMs() - t0);

    t0 = nowMs();
    std::vector<char> line(1 << 20);
    if ((int)a.num("--write-dbs", 0) != 0) {       // the `result` DB of the search step (Matcher::resultToBuffer lines)
        sddb::DbWriter w(tmp + "/result", sddb::DBTYPE_ALIGNMENT_RES);
        std::string buf;
        uint64_t i = 0;
        for (uint32_t qi = 0; qi < qdb.n(); qi++) {
            buf.clear();
            while (i < nA && alns[i].query == qi) {
                int len = sd_format_alignment(&alns[i], tdb.keys[alns[i].target], pool.data(), line.data(), (int)line.size());
                if (len <= 0) die("alignment line does not fit the format buffer");
                buf.append(line.data(), (size_t)len);
                i++;
            }
            w.add(qdb.keys[qi], buf);
        }
        if (!w.close(err)) die("%s", err.c_str());
        // ... and the intermediates of data/clustersearch.sh:119-141 under the reference's names, written by the stand-alone text
        // modules (csrc/aggmods.cpp) from that result DB, so that the reference (or sdb200 clusterhits) can resume from any of them
        auto runMod = [&](int (*fn)(int, char **), std::vector<std::string> args) {
            std::vector<char *> av;
            args.insert(args.begin(), {"sdb200", "module"});
            for (std::string &x : args) av.push_back(&x[0]);
            fn((int)av.size(), av.data());
        };
        runMod(modPrefixid, {tmp + "/result", tmp + "/result_prefixed"});
        runMod(modBesthitbyset, {a.pos[0], a.pos[1], tmp + "/result_prefixed", tmp + "/aggregate"});
        runMod(modMergeresultsbyset, {a.pos[0] + "_set_to_member", tmp + "/aggregate", tmp + "/aggregate_merged"});
        runMod(modCombinehits, {a.pos[0], a.pos[1], tmp + "/aggregate_merged", tmp + "/matches", tmp, "--alpha", a.str("--alpha", "1"),
                                "--filter-self-match", a.str("--filter-self-match", "0")});
    }
    // clusters / clusters_h (src/util/ClusterHits.cpp:455-483) and the TSV of summarizeresults (SummarizeResults.cpp:77-114)
    sddb::DbWriter wc(tmp + "/clusters", sddb::DBTYPE_ALIGNMENT_RES), wh(tmp + "/clusters_h", sddb::DBTYPE_GENERIC_DB);
    FILE *tsv = fopen(out.c_str(), "wb");
    if (!tsv) die("cannot write %s", out.c_str());
    std::string body, hdr, tsvBuf;
    char num[64];
    for (uint64_t c = 0; c < nC; c++) {
        const sd_cluster &cl = clusters[c];
        body.clear();
        hdr.clear();
        hdr += std::to_string(qdb.setKey[cl.q_set]) + "\t" + std::to_string(tdb.setKey[cl.t_set]) + "\t";
        snprintf(num, sizeof num, "%.3E", cl.p_co);
        std::string pco = num;
        snprintf(num, sizeof num, "%.3E", cl.p_mh);
        std::string pmh = num;
        hdr += pco + "\t" + pmh + "\t" + std::to_string(cl.n) + "\n";
        tsvBuf += "#" + std::to_string(c) + "\t" + qdb.setSource[cl.q_set] + "\t" + tdb.setSource[cl.t_set] + "\t" + pco + "\t" + pmh + "\t" +
                  std::to_string(cl.n) + "\n";
        for (uint32_t m = 0; m < cl.n; m++) {
            const sd_cluster_member &mem = members[cl.first + m];
            const sd_alignment &al = alns[mem.aln];
            int len = sd_format_alignment(&al, tdb.keys[al.target], pool.data(), line.data(), (int)line.size());
            if (len <= 0) die("alignment line does not fit the format buffer");
            // record: qKey tKey pval seqId E qS qE qL tS tE tL cigar  (prefixid + besthitbyset rewrite of column 3)
            const char *p = line.data(), *end = p + len;
            const char *f1 = (const char *)memchr(p, '\t', (size_t)(end - p)) + 1;          // after tKey
            const char *f2 = (const char *)memchr(f1, '\t', (size_t)(end - f1)) + 1;        // after bits
            snprintf(num, sizeof num, "%.3E", mem.pval);
            body += std::to_string(qdb.keys[al.query]) + "\t" + std::to_string(tdb.keys[al.target]) + "\t" + num + "\t";
            body.append(f2, (size_t)(end - f2));
            tsvBuf += ">" + qdb.names[al.query] + "\t" + tdb.names[al.target] + "\t" + num + "\t";
            tsvBuf.append(f2, (size_t)(end - f2));
        }
        wc.add((uint32_t)c, body);
        wh.add((uint32_t)c, hdr);
        if (tsvBuf.size() > (1u << 22)) { fwrite(tsvBuf.data(), 1, tsvBuf.size(), tsv); tsvBuf.clear(); }
    }
    if (!tsvBuf.empty()) fwrite(tsvBuf.data(), 1, tsvBuf.size(), tsv);
    fclose(tsv);
    if (!wc.close(err) || !wh.close(err)) die("%s", err.c_str());
    // post-processing of data/clustersearch.sh:153-165: clusters -> clu_to_seq (first column), swapped into <out>_seq_to_clu
    filterdbTrimToOneColumn(tmp + "/clusters", tmp + "/clu_to_seq");
    swapOneColumnDb(tmp + "/clu_to_seq", out + "_seq_to_clu");
    info("Wrote %llu clusters to %s, %.0f ms; total %.0f ms\n", (unsigned long long)nC, out.c_str(), nowMs() - t0, nowMs() - tAll);
    sd_query_batch_destroy(batch);
    sd_target_index_destroy(idx);
    sd_context_destroy(ctx);
    return EXIT_SUCCESS;
}

}  // namespace


int main(int argc, char **argv) {
    if (argc < 2) {
        fprintf(stderr, "sdb200 (libsdgpu %d): createsetdb | createindex | prefilter | align | search | clustersearch | prefixid | besthitbyset | mergeresultsbyset | combinehits | clusterhits | summarizeresults | filterdb | swapdb | subtractdbs | mergedbs\n", sd_version());
        return EXIT_FAILURE;
    }
    const std::string mod = argv[1];
    if (mod == "createsetdb") return modCreatesetdb(argc, argv);
    if (mod == "createindex") return modCreateindex(argc, argv);
    if (mod == "prefilter") return modPrefilter(argc, argv);
    if (mod == "search") return modSearch(argc, argv);
    if (mod == "rmdb") return modRmdb(argc, argv);
    if (mod == "align") return modAlign(argc, argv);
    if (mod == "clustersearch") return modClustersearch(argc, argv);
    if (mod == "summarizeresults") return modSummarizeresults(argc, argv);
    if (mod == "filterdb") return modFilterdb(argc, argv);
    if (mod == "swapdb") return modSwapdb(argc, argv);
    if (mod == "subtractdbs") return modSubtractdbs(argc, argv);
    if (mod == "mergedbs") return modMergedbs(argc, argv);
    if (mod == "prefixid") return modPrefixid(argc, argv);
    if (mod == "besthitbyset") return modBesthitbyset(argc, argv);
    if (mod == "mergeresultsbyset") return modMergeresultsbyset(argc, argv);
    if (mod == "combinehits") return modCombinehits(argc, argv);
    if (mod == "clusterhits") return modClusterhits(argc, argv);
    die("unknown module '%s'", mod.c_str());
}
