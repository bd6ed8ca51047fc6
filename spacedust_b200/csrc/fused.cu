// Fused `clustersearch` hot path: prefilter -> align -> besthitbyset/combinehits -> clusterhits in one call, the
// prefilter hit list never leaves HBM (data/clustersearch.sh:112-146 runs these as separate processes that exchange
// flat-file DBs; this is the same chain with the DBs kept on the device).
//
// After the prefilter the query batch is cut into chunks of whole query SETS (every stage after the prefilter is
// independent per query set: alignments per pair, besthitbyset per gene and target set, combinehits / clusterhits per
// (query set, target set)).  Two host threads, each with its own context (stream + scratch), take the chunks in turn; a
// token serialises the GPU-heavy part of the align stage, so one chunk's host post-processing (E-values, sorting,
// aggregation, clustering bookkeeping) runs under the other chunk's Smith-Waterman kernels.
#include <chrono>
#include <cstring>
#include <memory>
#include <mutex>

#include "sd_internal.h"

namespace {

struct ChunkResult {
    std::vector<sd_alignment> alns;
    sd::CharPool pool;
    std::vector<sd_cluster> clusters;
    std::vector<sd_cluster_member> members;
    // merged result of an exchange: views into the context's pinned receive buffer instead of owned vectors (valid until the
    // next exchange on that context; the caller copies the result out right away)
    const sd_alignment *vAlns = nullptr;
    const char *vPool = nullptr;
    const sd_cluster *vClusters = nullptr;
    const sd_cluster_member *vMembers = nullptr;
    uint64_t vNa = 0, vNp = 0, vNc = 0, vNm = 0;
    bool view = false;
    bool membersGlobal = false;          // member.aln already indexes the merged alignment array (gene-sharded step)
    uint64_t nAlns() const { return view ? vNa : alns.size(); }
    uint64_t nPool() const { return view ? vNp : pool.size(); }
    uint64_t nClusters() const { return view ? vNc : clusters.size(); }
    uint64_t nMembers() const { return view ? vNm : members.size(); }
    const sd_alignment *pAlns() const { return view ? vAlns : alns.data(); }
    const char *pPool() const { return view ? vPool : pool.data(); }
    const sd_cluster *pClusters() const { return view ? vClusters : clusters.data(); }
    const sd_cluster_member *pMembers() const { return view ? vMembers : members.data(); }
    sd_align_stats ast;
    sd_cluster_stats cst;
    int rc = SD_OK;
    std::string err;
};

}  // namespace

struct sd_search_result {
    std::vector<ChunkResult> chunks;     // ascending query sets
};

// cpar == nullptr: prefilter + align only (the gene-sharded multi-GPU step clusters after its first exchange); chunks are then cut
// at any query, not at set boundaries, and q_* / *_set_size are not read.
static int searchCore(sd_context *ctx, const sd_target_index *idx, const sd_query_batch *queries, uint32_t query_first,
                      const sd_prefilter_params *ppar, const sd_align_params *apar, const sd_cluster_params *cpar,
                      const uint32_t *q_set_of, const uint32_t *q_pos, const uint8_t *q_strand,
                      const uint32_t *t_set_of, const uint32_t *t_pos, const uint8_t *t_strand, uint32_t nt,
                      const uint32_t *q_set_size, uint32_t n_q_sets, const uint32_t *t_set_size, uint32_t n_t_sets,
                      uint64_t db_residues, sd_search_result **out, sd_prefilter_stats *pst, sd_align_stats *ast,
                      sd_cluster_stats *cst) {
    const bool doCluster = cpar != nullptr;
    if (!ctx || !idx || !queries || !ppar || !apar || !out || (doCluster && !q_set_of)) { sd::setError("sd_clustersearch_run: bad argument"); return SD_EINVAL; }
    *out = nullptr;
    if (ast) memset(ast, 0, sizeof *ast);
    if (cst) memset(cst, 0, sizeof *cst);
    std::unique_ptr<sd_search_result> r(new sd_search_result());
    if (queries->nq == 0) {      // a rank without query genes (more GPUs than genes): an empty result takes part in the exchanges
        if (pst) memset(pst, 0, sizeof *pst);
        *out = r.release();
        return SD_OK;
    }
    sd::DevBuf<sd_prefilter_hit> dHits;
    uint64_t nHits = 0;
    std::vector<uint64_t> hitOff;
    int rc = sd_prefilter_core(ctx, idx, queries, query_first, ppar, nullptr, 0, &nHits, pst, &dHits, &hitOff);
    if (rc) return rc;
    SD_CUDA(cudaStreamSynchronize(ctx->stream));      // the hit list is read from the helper context's stream too
    const uint32_t nq = queries->nq;

    // ---- chunks of whole query sets, about equal numbers of (query, target) pairs
    std::vector<uint32_t> cut;        // query index where a chunk starts; last = nq
    {
        bool contiguous = true;
        std::vector<uint32_t> setStart;                     // first query of every run of equal set ids
        for (uint32_t q = 0; q < nq; q++)
            if (!doCluster) {
                if ((q & 63u) == 0) setStart.push_back(q);                       // align only: any query may open a chunk
            } else if (q == 0 || q_set_of[q] != q_set_of[q - 1]) {
                if (q > 0 && q_set_of[q] < q_set_of[q - 1]) contiguous = false;   // sets must come in ascending blocks
                setStart.push_back(q);
            }
        const int envChunks = getenv("SD_CHUNKS") ? atoi(getenv("SD_CHUNKS")) : 0;      // test / tuning hook
        // Every chunk pays the critical paths of its small launches (16-bit re-runs, reverse pass, traceback: ~10 ms each on C2, set
        // by the longest pairs), a second chunk lets the host post-processing of the first hide under its kernels: one chunk for small
        // batches (a rank's share of a multi-GPU step), two otherwise (measured on C2: 4 chunks 997 ms, 3 chunks 981, 2 chunks 962),
        // more only to bound the scratch of very large batches (~20 M prefilter hits per chunk)
        // (without the cluster stages -- the gene-sharded multi-GPU step -- a chunk's host part is only E-values, sorting and the CIGAR
        // pool, so the second chunk pays off later -- C2 at 4 GPUs, 2.5 M hits per rank: 330 ms with two chunks, 306 ms with one; SD_CHUNK2_MIN /
        // SD_CHUNK2_MIN_ALIGN override the two thresholds)
        static const uint64_t twoFrom = getenv("SD_CHUNK2_MIN") ? strtoull(getenv("SD_CHUNK2_MIN"), nullptr, 10) : 2000000ull;
        static const uint64_t twoFromAlign = getenv("SD_CHUNK2_MIN_ALIGN") ? strtoull(getenv("SD_CHUNK2_MIN_ALIGN"), nullptr, 10) : 4000000ull;
        uint64_t want = envChunks > 0 ? (uint64_t)envChunks
                                      : std::max<uint64_t>(nHits >= (doCluster ? twoFrom : twoFromAlign) ? 2 : 1, (nHits + (20ull << 20) - 1) / (20ull << 20));
        if (!contiguous || setStart.size() < 2 || (envChunks <= 0 && nHits < 4096)) want = 1;
        want = std::min<uint64_t>(want, setStart.size());
        cut.push_back(0);
        // the host post-processing of the LAST chunk is not hidden under any kernel: it gets 40 % of an even share
        const double evenShare = 1.0 / ((double)want - 0.6);
        for (uint64_t c = 1; c < want; c++) {
            const uint64_t target = want > 1 ? (uint64_t)((double)nHits * evenShare * (double)c) : 0;
            // first set boundary whose hit offset reaches the target
            uint32_t best = nq;
            for (uint32_t sIdx = 0; sIdx < setStart.size(); sIdx++)
                if (hitOff[setStart[sIdx]] >= target) { best = setStart[sIdx]; break; }
            if (best > cut.back() && best < nq) cut.push_back(best);
        }
        cut.push_back(nq);
    }
    const size_t nChunks = cut.size() - 1;
    r->chunks.resize(nChunks);

    // ---- helper context for the second thread (created once, kept with the context)
    const bool twoThreads = nChunks > 1;
    if (twoThreads && !ctx->helper) {
        sd_context *h = nullptr;
        if ((rc = sd_context_create(ctx->device, &h))) return rc;
        ctx->helper = h;
    }
    std::mutex token;
    std::atomic<size_t> nextChunk{0};
    auto worker = [&](sd_context *c) {
        cudaSetDevice(c->device);
        for (;;) {
            const size_t k = nextChunk.fetch_add(1);
            if (k >= nChunks) break;
            ChunkResult &cr = r->chunks[k];
            memset(&cr.ast, 0, sizeof cr.ast);
            memset(&cr.cst, 0, sizeof cr.cst);
            const uint64_t lo = hitOff[cut[k]], hi = hitOff[cut[k + 1]];
            int rc2 = sd_align_device_hits(c, idx, queries, query_first, dHits.p + lo, hi - lo, db_residues, apar, cr.alns, cr.pool, &cr.ast);
            if (rc2 == SD_OK && doCluster) {
                uint64_t nC = 0, nM = 0;
                cr.clusters.resize(cr.alns.size() / 2 + 1);
                cr.members.resize(cr.alns.size() + 1);
                rc2 = sd_cluster_run(c, cr.alns.data(), cr.alns.size(), q_set_of, q_pos, q_strand, nq, t_set_of, t_pos, t_strand, nt,
                                     q_set_size, n_q_sets, t_set_size, n_t_sets, cpar, cr.clusters.data(), cr.clusters.size(), &nC,
                                     cr.members.data(), cr.members.size(), &nM, &cr.cst);
                if (rc2 == SD_OK) { cr.clusters.resize(nC); cr.members.resize(nM); }
            }
            if (rc2 != SD_OK) { cr.rc = rc2; cr.err = sd_last_error(); nextChunk.store(nChunks); break; }
        }
    };
    if (twoThreads) {
        ctx->gpuToken = &token;
        ctx->helper->gpuToken = &token;
        const uint64_t l0 = ctx->helper->launches;
        std::thread t(worker, ctx->helper);
        worker(ctx);
        t.join();
        ctx->launches += ctx->helper->launches - l0;
        ctx->gpuToken = nullptr;
        ctx->helper->gpuToken = nullptr;
    } else {
        worker(ctx);
    }
    dHits.release();
    for (const ChunkResult &cr : r->chunks)
        if (cr.rc != SD_OK) { sd::setError("%s", cr.err.c_str()); return cr.rc; }
    // ---- statistics: sums over the chunks (stage times are device times per chunk; chunks overlap in wall-clock)
    for (const ChunkResult &cr : r->chunks) {
        if (ast) {
            ast->alignments += cr.ast.alignments; ast->passed += cr.ast.passed; ast->word_mode += cr.ast.word_mode;
            ast->reverse_passes += cr.ast.reverse_passes; ast->tracebacks += cr.ast.tracebacks;
            ast->cells_forward += cr.ast.cells_forward; ast->cells_reverse += cr.ast.cells_reverse; ast->cells_traceback += cr.ast.cells_traceback; ast->target_residues += cr.ast.target_residues;
            ast->sw_ms += cr.ast.sw_ms; ast->traceback_ms += cr.ast.traceback_ms; ast->total_ms += cr.ast.total_ms;
        }
        if (cst) {
            cst->best_hits += cr.cst.best_hits; cst->genome_pairs += cr.cst.genome_pairs; cst->pairs_clustered += cr.cst.pairs_clustered;
            cst->max_hits_per_pair = std::max(cst->max_hits_per_pair, cr.cst.max_hits_per_pair);
            cst->merges += cr.cst.merges; cst->inexact_log += cr.cst.inexact_log;
            cst->kernel_ms += cr.cst.kernel_ms; cst->total_ms += cr.cst.total_ms;
        }
    }
    *out = r.release();
    return SD_OK;
}

extern "C" int sd_clustersearch_run(sd_context *ctx, const sd_target_index *idx, const sd_query_batch *queries, uint32_t query_first,
                                    const sd_prefilter_params *ppar, const sd_align_params *apar, const sd_cluster_params *cpar,
                                    const uint32_t *q_set_of, const uint32_t *q_pos, const uint8_t *q_strand,
                                    const uint32_t *t_set_of, const uint32_t *t_pos, const uint8_t *t_strand, uint32_t nt,
                                    const uint32_t *q_set_size, uint32_t n_q_sets, const uint32_t *t_set_size, uint32_t n_t_sets,
                                    uint64_t db_residues, sd_search_result **out, sd_prefilter_stats *pst, sd_align_stats *ast,
                                    sd_cluster_stats *cst) {
    if (!cpar) { sd::setError("sd_clustersearch_run: bad argument"); return SD_EINVAL; }
    return searchCore(ctx, idx, queries, query_first, ppar, apar, cpar, q_set_of, q_pos, q_strand, t_set_of, t_pos, t_strand, nt,
                      q_set_size, n_q_sets, t_set_size, n_t_sets, db_residues, out, pst, ast, cst);
}

// Multi-GPU step with gene-granular sharding (see include/sdgpu.h).  Query genes, not query sets, are the unit of the expensive
// stages: with few large sets (20 proteomes over 8 GPUs) whole-set blocks leave a 3-vs-2 imbalance, gene blocks do not.
//   1. prefilter + align of this rank's genes [query_first, query_first + nq)           (searchCore without the cluster stage)
//   2. ncclAllGather #1: alignments + CIGAR pools -> every rank holds all alignments, ascending genes
//   3. besthitbyset / combinehits / clusterhits for this rank's block of query SETS (blocks balanced by alignment count;
//      those stages need every gene of a set and every alignment of it, src/util/besthitbyset.cpp:41-144)
//   4. ncclAllGather #2: clusters + members (member.aln already indexes the merged alignment array)
extern "C" int sd_shard_clustersearch_run(sd_context *ctx, const sd_target_index *idx, const sd_query_batch *queries, uint32_t query_first,
                                          const sd_prefilter_params *ppar, const sd_align_params *apar, const sd_cluster_params *cpar,
                                          const uint32_t *q_set_of, const uint32_t *q_pos, const uint8_t *q_strand, uint32_t nq_total,
                                          const uint32_t *t_set_of, const uint32_t *t_pos, const uint8_t *t_strand, uint32_t nt,
                                          const uint32_t *q_set_size, uint32_t n_q_sets, const uint32_t *t_set_size, uint32_t n_t_sets,
                                          uint64_t db_residues, sd_search_result **out, sd_prefilter_stats *pst, sd_align_stats *ast,
                                          sd_cluster_stats *cst, uint64_t *bytes_gathered) {
    if (!ctx || !queries || !cpar || !q_set_of || !out) { sd::setError("sd_shard_clustersearch_run: bad argument"); return SD_EINVAL; }
    *out = nullptr;
    if (bytes_gathered) *bytes_gathered = 0;
    if (cst) memset(cst, 0, sizeof *cst);
    if ((uint64_t)query_first + queries->nq > nq_total) { sd::setError("sd_shard_clustersearch_run: gene block %u + %u exceeds %u query genes", query_first, queries->nq, nq_total); return SD_EINVAL; }
    const int world = ctx->world, rank = ctx->rank;
    if (!ctx->comm || world == 1) {      // no communicator: the block is the whole job
        if (query_first != 0 || queries->nq != nq_total) { sd::setError("sd_shard_clustersearch_run: call sd_shard_init first (the gene block is not the whole job)"); return SD_EINVAL; }
        return searchCore(ctx, idx, queries, 0, ppar, apar, cpar, q_set_of, q_pos, q_strand, t_set_of, t_pos, t_strand, nt, q_set_size, n_q_sets,
                          t_set_size, n_t_sets, db_residues, out, pst, ast, cst);
    }
    const bool timing = getenv("SD_TIMING") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    std::string tLog;
    auto tick = [&](const char *what) {
        if (!timing) return;
        const auto t1 = std::chrono::steady_clock::now();
        char b[64];
        snprintf(b, sizeof b, " %s %.1f", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        tLog += b;
        t0 = t1;
    };
    // 1
    sd_search_result *mineRaw = nullptr;
    int rc = searchCore(ctx, idx, queries, query_first, ppar, apar, nullptr, nullptr, nullptr, nullptr, t_set_of, t_pos, t_strand, nt,
                        nullptr, 0, t_set_size, n_t_sets, db_residues, &mineRaw, pst, ast, nullptr);
    if (rc) return rc;
    std::unique_ptr<sd_search_result> mine(mineRaw);
    tick("search");
    // 2
    sd_search_result *allRaw = nullptr;
    uint64_t bytes1 = 0, bytes2 = 0;
    if ((rc = sd::shardAllgather(ctx, mine.get(), query_first, 1, &allRaw, &bytes1))) return rc;
    std::unique_ptr<sd_search_result> all(allRaw);
    mine.reset();
    tick("gather1");
    // 3: alignments per query set (every rank computes the same partition from the same merged list)
    std::vector<uint64_t> setAlns(n_q_sets, 0), chunkBase(all->chunks.size() + 1, 0);
    for (size_t k = 0; k < all->chunks.size(); k++) chunkBase[k + 1] = chunkBase[k] + all->chunks[k].nAlns();
    const uint64_t nAll = chunkBase.back();
    {
        const unsigned nPart = 32;
        std::vector<std::vector<uint64_t>> part(nPart, std::vector<uint64_t>(n_q_sets, 0));
        for (const ChunkResult &cr : all->chunks) {
            const sd_alignment *a = cr.pAlns();
            sd::parallelChunks(cr.nAlns(), nPart, [&](unsigned c, uint64_t lo, uint64_t hi) {
                for (uint64_t i = lo; i < hi; i++) {
                    if (a[i].query < nq_total) part[c][q_set_of[a[i].query]]++;
                }
            });
        }
        for (unsigned c = 0; c < nPart; c++)
            for (uint32_t sIdx = 0; sIdx < n_q_sets; sIdx++) setAlns[sIdx] += part[c][sIdx];
    }
    std::vector<uint64_t> weight(n_q_sets);
    for (uint32_t sIdx = 0; sIdx < n_q_sets; sIdx++) weight[sIdx] = setAlns[sIdx] + 1;
    std::vector<uint32_t> firstSet((size_t)world + 1);
    if ((rc = sd_shard_partition(weight.data(), n_q_sets, world, firstSet.data()))) return rc;
    uint64_t a0 = 0, a1 = 0;
    for (uint32_t sIdx = 0; sIdx < firstSet[rank]; sIdx++) a0 += setAlns[sIdx];
    a1 = a0;
    for (uint32_t sIdx = firstSet[rank]; sIdx < firstSet[rank + 1]; sIdx++) a1 += setAlns[sIdx];
    if (a1 > nAll) { sd::setError("sd_shard_clustersearch_run: query sets are not contiguous ascending blocks of genes"); return SD_EINVAL; }
    tick("partition");
    std::unique_ptr<sd_search_result> mineC(new sd_search_result());
    mineC->chunks.resize(1);
    ChunkResult &cc = mineC->chunks[0];
    memset(&cc.ast, 0, sizeof cc.ast);
    memset(&cc.cst, 0, sizeof cc.cst);
    {
        std::vector<sd_alignment> myAlns(a1 - a0);
        for (size_t k = 0; k < all->chunks.size(); k++) {       // the range may span several rank blocks
            const uint64_t lo = std::max(a0, chunkBase[k]), hi = std::min(a1, chunkBase[k + 1]);
            if (lo >= hi) continue;
            const sd_alignment *src = all->chunks[k].pAlns() + (lo - chunkBase[k]);
            sd_alignment *dst = myAlns.data() + (lo - a0);
            sd::parallelChunks(hi - lo, 32, [&](unsigned, uint64_t l, uint64_t h) { memcpy(dst + l, src + l, (h - l) * sizeof(sd_alignment)); });
        }
        for (uint64_t i = 0; i < myAlns.size(); i += std::max<uint64_t>(1, myAlns.size() - 1)) {     // first and last belong to this rank's sets
            const uint32_t sIdx = q_set_of[myAlns[i].query];
            if (sIdx < firstSet[rank] || sIdx >= firstSet[rank + 1]) { sd::setError("sd_shard_clustersearch_run: query sets are not contiguous ascending blocks of genes"); return SD_EINVAL; }
        }
        uint64_t nC = 0, nM = 0;
        cc.clusters.resize(myAlns.size() / 2 + 1);
        cc.members.resize(myAlns.size() + 1);
        rc = sd_cluster_run(ctx, myAlns.data(), myAlns.size(), q_set_of, q_pos, q_strand, nq_total, t_set_of, t_pos, t_strand, nt, q_set_size, n_q_sets,
                            t_set_size, n_t_sets, cpar, cc.clusters.data(), cc.clusters.size(), &nC, cc.members.data(), cc.members.size(), &nM, &cc.cst);
        if (rc) return rc;
        cc.clusters.resize(nC);
        cc.members.resize(nM);
        sd_cluster_member *m = cc.members.data();
        sd::parallelFor(nM, [&](uint64_t i) { m[i].aln += (uint32_t)a0; });
        if (cst) *cst = cc.cst;
    }
    tick("cluster");
    // 4
    sd_search_result *allCRaw = nullptr;
    if ((rc = sd::shardAllgather(ctx, mineC.get(), 0, 3, &allCRaw, &bytes2))) return rc;
    std::unique_ptr<sd_search_result> allC(allCRaw);
    for (ChunkResult &cr : allC->chunks) { cr.membersGlobal = true; all->chunks.push_back(std::move(cr)); }
    tick("gather2");
    if (timing) fprintf(stderr, "[sd_shard rank %d]%s (ms; prefilter %.1f align %.1f)\n", rank, tLog.c_str(), pst ? pst->total_ms : 0.0, ast ? ast->total_ms : 0.0);
    if (bytes_gathered) *bytes_gathered = bytes1 + bytes2;
    *out = all.release();
    return SD_OK;
}

extern "C" void sd_search_result_sizes(const sd_search_result *r, uint64_t *n_alns, uint64_t *pool_bytes, uint64_t *n_clusters, uint64_t *n_members) {
    uint64_t a = 0, p = 0, c = 0, m = 0;
    if (r)
        for (const ChunkResult &cr : r->chunks) { a += cr.nAlns(); p += cr.nPool(); c += cr.nClusters(); m += cr.nMembers(); }
    if (n_alns) *n_alns = a;
    if (pool_bytes) *pool_bytes = p;
    if (n_clusters) *n_clusters = c;
    if (n_members) *n_members = m;
}

// Chunk results are concatenated in chunk (= ascending query set) order; the cross references are rebased on the way:
// alignment.bt_off -> pool, member.aln -> alignment array, cluster.first -> member array.
extern "C" int sd_search_result_copy(const sd_search_result *r, sd_alignment *alns, char *pool, sd_cluster *clusters, sd_cluster_member *members) {
    if (!r) { sd::setError("sd_search_result_copy: null result"); return SD_EINVAL; }
    uint64_t a0 = 0, p0 = 0, c0 = 0, m0 = 0;
    for (const ChunkResult &cr : r->chunks) {
        const uint64_t na = cr.nAlns(), np = cr.nPool(), nc = cr.nClusters(), nm = cr.nMembers();
        if (alns && na) {
            sd_alignment *dst = alns + a0;
            const sd_alignment *src = cr.pAlns();
            sd::parallelFor(na, [&](uint64_t i) { dst[i] = src[i]; dst[i].bt_off += p0; });
        }
        if (pool && np) {
            const char *src = cr.pPool();
            char *dst = pool + p0;
            sd::parallelChunks(np, 64, [&](unsigned, uint64_t lo, uint64_t hi) { memcpy(dst + lo, src + lo, hi - lo); });
        }
        if (clusters && nc) {
            sd_cluster *dst = clusters + c0;
            const sd_cluster *src = cr.pClusters();
            sd::parallelFor(nc, [&](uint64_t i) { dst[i] = src[i]; dst[i].first += (uint32_t)m0; });
        }
        if (members && nm) {
            sd_cluster_member *dst = members + m0;
            const sd_cluster_member *src = cr.pMembers();
            const uint32_t rebase = cr.membersGlobal ? 0u : (uint32_t)a0;
            sd::parallelFor(nm, [&](uint64_t i) { dst[i] = src[i]; dst[i].aln += rebase; });
        }
        a0 += na; p0 += np; c0 += nc; m0 += nm;
    }
    return SD_OK;
}

extern "C" void sd_search_result_destroy(sd_search_result *r) { delete r; }

int sd_search_result_from_blocks(const char *blocks, uint64_t stride, const uint64_t *sizes, int n, sd_search_result **out) {
    std::unique_ptr<sd_search_result> r(new sd_search_result());
    r->chunks.resize((size_t)n);
    auto al = [](uint64_t x) { return (x + 15) & ~15ull; };
    for (int k = 0; k < n; k++) {
        const char *b = blocks + (uint64_t)k * stride;
        uint64_t head[4];
        memcpy(head, b, sizeof head);
        const uint64_t na = head[0], np = head[1], nc = head[2], nm = head[3];
        const uint64_t oA = 32, oP = al(oA + na * sizeof(sd_alignment)), oC = al(oP + np), oM = al(oC + nc * sizeof(sd_cluster));
        if (al(oM + nm * sizeof(sd_cluster_member)) != sizes[k]) { sd::setError("result block of rank %d is inconsistent", k); return SD_EINVAL; }
        ChunkResult &cr = r->chunks[(size_t)k];
        memset(&cr.ast, 0, sizeof cr.ast);
        memset(&cr.cst, 0, sizeof cr.cst);
        cr.view = true;
        cr.vAlns = reinterpret_cast<const sd_alignment *>(b + oA); cr.vNa = na;
        cr.vPool = b + oP; cr.vNp = np;
        cr.vClusters = reinterpret_cast<const sd_cluster *>(b + oC); cr.vNc = nc;
        cr.vMembers = reinterpret_cast<const sd_cluster_member *>(b + oM); cr.vNm = nm;
    }
    *out = r.release();
    return SD_OK;
}
