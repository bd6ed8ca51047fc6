// Fused `clustersearch` hot path: prefilter -> align -> besthitbyset/combinehits -> clusterhits in one call, the
// prefilter hit list never leaves HBM (data/clustersearch.sh:112-146 runs these as separate processes that exchange
// flat-file DBs; this is the same chain with the DBs kept on the device).
//
// After the prefilter the query batch is cut into chunks of whole query SETS (every stage after the prefilter is
// independent per query set: alignments per pair, besthitbyset per gene and target set, combinehits / clusterhits per
// (query set, target set)).  Two host threads, each with its own context (stream + scratch), take the chunks in turn; a
// token serialises the GPU-heavy part of the align stage, so one chunk's host post-processing (E-values, sorting,
// aggregation, clustering bookkeeping) runs under the other chunk's Smith-Waterman kernels.
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

extern "C" int sd_clustersearch_run(sd_context *ctx, const sd_target_index *idx, const sd_query_batch *queries, uint32_t query_first,
                                    const sd_prefilter_params *ppar, const sd_align_params *apar, const sd_cluster_params *cpar,
                                    const uint32_t *q_set_of, const uint32_t *q_pos, const uint8_t *q_strand,
                                    const uint32_t *t_set_of, const uint32_t *t_pos, const uint8_t *t_strand, uint32_t nt,
                                    const uint32_t *q_set_size, uint32_t n_q_sets, const uint32_t *t_set_size, uint32_t n_t_sets,
                                    uint64_t db_residues, sd_search_result **out, sd_prefilter_stats *pst, sd_align_stats *ast,
                                    sd_cluster_stats *cst) {
    if (!ctx || !idx || !queries || !ppar || !apar || !cpar || !out || !q_set_of) { sd::setError("sd_clustersearch_run: bad argument"); return SD_EINVAL; }
    *out = nullptr;
    if (ast) memset(ast, 0, sizeof *ast);
    if (cst) memset(cst, 0, sizeof *cst);
    std::unique_ptr<sd_search_result> r(new sd_search_result());
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
            if (q == 0 || q_set_of[q] != q_set_of[q - 1]) {
                if (q > 0 && q_set_of[q] < q_set_of[q - 1]) contiguous = false;   // sets must come in ascending blocks
                setStart.push_back(q);
            }
        const int envChunks = getenv("SD_CHUNKS") ? atoi(getenv("SD_CHUNKS")) : 0;      // test / tuning hook
        uint64_t want = envChunks > 0 ? (uint64_t)envChunks : std::max<uint64_t>(4, (nHits + (20ull << 20) - 1) / (20ull << 20));
        if (!contiguous || setStart.size() < 2 || (envChunks <= 0 && nHits < 4096)) want = 1;
        want = std::min<uint64_t>(want, setStart.size());
        cut.push_back(0);
        for (uint64_t c = 1; c < want; c++) {
            const uint64_t target = nHits * c / want;
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
            if (rc2 == SD_OK) {
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
            sd::parallelFor(nm, [&](uint64_t i) { dst[i] = src[i]; dst[i].aln += (uint32_t)a0; });
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
