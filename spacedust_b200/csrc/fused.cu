// Fused `clustersearch` hot path: prefilter -> align -> besthitbyset/combinehits -> clusterhits in one call, the
// prefilter hit list never leaves HBM (data/clustersearch.sh:112-146 runs these as separate processes that exchange
// flat-file DBs; this is the same chain with the DBs kept on the device).
#include <cstring>
#include <memory>

#include "sd_internal.h"

struct sd_search_result {
    std::vector<sd_alignment> alns;
    sd::CharPool pool;
    std::vector<sd_cluster> clusters;
    std::vector<sd_cluster_member> members;
};

extern "C" int sd_clustersearch_run(sd_context *ctx, const sd_target_index *idx, const sd_query_batch *queries, uint32_t query_first,
                                    const sd_prefilter_params *ppar, const sd_align_params *apar, const sd_cluster_params *cpar,
                                    const uint32_t *q_set_of, const uint32_t *q_pos, const uint8_t *q_strand,
                                    const uint32_t *t_set_of, const uint32_t *t_pos, const uint8_t *t_strand, uint32_t nt,
                                    const uint32_t *q_set_size, uint32_t n_q_sets, const uint32_t *t_set_size, uint32_t n_t_sets,
                                    uint64_t db_residues, sd_search_result **out, sd_prefilter_stats *pst, sd_align_stats *ast,
                                    sd_cluster_stats *cst) {
    if (!ctx || !idx || !queries || !ppar || !apar || !cpar || !out) { sd::setError("sd_clustersearch_run: bad argument"); return SD_EINVAL; }
    *out = nullptr;
    std::unique_ptr<sd_search_result> r(new sd_search_result());
    sd::DevBuf<sd_prefilter_hit> dHits;
    uint64_t nHits = 0;
    int rc = sd_prefilter_core(ctx, idx, queries, query_first, ppar, nullptr, 0, &nHits, pst, &dHits);
    if (rc) return rc;
    rc = sd_align_device_hits(ctx, idx, queries, query_first, dHits.p, nHits, db_residues, apar, r->alns, r->pool, ast);
    if (rc) return rc;
    dHits.release();
    uint64_t nC = 0, nM = 0;
    r->clusters.resize(r->alns.size() / 2 + 1);
    r->members.resize(r->alns.size() + 1);
    rc = sd_cluster_run(ctx, r->alns.data(), r->alns.size(), q_set_of, q_pos, q_strand, queries->nq, t_set_of, t_pos, t_strand, nt,
                        q_set_size, n_q_sets, t_set_size, n_t_sets, cpar, r->clusters.data(), r->clusters.size(), &nC,
                        r->members.data(), r->members.size(), &nM, cst);
    if (rc) return rc;
    r->clusters.resize(nC);
    r->members.resize(nM);
    *out = r.release();
    return SD_OK;
}

extern "C" void sd_search_result_sizes(const sd_search_result *r, uint64_t *n_alns, uint64_t *pool_bytes, uint64_t *n_clusters, uint64_t *n_members) {
    if (n_alns) *n_alns = r ? r->alns.size() : 0;
    if (pool_bytes) *pool_bytes = r ? r->pool.size() : 0;
    if (n_clusters) *n_clusters = r ? r->clusters.size() : 0;
    if (n_members) *n_members = r ? r->members.size() : 0;
}

extern "C" int sd_search_result_copy(const sd_search_result *r, sd_alignment *alns, char *pool, sd_cluster *clusters, sd_cluster_member *members) {
    if (!r) { sd::setError("sd_search_result_copy: null result"); return SD_EINVAL; }
    if (alns && !r->alns.empty()) memcpy(alns, r->alns.data(), r->alns.size() * sizeof(sd_alignment));
    if (pool && !r->pool.empty()) memcpy(pool, r->pool.data(), r->pool.size());
    if (clusters && !r->clusters.empty()) memcpy(clusters, r->clusters.data(), r->clusters.size() * sizeof(sd_cluster));
    if (members && !r->members.empty()) memcpy(members, r->members.data(), r->members.size() * sizeof(sd_cluster_member));
    return SD_OK;
}

extern "C" void sd_search_result_destroy(sd_search_result *r) { delete r; }
