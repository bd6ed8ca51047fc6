/*
 * sdgpu.h -- C ABI of libsdgpu.so, the B200 (sm_100a) implementation of the Spacedust
 * `clustersearch --search-mode 0` hot path.
 *
 * The reference (soedinglab/spacedust) has no in-process plugin interface for this path: every
 * stage is a module `int f(int argc, const char **argv, const Command&)`
 * (lib/mmseqs/src/commons/Command.h:94-105) run as its own process by data/clustersearch.sh:112-151
 * and lib/mmseqs/data/workflow/blastp.sh:70,85, talking through flat-file DBs.  The entry points below
 * are therefore what a binding for that path has to offer: one call per reference module body, on
 * plain host pointers, so that a maintainer can replace the body of
 *   Prefiltering::runSplit      (lib/mmseqs/src/prefiltering/Prefiltering.cpp:725-951)  -> sd_prefilter_*
 *   Alignment::run              (lib/mmseqs/src/alignment/Alignment.cpp:248-542)        -> sd_align_*
 *   clusterhits() pair loop     (src/util/ClusterHits.cpp:294-490)                      -> sd_clusterhits_*
 * with a call into this library (INTEGRATION.md shows the stubs).  No torch / C++ types cross the
 * boundary.  All functions return 0 on success and a negative SD_E* code on failure;
 * sd_last_error() returns a thread-local message (the reference prints via Debug(Debug::ERROR) and
 * EXIT(EXIT_FAILURE), lib/mmseqs/src/commons/MMseqsMPI.h:28-36 -- the front-end does the same with
 * this message).  There is no CPU fallback: without a CUDA device every compute entry point fails
 * with SD_ENODEV.
 */
#ifndef SDGPU_H
#define SDGPU_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SD_OK 0
#define SD_EINVAL (-1)   /* bad argument                                  */
#define SD_ENODEV (-2)   /* no CUDA device / CUDA runtime error           */
#define SD_ENOMEM (-3)   /* host or device allocation failed              */
#define SD_ELIMIT (-4)   /* an implementation limit was hit (see message) */
#define SD_ESMALL (-5)   /* caller's output buffer too small              */

typedef struct sd_context sd_context;        /* one per process per GPU            */
typedef struct sd_target_index sd_target_index; /* device-resident target DB + k-mer index */

const char *sd_last_error(void);
int sd_version(void);

/* device < 0: use the current CUDA device. */
int sd_context_create(int device, sd_context **out);
void sd_context_destroy(sd_context *ctx);
/* Number of CUDA kernels launched through this context since creation. */
uint64_t sd_context_launches(const sd_context *ctx);

/* ---- query batches: residues uploaded and mapped once, reusable by the prefilter and align calls ---- */
typedef struct sd_query_batch sd_query_batch;
int sd_query_batch_create(sd_context *ctx, const char *seqs, const uint64_t *off, uint32_t n, sd_query_batch **out);
void sd_query_batch_destroy(sd_query_batch *b);

/* ---- prefilter (reference flags of `prefilter`) ---------------------- */
typedef struct {
    float sensitivity;        /* -s                     (5.7)  */
    int max_seqs;             /* --max-seqs             (300)  */
    int min_ungapped_score;   /* --min-ungapped-score   (15)   */
    float cov_thr;            /* -c                     (0.8)  */
    int cov_mode;             /* --cov-mode             (2)    */
    int mask;                 /* --mask                 (1)    */
    float mask_prob;          /* --mask-prob            (0.9)  */
    int comp_bias_corr;       /* --comp-bias-corr       (1)    */
    int bin_size;             /* tie-order radix of the reference when --max-seqs truncates; 0 = derive it
                                 like QueryMatcher::initDiagonalMatcher (QueryMatcher.cpp:422-449) from the
                                 host L2 size and the number of targets */
    int same_db;              /* query DB is the target DB: identity hit first with score 65535 */
} sd_prefilter_params;

void sd_prefilter_params_default(sd_prefilter_params *p);

typedef struct {
    uint32_t query;           /* index of the query in the call's query array   */
    uint32_t target;          /* index of the target in the index's target array */
    int32_t score;            /* ungapped diagonal score (65535 for identity)   */
    int16_t diag;             /* i - j, printed as signed short by the reference */
    int16_t pad;
} sd_prefilter_hit;

typedef struct {
    uint64_t probes;          /* P: similar k-mers looked up                    */
    uint64_t matches;         /* M: index entries returned                      */
    uint64_t diag_residues;   /* R: residues read by the diagonal scorer        */
    uint64_t query_residues;  /* sum of query lengths                           */
    uint64_t hits;            /* H: hits before the coverage pre-filter         */
    uint64_t candidates;      /* diagonal-scored (target, diagonal) candidates  */
    uint64_t truncated;       /* queries cut by --max-seqs                      */
    double kernel_ms;         /* device time of the main prefilter kernel(s), CUDA events */
    double total_ms;          /* device time of the whole call incl. copies     */
} sd_prefilter_stats;

/* Build the device-resident target side: residue mapping, tantan masking (on the device), k-mer index
 * (IndexBuilder::fillDatabase, lib/mmseqs/src/prefiltering/IndexBuilder.cpp:55-239).
 * seqs: ASCII residues of all targets concatenated; off[n+1]: byte offsets. */
int sd_target_index_build(sd_context *ctx, const char *seqs, const uint64_t *off, uint32_t n,
                          const sd_prefilter_params *par, sd_target_index **out);
void sd_target_index_destroy(sd_target_index *idx);
/* Copy the masked numeric residues (sum of lengths bytes) back; test hook. */
int sd_target_index_masked(const sd_target_index *idx, uint8_t *out, uint64_t cap);
uint64_t sd_target_index_entries(const sd_target_index *idx);
uint32_t sd_target_index_targets(const sd_target_index *idx);
/* Persisted GPU-ready index (the reference's counterpart: `createindex` + PrefilteringIndexReader, lib/mmseqs/src/prefiltering/
 * PrefilteringIndexReader.cpp): the device buffers of a built index in one file; loading is read + upload, nothing is recomputed.
 * sd_target_index_load refuses a file built with another sensitivity / masking. */
int sd_target_index_save(const sd_target_index *idx, const sd_prefilter_params *par, const char *path);
int sd_target_index_load(sd_context *ctx, const char *path, const sd_prefilter_params *par, sd_target_index **out);

/* Run the prefilter for nq queries against the index (body of the query loop of
 * Prefiltering::runSplit, Prefiltering.cpp:790-880).  Hits are grouped by query in query order;
 * inside a query they are in the reference's output order.  query_first (optional) is the index of
 * query 0 in the target array when same_db is set (identity = query_first + q).
 * Returns SD_ESMALL (and *n_hits = required) when cap is too small. */
int sd_prefilter_run(sd_context *ctx, const sd_target_index *idx, const char *qseqs,
                     const uint64_t *qoff, uint32_t nq, uint32_t query_first,
                     const sd_prefilter_params *par, sd_prefilter_hit *hits, uint64_t cap,
                     uint64_t *n_hits, sd_prefilter_stats *stats);

/* Same with a device-resident query batch (no host->device copy of query residues in the call). */
int sd_prefilter_run_batch(sd_context *ctx, const sd_target_index *idx, const sd_query_batch *queries, uint32_t query_first,
                           const sd_prefilter_params *par, sd_prefilter_hit *hits, uint64_t cap, uint64_t *n_hits,
                           sd_prefilter_stats *stats);

/* ---- align (reference flags of `align`) -------------------------------- */
typedef struct {
    double evalue_thr;        /* -e                (10)   */
    double min_seq_id;        /* --min-seq-id      (0)    */
    float cov_thr;            /* -c                (0.8)  */
    int cov_mode;             /* --cov-mode        (2)    */
    int min_aln_len;          /* --min-aln-len     (30)   */
    int comp_bias_corr;       /* --comp-bias-corr  (1)    */
    int gap_open;             /* --gap-open        (11)   */
    int gap_extend;           /* --gap-extend      (1)    */
    float score_bias;         /* --score-bias      (0)    */
    int same_db;              /* query DB is the target DB: identity pairs take the scoreIdentical shortcut */
} sd_align_params;

void sd_align_params_default(sd_align_params *p);

typedef struct {
    uint32_t query;           /* index into the call's query array                     */
    uint32_t target;          /* index into the index's target array                   */
    int32_t score;            /* raw Smith-Waterman score                              */
    int32_t bits;             /* int(bit score + 0.5)                                  */
    float seq_id;             /* identities / alignment length                         */
    float qcov, tcov;
    double evalue;
    int32_t q_start, q_end, q_len, t_start, t_end, t_len, aln_len;
    uint64_t bt_off;          /* offset of the backtrace string in the backtrace pool  */
    uint32_t bt_len;          /* its length in bytes                                   */
    uint32_t bt_cigar;        /* 1: the string is the run-length CIGAR text ("35M2D20M", what the library returns --
                                 Matcher::compressAlignment, Matcher.cpp:166-185, done on the device); 0: one M/I/D
                                 byte per alignment column (accepted on input, e.g. by sd_format_alignment) */
} sd_alignment;

typedef struct {
    uint64_t alignments;      /* pairs aligned (after the length pre-filter)            */
    uint64_t passed;          /* alignments accepted by Alignment::checkCriteria        */
    uint64_t word_mode;       /* pairs whose 8-bit pass overflowed (re-run in 16-bit semantics) */
    uint64_t reverse_passes;
    uint64_t tracebacks;
    uint64_t cells_forward, cells_reverse, cells_traceback;
    double sw_ms;             /* device time of the score/coordinate kernels, CUDA events */
    double traceback_ms;
    double total_ms;
    uint64_t target_residues; /* sum of target lengths over the aligned pairs (algorithmic bytes: L_t + 64 per pair) */
} sd_align_stats;

/* Body of the query loop of Alignment::run (lib/mmseqs/src/alignment/Alignment.cpp:312-511) for the
 * (query, target) pairs of a prefilter result: pairs must be grouped by query.  Accepted alignments come
 * back grouped by query (ascending query index) and, inside a query, in Matcher::compareHits order
 * (Matcher.h:157-168; its last tie-break, the target's DB key, is the target INDEX here: index the targets in ascending key order,
 * as createdb / createsetdb number them -- sdb200 refuses DBs where that does not hold).  db_residues is the E-value
 * search-space size (0: residues of the index).
 * Returns SD_ESMALL with *n_out / *bt_used = required sizes when a buffer is too small. */
int sd_align_run(sd_context *ctx, const sd_target_index *idx, const char *qseqs, const uint64_t *qoff, uint32_t nq,
                 uint32_t query_first, const uint32_t *pair_q, const uint32_t *pair_t, uint64_t n_pairs,
                 uint64_t db_residues, const sd_align_params *par, sd_alignment *out, uint64_t cap, uint64_t *n_out,
                 char *bt_pool, uint64_t bt_cap, uint64_t *bt_used, sd_align_stats *stats);

int sd_align_run_batch(sd_context *ctx, const sd_target_index *idx, const sd_query_batch *queries, uint32_t query_first,
                       const uint32_t *pair_q, const uint32_t *pair_t, uint64_t n_pairs, uint64_t db_residues,
                       const sd_align_params *par, sd_alignment *out, uint64_t cap, uint64_t *n_out, char *bt_pool,
                       uint64_t bt_cap, uint64_t *bt_used, sd_align_stats *stats);

/* Matcher::resultToBuffer (Matcher.cpp:280-327): one text line "tKey bits seqId E qS qE qL tS tE tL cigar\n";
 * returns the number of bytes written (0 if cap is too small). */
int sd_format_alignment(const sd_alignment *a, uint32_t target_key, const char *bt_pool, char *buf, int cap);

/* ---- besthitbyset + combinehits + clusterhits (reference flags) ------- */
typedef struct {
    float alpha;              /* --alpha             (1)    */
    int filter_self_match;    /* --filter-self-match (0)    */
    int max_gene_gap;         /* --max-gene-gap      (3)    */
    int cluster_size;         /* --cluster-size      (2)    */
    double cluster_pval;      /* --cluster-pval      (0.01) */
    double multihit_pval;     /* --multihit-pval     (0.01) */
} sd_cluster_params;

void sd_cluster_params_default(sd_cluster_params *p);

typedef struct {
    uint32_t q_set, t_set;    /* dense set indices (= set keys of createsetdb)           */
    double p_co;              /* clustering + ordering P-value, exp(-score)             */
    double p_mh;              /* multihit P-value of the cluster                        */
    uint32_t n;               /* member hits                                            */
    uint32_t first;           /* offset of the first member in the member array         */
} sd_cluster;

typedef struct {
    uint32_t aln;             /* index into the alignment array given to sd_cluster_run */
    uint32_t pad;
    double pval;              /* per-hit P-value as the reference prints it in column 3 */
} sd_cluster_member;

typedef struct {
    uint64_t best_hits;       /* (query gene, target set) best hits after besthitbyset  */
    uint64_t genome_pairs;    /* (query set, target set) matches after combinehits      */
    uint64_t pairs_clustered; /* matches with at least two hits                         */
    uint64_t max_hits_per_pair;
    uint64_t merges;          /* agglomeration steps executed on the device             */
    uint64_t inexact_log;     /* cluster sizes beyond the host-precomputed log table (device log used) */
    double kernel_ms;         /* device time of the clustering kernel, CUDA events      */
    double total_ms;
} sd_cluster_stats;

/* Bodies of besthitbyset (src/util/besthitbyset.cpp:41-144), mergeresultsbyset, combinehits
 * (src/util/combinehits.cpp:74-234) and clusterhits (src/util/ClusterHits.cpp:294-490) on the accepted
 * alignments of sd_align_run (grouped by ascending query, reference order inside a query).
 * q_set_of/q_pos/q_strand: per query gene its set, gene index in the set (3rd-last field of the lookup name) and
 * strand (1: start < end); same for targets.  Sets are dense indices in ascending key order.
 * Clusters come back in the reference's order: ascending (query set, target set), then node index. */
int sd_cluster_run(sd_context *ctx, const sd_alignment *alns, uint64_t n_alns,
                   const uint32_t *q_set_of, const uint32_t *q_pos, const uint8_t *q_strand, uint32_t nq,
                   const uint32_t *t_set_of, const uint32_t *t_pos, const uint8_t *t_strand, uint32_t nt,
                   const uint32_t *q_set_size, uint32_t n_q_sets, const uint32_t *t_set_size, uint32_t n_t_sets,
                   const sd_cluster_params *par, sd_cluster *clusters, uint64_t cap, uint64_t *n_clusters,
                   sd_cluster_member *members, uint64_t mcap, uint64_t *n_members, sd_cluster_stats *stats);

/* The clusterhits module body alone (src/util/ClusterHits.cpp:294-485), for matches read from a `matches` / `matches_h` DB pair
 * (`sdb200 clusterhits`): match i owns hits first[i] .. first[i+1]-1; per hit the gene index of the query / target gene in its set,
 * whether both genes lie on the same strand, and the P-value column; nq[i] = number of genes of the match's query set (header
 * column 3); q_set / t_set are passed through to the clusters (set KEYS here).  members[].aln = index of the hit. */
int sd_clusterhits_run(sd_context *ctx, uint32_t n_matches, const uint64_t *first, const uint32_t *q_set, const uint32_t *t_set,
                       const uint32_t *nq, const uint32_t *hit_q_pos, const uint32_t *hit_t_pos, const uint8_t *hit_same_strand,
                       const double *hit_pval, uint32_t max_set_size, const sd_cluster_params *par, sd_cluster *clusters, uint64_t cap,
                       uint64_t *n_clusters, sd_cluster_member *members, uint64_t mcap, uint64_t *n_members, sd_cluster_stats *stats);

/* ---- fused hot path: prefilter -> align -> besthitbyset/combinehits -> clusterhits (data/clustersearch.sh:112-146) ----
 * Same results as calling the three stages one after the other, but the prefilter hit list stays in HBM.  The result
 * object owns host copies of the accepted alignments (+ CIGAR strings) and the clusters.  Internally the query batch is cut
 * into chunks of whole query sets that two host threads process in turn (host post-processing of one chunk under the kernels of
 * the next); q_set_of must list the sets in ascending contiguous blocks for that, otherwise the call runs as one chunk. */
typedef struct sd_search_result sd_search_result;
int sd_clustersearch_run(sd_context *ctx, const sd_target_index *idx, const sd_query_batch *queries, uint32_t query_first,
                         const sd_prefilter_params *ppar, const sd_align_params *apar, const sd_cluster_params *cpar,
                         const uint32_t *q_set_of, const uint32_t *q_pos, const uint8_t *q_strand,
                         const uint32_t *t_set_of, const uint32_t *t_pos, const uint8_t *t_strand, uint32_t nt,
                         const uint32_t *q_set_size, uint32_t n_q_sets, const uint32_t *t_set_size, uint32_t n_t_sets,
                         uint64_t db_residues, sd_search_result **out, sd_prefilter_stats *pst, sd_align_stats *ast,
                         sd_cluster_stats *cst);
void sd_search_result_sizes(const sd_search_result *r, uint64_t *n_alns, uint64_t *pool_bytes, uint64_t *n_clusters, uint64_t *n_members);
int sd_search_result_copy(const sd_search_result *r, sd_alignment *alns, char *pool, sd_cluster *clusters, sd_cluster_member *members);
void sd_search_result_destroy(sd_search_result *r);

/* ---- multi-GPU: one process per GPU, NCCL over NVLink (csrc/shard.cu) -------------------------------------------------
 * The reference's counterpart is its MPI mode (Prefiltering.cpp:273-377 splits the target DB per node and merges the per-split
 * lists, :662-723; Alignment/ClusterHits run per DB chunk): here the target index is replicated, the QUERY SETS are split into
 * contiguous residue-balanced blocks (sd_shard_partition), every rank runs sd_clustersearch_run on its block with query_first =
 * first gene of the block, and ONE all-gather leaves every rank with all ranks' result lists. */
#define SD_NCCL_ID_BYTES 128
int sd_shard_unique_id(uint8_t *id /* SD_NCCL_ID_BYTES, filled by rank 0 and sent to the other ranks by the caller */);
int sd_shard_init(sd_context *ctx, int rank, int world, const uint8_t *id);
void sd_shard_finalize(sd_context *ctx);
int sd_shard_rank(const sd_context *ctx);
int sd_shard_world(const sd_context *ctx);
/* first_set[world + 1]: rank r owns query sets first_set[r] .. first_set[r+1]-1 */
int sd_shard_partition(const uint64_t *set_residues, uint32_t n_sets, int world, uint32_t *first_set);
/* mine: this rank's result (query indices relative to its block; query_first makes them global); *all: every rank's lists in
 * rank order = ascending query sets, same layout a single-GPU run returns (it refers to the context's receive buffer: copy it out with
 * sd_search_result_copy before the next exchange on this context); bytes_gathered: payload bytes over NVLink. */
int sd_shard_allgather_result(sd_context *ctx, const sd_search_result *mine, uint32_t query_first, sd_search_result **all,
                              uint64_t *bytes_gathered);


/* Gene-granular sharding (the layout `sdb200 clustersearch --gpus N` and bench.py use): query GENES, not query sets, are the
 * unit of the expensive stages, so a job with few large sets (20 proteomes over 8 GPUs) is still balanced.
 * sd_shard_partition_genes: first_gene[world + 1], contiguous residue-balanced gene blocks (gene_off: n_genes + 1 offsets).
 * sd_shard_clustersearch_run (after sd_shard_init; with world == 1 it equals sd_clustersearch_run):
 *   1. prefilter + align of this rank's genes (`queries` = genes query_first .. query_first + nq - 1 of the job),
 *   2. ncclAllGather of the alignment lists (every rank then holds all alignments in ascending gene order),
 *   3. besthitbyset / combinehits / clusterhits for this rank's block of query SETS (blocks balanced by alignments; these stages
 *      need all genes of a set: src/util/besthitbyset.cpp:41-144, ClusterHits.cpp:359-454),
 *   4. ncclAllGather of the cluster lists.
 * q_set_of / q_pos / q_strand describe ALL nq_total query genes of the job (sets in ascending contiguous blocks); the result is
 * the whole job's, identical to a single-GPU run, and refers to the context's receive buffers like sd_shard_allgather_result. */
int sd_shard_partition_genes(const uint64_t *gene_off, uint32_t n_genes, int world, uint32_t *first_gene);
int sd_shard_clustersearch_run(sd_context *ctx, const sd_target_index *idx, const sd_query_batch *queries, uint32_t query_first,
                               const sd_prefilter_params *ppar, const sd_align_params *apar, const sd_cluster_params *cpar,
                               const uint32_t *q_set_of, const uint32_t *q_pos, const uint8_t *q_strand, uint32_t nq_total,
                               const uint32_t *t_set_of, const uint32_t *t_pos, const uint8_t *t_strand, uint32_t nt,
                               const uint32_t *q_set_size, uint32_t n_q_sets, const uint32_t *t_set_size, uint32_t n_t_sets,
                               uint64_t db_residues, sd_search_result **out, sd_prefilter_stats *pst, sd_align_stats *ast,
                               sd_cluster_stats *cst, uint64_t *bytes_gathered);

#ifdef __cplusplus
}
#endif
#endif
