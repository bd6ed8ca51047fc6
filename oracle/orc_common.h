/*
 * oracle/ -- TEST INFRASTRUCTURE ONLY.
 *
 * Scalar CPU restatement of the `spacedust clustersearch --search-mode 0` hot path, written from
 * the behaviour of the reference (soedinglab/spacedust @ /root/reference; every function cites the
 * reference file:line it follows).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product (spacedust_b200/)
 * never includes, links or calls anything under oracle/.
 *
 * Parity pin: this restatement is checked against the UNMODIFIED reference binary built by
 * oracle/build_ref.py (oracle/_ref/spacedust) -- tests/test_oracle_vs_ref.py diffs the prefilter,
 * align, aggregate, matches and clusters DBs -- and against the committed fixtures in tests/golden/.
 */
#ifndef ORC_COMMON_H
#define ORC_COMMON_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_AA 21          /* alphabet incl. X (code 20) */
#define ORC_X 20

typedef struct {
    short sub[ORC_AA * ORC_AA];   /* rounded integer scores              */
    double pBack[ORC_AA];         /* background used for composition bias */
    double prob[ORC_AA * ORC_AA]; /* joint probabilities                  */
} orc_submat;

/* which: 0 = BLOSUM62, 1 = VTML80 */
void orc_build_submat(int which, float bitFactor, float scoreBias, orc_submat *out);
void orc_map_sequence(const char *s, int len, unsigned char *out);
void orc_comp_bias(const orc_submat *m, const unsigned char *seq, int L, float *bias, float scale);
int orc_tantan_mask(unsigned char *seq, int L, const orc_submat *m, double minMaskProb);
int orc_kmer_threshold(float sensitivity, int kmerSize);

typedef struct {
    float sensitivity;     /* -s            */
    int maxSeqs;           /* --max-seqs    */
    int minDiagScore;      /* --min-ungapped-score */
    float covThr;          /* -c            */
    int covMode;           /* --cov-mode    */
    int mask;              /* --mask        */
    float maskProb;        /* --mask-prob   */
    int compBias;          /* --comp-bias-corr */
    int binSize;           /* radix bin count of the reference's CacheFriendlyOperations (host-L2 dependent) */
    int sameDB;            /* query DB == target DB: identity hit first */
    uint64_t maxDbMatches;        /* test hook: size of the hit buffer (0 = reference: 2*max(10^6, nT), QueryMatcher.cpp:43) */
    uint64_t foundDiagonalsSize;  /* test hook: size of the candidate array (0 = reference: max(10^6, nT), QueryMatcher.cpp:42) */
} orc_pref_params;

typedef struct {
    uint64_t probes;       /* P: similar k-mers looked up                      */
    uint64_t matches;      /* M: index entries returned                        */
    uint64_t diagResidues; /* R: residues read by diagonal scoring             */
    uint64_t queryResidues;/* sum L_q                                          */
    uint64_t hits;         /* H: hits written (before the coverage filter)     */
    uint64_t candidates;   /* double-hit candidates that were diagonal-scored  */
    uint64_t truncated;    /* queries where the --max-seqs cut dropped hits    */
    uint64_t overflowQueries;      /* queries whose index hits overflowed the hit buffer at least once (QueryMatcher.cpp:280) */
    uint64_t multiOverflowQueries; /* ... at least twice (reversed merge + intermediate keepMax, :296-303) */
} orc_pref_stats;

typedef struct {
    uint32_t query;        /* query index                                       */
    uint32_t target;       /* target index                                      */
    int32_t score;
    int16_t diag;
} orc_pref_hit;

/* Sequences are ASCII residues, concatenated; off[n+1] byte offsets.  Returns number of hits written
 * (after the coverage filter), or -1 if outCap is too small.  Hits are grouped by query in query
 * order and, inside a query, in the reference's output order. */
long orc_prefilter(const char *tSeq, const uint64_t *tOff, uint32_t nT,
                   const char *qSeq, const uint64_t *qOff, uint32_t nQ,
                   const orc_pref_params *par, orc_pref_hit *out, long outCap,
                   orc_pref_stats *stats, unsigned char *maskedOut /* optional, sum L_t bytes */);
/* profile queries: profData = entries concatenated (25 bytes per column), qOff in columns; sameDB must be 0 */
long orc_prefilter_profile(const char *tSeq, const uint64_t *tOff, uint32_t nT, const unsigned char *profData,
                           const uint64_t *qOff, uint32_t nQ, const orc_pref_params *par, orc_pref_hit *out, long outCap,
                           orc_pref_stats *stats);
int orc_kmer_threshold_profile(float sensitivity, int kmerSize);

#ifdef __cplusplus
}
#endif
#endif
