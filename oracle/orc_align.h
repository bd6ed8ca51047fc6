/* oracle (test infrastructure) -- see orc_common.h */
#ifndef ORC_ALIGN_H
#define ORC_ALIGN_H
#include "orc_common.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    double evalThr;     /* -e              */
    double seqIdThr;    /* --min-seq-id    */
    float covThr;       /* -c              */
    int covMode;        /* --cov-mode      */
    int minAlnLen;      /* --min-aln-len   */
    int compBias;       /* --comp-bias-corr*/
    int gapOpen;        /* --gap-open      */
    int gapExtend;      /* --gap-extend    */
    float scoreBias;    /* --score-bias    */
    int sameDB;         /* query DB is the target DB (identity shortcut) */
} orc_align_params;

typedef struct {
    uint32_t query, target, targetKey;
    int score, bits;
    float seqId, qcov, tcov;
    double evalue;
    int qStart, qEnd, qLen, tStart, tEnd, tLen, alnLen;
    uint64_t btOff;
    uint32_t btLen;
} orc_aln;

typedef struct {
    uint64_t alignments, passed, wordMode, reversePasses, tracebacks, tracebackFailed, scoreMismatch;
    uint64_t cellsForward, cellsReverse, cellsTraceback;
} orc_align_stats;

double orc_evalue(double score, double qLen, double dbResidues);
int orc_bitscore(double score);

/* pairs (pairQ[k], pairT[k]) must be grouped by query, in prefilter order.  Results are grouped by
 * query in that order and sorted inside a query like the reference; backtraces go to btPool. */
long orc_align(const char *tSeq, const uint64_t *tOff, const uint32_t *tKeys, uint32_t nT,
               const char *qSeq, const uint64_t *qOff, const uint32_t *qKeys, uint32_t nQ,
               const uint32_t *pairQ, const uint32_t *pairT, uint64_t nPairs, uint64_t dbResidues,
               const orc_align_params *par, orc_aln *out, long outCap, char *btPool, uint64_t btCap,
               orc_align_stats *stats);
long orc_align_profile(const char *tSeq, const uint64_t *tOff, const uint32_t *tKeys, uint32_t nT,
                       const unsigned char *profData, const uint64_t *qOff, const uint32_t *qKeys, uint32_t nQ,
                       const uint32_t *pairQ, const uint32_t *pairT, uint64_t nPairs, uint64_t dbResidues,
                       const orc_align_params *par, orc_aln *out, long outCap, char *btPool, uint64_t btCap,
                       orc_align_stats *stats);
int orc_format_alignment(const orc_aln *r, const char *btPool, char *buf, int cap);

#ifdef __cplusplus
}
#endif
#endif
