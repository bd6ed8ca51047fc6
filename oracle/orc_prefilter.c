/* oracle (test infrastructure): scalar restatement of the MMseqs2 k-mer prefilter as Spacedust's
 * clustersearch runs it (k=6 spaced 1101010011, VTML80 seeds, BLOSUM62 diagonal scoring).
 * See orc_common.h for the rules that govern this directory. */
#include "orc_common.h"
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define KMER 6
#define SPAN 10
#define N3 8000          /* 20^3 */
#define NK 64000000      /* 20^6 */
static const int SEED_POS[KMER] = {0, 1, 3, 5, 8, 9};   /* Sequence.h:24 spaced_seed_6 */

/* ---- extended 3-mer matrix: ExtendedSubstitutionMatrix::calcScoreMatrix (ExtendedSubstitutionMatrix.cpp:20-69).
 * Row for 3-mer x: all 3-mers y sorted by score(x,y) descending; std::stable_sort over the cartesian
 * product enumerated first-letter-major, so ties keep (y0,y1,y2) lexicographic order while the stored
 * index is y0 + 20*y1 + 400*y2 (Indexer::int2index, Indexer.h:20-85). */
typedef struct { short *score; unsigned short *index; } ext3_t;

static void build_ext3(const short *sub /*21x21*/, ext3_t *e) {
    e->score = (short *)malloc(sizeof(short) * (size_t)N3 * N3);
    e->index = (unsigned short *)malloc(sizeof(unsigned short) * (size_t)N3 * N3);
    short *tmpS = (short *)malloc(sizeof(short) * N3);
    int cnt[512];
    for (int x0 = 0; x0 < 20; x0++)
        for (int x1 = 0; x1 < 20; x1++)
            for (int x2 = 0; x2 < 20; x2++) {
                int xi = x0 + 20 * x1 + 400 * x2;
                /* counting sort by score (range -256..255), stable in enumeration order */
                memset(cnt, 0, sizeof cnt);
                int n = 0;
                for (int y0 = 0; y0 < 20; y0++)
                    for (int y1 = 0; y1 < 20; y1++)
                        for (int y2 = 0; y2 < 20; y2++) {
                            short s = (short)(sub[x0 * 21 + y0] + sub[x1 * 21 + y1] + sub[x2 * 21 + y2]);
                            tmpS[n++] = s;
                            cnt[255 - s]++;       /* bucket 0 = highest score */
                        }
                int start[512], acc = 0;
                for (int b = 0; b < 512; b++) { start[b] = acc; acc += cnt[b]; }
                n = 0;
                short *rs = e->score + (size_t)xi * N3;
                unsigned short *ri = e->index + (size_t)xi * N3;
                for (int y0 = 0; y0 < 20; y0++)
                    for (int y1 = 0; y1 < 20; y1++)
                        for (int y2 = 0; y2 < 20; y2++) {
                            short s = tmpS[n++];
                            int p = start[255 - s]++;
                            rs[p] = s;
                            ri[p] = (unsigned short)(y0 + 20 * y1 + 400 * y2);
                        }
            }
    free(tmpS);
}

/* ---- target index: IndexBuilder::fillDatabase (IndexBuilder.cpp:55-239), IndexTable::addKmerCount /
 * addSequence (IndexTable.h:133-173, 348-401): k-mers without X whose self score >= kmerThr; each
 * distinct k-mer of a sequence once, at its first position; buckets ordered by (seqId, pos). */
typedef struct { uint32_t seq; uint16_t pos; } entry_t;
typedef struct {
    uint32_t *off;       /* NK + 1 */
    entry_t *ent;
    uint64_t nEnt;
} index_t;

typedef struct { uint32_t kmer; uint16_t pos; } kp_t;
static int cmp_kp(const void *a, const void *b) {
    const kp_t *x = (const kp_t *)a, *y = (const kp_t *)b;
    if (x->kmer != y->kmer) return x->kmer < y->kmer ? -1 : 1;
    if (x->pos != y->pos) return x->pos < y->pos ? -1 : 1;
    return 0;
}

static uint32_t collect_kmers(const unsigned char *s, int L, const short *kmat, int thr, kp_t *buf) {
    uint32_t n = 0;
    for (int p = 0; p + SPAN <= L; p++) {
        int hasX = 0, self = 0;
        uint32_t idx = 0, mul = 1;
        for (int k = 0; k < KMER; k++) {
            unsigned char c = s[p + SEED_POS[k]];
            if (c == ORC_X) { hasX = 1; break; }
            self += (signed char)kmat[c * 21 + c];
            idx += c * mul;
            mul *= 20;
        }
        if (hasX) continue;
        if (thr > 0 && self < thr) continue;
        buf[n].kmer = idx;
        buf[n].pos = (uint16_t)p;
        n++;
    }
    if (n > 1) qsort(buf, n, sizeof(kp_t), cmp_kp);
    uint32_t m = 0;
    for (uint32_t i = 0; i < n; i++)
        if (i == 0 || buf[i].kmer != buf[i - 1].kmer) buf[m++] = buf[i];
    return m;
}

static void build_index(const unsigned char *res, const uint64_t *off, uint32_t nT, const short *kmat,
                        int thr, index_t *ix) {
    ix->off = (uint32_t *)calloc((size_t)NK + 1, sizeof(uint32_t));
    int maxL = 0;
    for (uint32_t t = 0; t < nT; t++) {
        int L = (int)(off[t + 1] - off[t]);
        if (L > maxL) maxL = L;
    }
    kp_t *buf = (kp_t *)malloc(sizeof(kp_t) * (size_t)(maxL + 1));
    for (uint32_t t = 0; t < nT; t++) {
        uint32_t m = collect_kmers(res + off[t], (int)(off[t + 1] - off[t]), kmat, thr, buf);
        for (uint32_t i = 0; i < m; i++) ix->off[buf[i].kmer + 1]++;
    }
    for (size_t k = 0; k < NK; k++) ix->off[k + 1] += ix->off[k];
    ix->nEnt = ix->off[NK];
    ix->ent = (entry_t *)malloc(sizeof(entry_t) * (size_t)(ix->nEnt + 1));
    uint32_t *fill = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)NK);
    memcpy(fill, ix->off, sizeof(uint32_t) * (size_t)NK);
    for (uint32_t t = 0; t < nT; t++) {      /* ascending t => buckets already (seqId,pos)-sorted */
        uint32_t m = collect_kmers(res + off[t], (int)(off[t + 1] - off[t]), kmat, thr, buf);
        for (uint32_t i = 0; i < m; i++) {
            entry_t *e = &ix->ent[fill[buf[i].kmer]++];
            e->seq = t;
            e->pos = buf[i].pos;
        }
    }
    free(fill);
    free(buf);
}

/* ---- ungapped diagonal score: UngappedAlignment::scalarDiagonalScoring / computeSingelSequenceScores
 * (UngappedAlignment.cpp:30-43, 416-430).  profile[pos*21 + aa]; diag = i - j as a signed 16-bit value. */
static int diag_score(const signed char *prof, int Lq, const unsigned char *t, int Lt, int diag,
                      uint64_t *residues) {
    int qs, ts, len;
    if (diag >= 0) {
        if (diag >= Lq) return 0;
        qs = diag; ts = 0;
        len = Lt < Lq - diag ? Lt : Lq - diag;
    } else {
        int d = -diag;
        if (d >= Lt) return 0;
        qs = 0; ts = d;
        len = Lt - d < Lq ? Lt - d : Lq;
    }
    int best = 0, run = 0;
    for (int p = 0; p < len; p++) {
        run += prof[(qs + p) * 21 + t[ts + p]];
        if (run < 0) run = 0;
        if (run > best) best = run;
    }
    if (residues) *residues += (uint64_t)len;
    return best;
}

typedef struct { uint32_t id; uint16_t diag; unsigned char count; } cand_t;
typedef struct { uint32_t id; uint16_t diag; } hit_t;
typedef struct { uint32_t seqId; int score; uint16_t diag; } res_t;

static int cmp_res(const void *a, const void *b) {   /* hit_t::compareHitsByScoreAndId QueryMatcher.h:38-48 */
    const res_t *x = (const res_t *)a, *y = (const res_t *)b;
    int ax = abs(x->score), ay = abs(y->score);
    if (ax != ay) return ax > ay ? -1 : 1;
    if (x->seqId != y->seqId) return x->seqId < y->seqId ? -1 : 1;
    return 0;
}

static int can_be_covered(float covThr, int covMode, float q, float t) {   /* Util.cpp:477-494 */
    switch (covMode) {
        case 0: return (q / t >= covThr) && (t / q >= covThr);
        case 1: return (q / t) >= covThr;
        case 2: return (t / q) >= covThr;
        case 3: return ((t / q) >= covThr) && (t / q) <= 1.0;
        case 4: return ((q / t) >= covThr) && (q / t) <= 1.0;
        case 5: return ((t < q ? t : q) / (t > q ? t : q)) >= covThr;
        default: return 1;
    }
}

/* ---- profile queries (row a22 of SURVEY.md section 8: `prefilter <profileDB> <targetDB>` of the iterative search) ----
 * A profile entry has 25 bytes per column (Sequence.h:458-471): 20 scores (x4 scaled), query letter, consensus, Neff, 2 gap
 * bytes.  Sequence::mapProfile (Sequence.cpp:241-283) sorts every column's 20 scores descending with a fixed 20-input sorting
 * network (Util::rankedDescSort20, Util.cpp:88-114; the comparator order decides ties, so it is kept as a table), the
 * k-mer generator then combines the six columns of a spaced window list by list (KmerGenerator::setDivideStrategy(ScoreMatrix**)
 * :30-39, generateKmerList :107-183, calculateArrayProduct :186-216), the diagonal scorer uses score/4 with X = 0, there is no
 * composition bias (QueryMatcher.cpp:91-96), the k-mer threshold comes from the profile table (Prefiltering.cpp:1019-1048) and
 * the target index keeps every k-mer (threshold 0, Prefiltering.cpp:525-527). */
static const unsigned char SORT20[][2] = {
    {0,16},{1,17},{2,18},{3,19},{4,12},{5,13},{6,14},{7,15},
    {0,8},{1,9},{2,10},{3,11},
    {8,16},{9,17},{10,18},{11,19},{0,4},{1,5},{2,6},{3,7},
    {8,12},{9,13},{10,14},{11,15},{4,16},{5,17},{6,18},{7,19},{0,2},{1,3},
    {4,8},{5,9},{6,10},{7,11},{12,16},{13,17},{14,18},{15,19},{0,1},
    {4,6},{5,7},{8,10},{9,11},{12,14},{13,15},{16,18},{17,19},
    {2,16},{3,17},{6,12},{7,13},{18,19},
    {2,8},{3,9},{10,16},{11,17},
    {2,4},{3,5},{6,8},{7,9},{10,12},{11,13},{14,16},{15,17},
    {2,3},{4,5},{6,7},{8,9},{10,11},{12,13},{14,15},{16,17},
    {1,16},{3,18},{5,12},{7,14},
    {1,8},{3,10},{9,16},{11,18},
    {1,4},{3,6},{5,8},{7,10},{9,12},{11,14},{13,16},{15,18},
    {1,2},{3,4},{5,6},{7,8},{9,10},{11,12},{13,14},{15,16},{17,18}};

static void ranked_desc_sort20(short *val, unsigned int *index) {
    for (size_t c = 0; c < sizeof SORT20 / sizeof SORT20[0]; c++) {
        const int x = SORT20[c][0], y = SORT20[c][1];
        if (val[x] < val[y]) {
            short t1 = val[x]; val[x] = val[y]; val[y] = t1;
            unsigned int t2 = index[x]; index[x] = index[y]; index[y] = t2;
        }
    }
}

int orc_kmer_threshold_profile(float sensitivity, int kmerSize) {     /* Prefiltering.cpp:1033-1046, no context pseudo counts */
    float thr;
    if (kmerSize == 5) { float base = 108.8; thr = base - (sensitivity * 4.7); }
    else if (kmerSize == 6) { float base = 134.35; thr = base - (sensitivity * 6.15); }
    else { float base = 149.15; thr = base - (sensitivity * 6.85); }
    return (int)thr;
}

#define ORC_MAX_KMER_RESULT ((size_t)262144 * 32)

/* similar k-mers of one profile window: six sorted columns combined list by list; returns the list length, order as produced */
static size_t profile_kmer_list(short sc[KMER][20], unsigned int ix[KMER][20], short thr, short **bufS, size_t **bufI,
                                const size_t **result) {
    static const size_t POW20[KMER] = {1, 20, 400, 8000, 160000, 3200000};
    short rest[KMER];
    rest[KMER - 1] = 0;
    for (int i = KMER - 1; i >= 1; i--) rest[i - 1] = (short)(sc[i][0] + rest[i]);
    short cutoff1 = (short)(thr - rest[0]);
    for (int pos = 0; pos < 20 && sc[0][pos] >= cutoff1; pos++) bufI[1][pos] = ix[0][pos];
    const short *inS = sc[0];
    const size_t *inI = bufI[1];
    size_t n = 20;
    int i;
    for (i = 0; i < KMER - 1; i++) {
        short *outS = bufS[i % 2];
        size_t *outI = bufI[i % 2];
        size_t cnt = 0;
        int full = 0;
        for (size_t a = 0; a < n; a++) {
            const short sa = inS[a];
            if (sa < cutoff1) break;
            const short cutoff2 = (short)(thr - sa - rest[i + 1]);
            for (int j = 0; j < 20 && cnt + 1 < ORC_MAX_KMER_RESULT && sc[i + 1][j] >= cutoff2; j++) {
                outS[cnt] = (short)(sa + sc[i + 1][j]);
                outI[cnt] = inI[a] + (size_t)ix[i + 1][j] * POW20[i + 1];
                cnt++;
            }
            if (cnt + 1 >= ORC_MAX_KMER_RESULT) { full = 1; break; }
        }
        (void)full;
        inS = outS;
        inI = outI;
        cutoff1 = -1000;
        n = cnt;
    }
    *result = bufI[(i - 1) % 2];
    return n;
}


/* ---- CacheFriendlyOperations, restated on flat arrays.  The reference first hashes the input into radix bins
 * (id & binMask), stable inside a bin, and then walks bin by bin; `last` is its duplicateBitArray (index id >> binBits). */
typedef struct {
    unsigned char *last;       /* duplicateBitArray */
    cand_t *tmp;               /* tmpElementBuffer / scratch copy */
    size_t tmpCap;
    uint32_t binMask;
    int binBits;
} cfo_t;

static void cfo_reserve(cfo_t *c, size_t n) {
    if (n + 1 > c->tmpCap) {
        while (n + 1 > c->tmpCap) c->tmpCap *= 2;
        c->tmp = (cand_t *)realloc(c->tmp, sizeof(cand_t) * c->tmpCap);
    }
}

/* findDuplicates (CacheFriendlyOperations.cpp:185-274, diagonal-scoring branch): a hit is kept when its 8-bit diagonal equals the
 * previous 8-bit diagonal seen for that target (table starts at 0); kept hits of one target with equal consecutive 8-bit
 * diagonals collapse to the first.  Returns early when the output array would overflow (:208-211). */
static size_t cfo_find_duplicates(cfo_t *c, const hit_t *hits, size_t nHit, cand_t *out, size_t outputSize) {
    size_t nOut = 0;
    cfo_reserve(c, nHit);
    for (uint32_t bin = 0; bin <= c->binMask; bin++) {
        size_t nTmp = 0, binCount = 0;
        for (size_t h = 0; h < nHit; h++) {
            if ((hits[h].id & c->binMask) != bin) continue;
            binCount++;
            uint32_t slot = hits[h].id >> c->binBits;
            unsigned char d8 = (unsigned char)hits[h].diag;
            if (d8 == c->last[slot]) { c->tmp[nTmp].id = hits[h].id; c->tmp[nTmp].diag = hits[h].diag; nTmp++; }
            c->last[slot] = d8;
        }
        if (nOut + (nTmp < binCount / 2 ? nTmp : binCount / 2) >= outputSize) return nOut;
        for (size_t n = nTmp; n-- > 0;) c->last[c->tmp[n].id >> c->binBits] = (unsigned char)((unsigned char)c->tmp[n].diag + 1);
        for (size_t n = 0; n < nTmp; n++) {
            uint32_t slot = c->tmp[n].id >> c->binBits;
            unsigned char d8 = (unsigned char)c->tmp[n].diag;
            out[nOut].id = c->tmp[n].id; out[nOut].diag = c->tmp[n].diag; out[nOut].count = 0;
            if (c->last[slot] != d8) nOut++;
            c->last[slot] = d8;
        }
        for (size_t h = 0; h < nHit; h++)
            if ((hits[h].id & c->binMask) == bin) c->last[hits[h].id >> c->binBits] = 0;
    }
    return nOut;
}

/* mergeDiagonalDuplicates (CacheFriendlyOperations.cpp:84-113): per target, consecutive elements with the same 8-bit
 * diagonal collapse to the first; output bin-major.  The table is NOT cleared afterwards (the caller's next use clears or
 * overwrites it) -- cleared here, which is equivalent for every caller on this path. */
static size_t cfo_merge_diagonal(cfo_t *c, cand_t *arr, size_t N) {
    cfo_reserve(c, N);
    memcpy(c->tmp, arr, N * sizeof(cand_t));
    size_t nOut = 0;
    for (uint32_t bin = 0; bin <= c->binMask; bin++) {
        for (size_t n = N; n-- > 0;)
            if ((c->tmp[n].id & c->binMask) == bin) c->last[c->tmp[n].id >> c->binBits] = (unsigned char)((unsigned char)c->tmp[n].diag + 1);
        for (size_t n = 0; n < N; n++) {
            if ((c->tmp[n].id & c->binMask) != bin) continue;
            uint32_t slot = c->tmp[n].id >> c->binBits;
            unsigned char d8 = (unsigned char)c->tmp[n].diag;
            arr[nOut] = c->tmp[n];
            if (c->last[slot] != d8) nOut++;
            c->last[slot] = d8;
        }
        for (size_t n = 0; n < N; n++)
            if ((c->tmp[n].id & c->binMask) == bin) c->last[c->tmp[n].id >> c->binBits] = 0;
    }
    return nOut;
}

/* mergeDiagonalKeepScoredHitsDuplicates (CacheFriendlyOperations.cpp:116-145): walks every bin BACKWARDS; an element survives
 * when it already carries a score or its 8-bit diagonal differs from the element after it (same target); output order is the
 * reversed walk. */
static size_t cfo_merge_keep_scored(cfo_t *c, cand_t *arr, size_t N) {
    cfo_reserve(c, N);
    memcpy(c->tmp, arr, N * sizeof(cand_t));
    size_t nOut = 0;
    for (uint32_t bin = 0; bin <= c->binMask; bin++) {
        for (size_t n = 0; n < N; n++)
            if ((c->tmp[n].id & c->binMask) == bin) c->last[c->tmp[n].id >> c->binBits] = (unsigned char)((unsigned char)c->tmp[n].diag + 1);
        for (size_t n = N; n-- > 0;) {
            if ((c->tmp[n].id & c->binMask) != bin) continue;
            uint32_t slot = c->tmp[n].id >> c->binBits;
            unsigned char d8 = (unsigned char)c->tmp[n].diag;
            arr[nOut] = c->tmp[n];
            if (arr[nOut].count != 0 || c->last[slot] != d8) nOut++;
            c->last[slot] = d8;
        }
        for (size_t n = 0; n < N; n++)
            if ((c->tmp[n].id & c->binMask) == bin) c->last[c->tmp[n].id >> c->binBits] = 0;
    }
    return nOut;
}

/* keepMaxScoreElementOnly (CacheFriendlyOperations.cpp:350-380): per target the first element carrying the maximal score;
 * output again bin-major, arrival order inside a bin. */
static size_t cfo_keep_max(cfo_t *c, cand_t *arr, size_t N) {
    cfo_reserve(c, N);
    memcpy(c->tmp, arr, N * sizeof(cand_t));
    size_t nKeep = 0;
    for (uint32_t bin = 0; bin <= c->binMask; bin++) {
        for (size_t i = 0; i < N; i++) {
            if ((c->tmp[i].id & c->binMask) != bin) continue;
            uint32_t slot = c->tmp[i].id >> c->binBits;
            if (c->tmp[i].count > c->last[slot]) c->last[slot] = c->tmp[i].count;
        }
        for (size_t i = 0; i < N; i++) {
            if ((c->tmp[i].id & c->binMask) != bin) continue;
            uint32_t slot = c->tmp[i].id >> c->binBits;
            int found = (c->last[slot] == c->tmp[i].count);
            if (found) arr[nKeep++] = c->tmp[i];
            c->last[slot] = (unsigned char)(c->last[slot] * (1 - found));
        }
        for (size_t i = 0; i < N; i++)
            if ((c->tmp[i].id & c->binMask) == bin) c->last[c->tmp[i].id >> c->binBits] = 0;
    }
    return nKeep;
}

/* UngappedAlignment::align -> computeScores (UngappedAlignment.cpp:25-27, 331-362): elements that already carry a score are
 * skipped; the others get min(255, diagonal score) */
static void align_cands(cand_t *arr, size_t N, const signed char *prof, int L, const unsigned char *tres, const uint64_t *tOff,
                        uint64_t *residues) {
    for (size_t c = 0; c < N; c++) {
        if (arr[c].count != 0) continue;
        uint32_t t = arr[c].id;
        int s = diag_score(prof, L, tres + tOff[t], (int)(tOff[t + 1] - tOff[t]), (short)arr[c].diag, residues);
        arr[c].count = (unsigned char)(s > 255 ? 255 : s);
    }
}

/* QueryMatcher::match (QueryMatcher.cpp:213-346): index entries are buffered in databaseHits (maxDbMatches entries).  When the
 * next bucket would reach the end of the buffer (:280) the buffered part is reduced by findDuplicates (fresh table) into
 * foundDiagonals behind the results of earlier overflows; from the second overflow on the list is merged with
 * mergeDiagonalKeepScoredHitsDuplicates, scored and cut to the best diagonal per target (:296-303). */
typedef struct {
    hit_t *hits;
    size_t nHit, maxDbMatches;
    size_t overflowHitCount, overflowNumMatches, nOverflows;
    cand_t *found;
    size_t foundSize;
    cfo_t *cfo;
    const signed char *prof;
    int L;
    const unsigned char *tres;
    const uint64_t *tOff;
    uint64_t *residues;
} match_t;

/* returns 1 for the reference's `goto outer` (a single bucket as large as the whole buffer) */
static int match_bucket(match_t *m, const index_t *ix, uint32_t kmer, int p) {
    const uint32_t lo = ix->off[kmer], hi = ix->off[kmer + 1];
    const size_t len = hi - lo;
    if (m->nHit + len >= m->maxDbMatches) {
        const size_t hitCount = cfo_find_duplicates(m->cfo, m->hits, m->nHit, m->found + m->overflowHitCount,
                                                    m->foundSize - m->overflowHitCount);
        if (m->overflowHitCount != 0) {
            m->overflowHitCount = cfo_merge_keep_scored(m->cfo, m->found, hitCount + m->overflowHitCount);
            align_cands(m->found, m->overflowHitCount, m->prof, m->L, m->tres, m->tOff, m->residues);
            m->overflowHitCount = cfo_keep_max(m->cfo, m->found, m->overflowHitCount);
        } else m->overflowHitCount = hitCount;
        m->overflowNumMatches += m->nHit;
        m->nHit = 0;
        m->nOverflows++;
        if (len >= m->maxDbMatches) return 1;
    }
    for (uint32_t e = lo; e < hi; e++) {
        m->hits[m->nHit].id = ix->ent[e].seq;
        m->hits[m->nHit].diag = (uint16_t)(p - ix->ent[e].pos);
        m->nHit++;
    }
    return 0;
}

static long prefilter_core(const char *tSeq, const uint64_t *tOff, uint32_t nT,
                           const char *qSeq, const uint64_t *qOff, uint32_t nQ,
                           const orc_pref_params *par, orc_pref_hit *out, long outCap,
                           orc_pref_stats *stats, unsigned char *maskedOut, const unsigned char *profData);

long orc_prefilter(const char *tSeq, const uint64_t *tOff, uint32_t nT,
                   const char *qSeq, const uint64_t *qOff, uint32_t nQ,
                   const orc_pref_params *par, orc_pref_hit *out, long outCap,
                   orc_pref_stats *stats, unsigned char *maskedOut) {
    return prefilter_core(tSeq, tOff, nT, qSeq, qOff, nQ, par, out, outCap, stats, maskedOut, NULL);
}

/* profData: the profile entries concatenated, 25 bytes per column; qOff in COLUMNS */
long orc_prefilter_profile(const char *tSeq, const uint64_t *tOff, uint32_t nT,
                           const unsigned char *profData, const uint64_t *qOff, uint32_t nQ,
                           const orc_pref_params *par, orc_pref_hit *out, long outCap, orc_pref_stats *stats) {
    return prefilter_core(tSeq, tOff, nT, NULL, qOff, nQ, par, out, outCap, stats, NULL, profData);
}

static long prefilter_core(const char *tSeq, const uint64_t *tOff, uint32_t nT,
                           const char *qSeq, const uint64_t *qOff, uint32_t nQ,
                           const orc_pref_params *par, orc_pref_hit *out, long outCap,
                           orc_pref_stats *stats, unsigned char *maskedOut, const unsigned char *profData) {
    orc_submat kmat, umat;
    const int isProfile = profData != NULL;
    /* Prefiltering.cpp:67-77: sequence queries seed with VTML80 x8, profile queries take BLOSUM62 x8 (background distributions and
     * the masker's likelihood ratios then come from BLOSUM62, so the masked target residues differ between the two modes) */
    orc_build_submat(isProfile ? 0 : 1, 8.0f, -0.2f, &kmat);
    orc_build_submat(0, 2.0f, -0.2f, &umat);      /* Prefiltering.cpp:69 diagonal matrix BLOSUM62 x2 */
    const int kmerThr = isProfile ? orc_kmer_threshold_profile(par->sensitivity, KMER) : orc_kmer_threshold(par->sensitivity, KMER);
    orc_pref_stats st;
    memset(&st, 0, sizeof st);

    /* target residues: map + mask (IndexBuilder.cpp:117-133) */
    unsigned char *tres = (unsigned char *)malloc((size_t)tOff[nT] + 1);
    orc_map_sequence(tSeq, (int)tOff[nT], tres);
    if (par->mask)
        for (uint32_t t = 0; t < nT; t++)
            orc_tantan_mask(tres + tOff[t], (int)(tOff[t + 1] - tOff[t]), &kmat, (double)par->maskProb);
    if (maskedOut) memcpy(maskedOut, tres, (size_t)tOff[nT]);

    index_t ix;
    build_index(tres, tOff, nT, kmat.sub, isProfile ? 0 : kmerThr, &ix);
    ext3_t e3;
    e3.score = NULL; e3.index = NULL;
    if (!isProfile) build_ext3(kmat.sub, &e3);
    short *pkS[2] = {NULL, NULL};
    size_t *pkI[2] = {NULL, NULL};
    if (isProfile)
        for (int b = 0; b < 2; b++) {
            pkS[b] = (short *)malloc(sizeof(short) * ORC_MAX_KMER_RESULT);
            pkI[b] = (size_t *)malloc(sizeof(size_t) * ORC_MAX_KMER_RESULT);
        }

    int maxLq = 0;
    for (uint32_t q = 0; q < nQ; q++) {
        int L = (int)(qOff[q + 1] - qOff[q]);
        if (L > maxLq) maxLq = L;
    }
    unsigned char *qres = (unsigned char *)malloc((size_t)maxLq + 1);
    float *bias = (float *)malloc(sizeof(float) * (size_t)(maxLq + 1));
    signed char *prof = (signed char *)malloc((size_t)21 * (size_t)(maxLq + 1));
    /* QueryMatcher.cpp:42-43: hit buffer of maxDbMatches index entries, candidate array of foundDiagonalsSize elements
     * (par->maxDbMatches / par->foundDiagonalsSize: test hooks that shrink them; 0 = the reference's values) */
    const size_t refFound = (size_t)nT > 1000000 ? (size_t)nT : 1000000;
    const size_t maxDbMatches = par->maxDbMatches ? (size_t)par->maxDbMatches : refFound * 2;
    const size_t foundDiagonalsSize = par->foundDiagonalsSize ? (size_t)par->foundDiagonalsSize : refFound;
    size_t nCand;
    hit_t *hits = (hit_t *)malloc(sizeof(hit_t) * (maxDbMatches + 1));
    cand_t *cand = (cand_t *)calloc(foundDiagonalsSize + 1, sizeof(cand_t));     /* foundDiagonals */
    cfo_t cfo;
    cfo.tmpCap = 1 << 16;
    cfo.tmp = (cand_t *)malloc(sizeof(cand_t) * cfo.tmpCap);
    cfo.last = (unsigned char *)calloc((size_t)nT + 1, 1);
    const size_t maxHits = (size_t)par->maxSeqs < (size_t)nT ? (size_t)par->maxSeqs : (size_t)nT;
    res_t *res = (res_t *)malloc(sizeof(res_t) * (maxHits + 2));
    const uint32_t binMask = (uint32_t)par->binSize - 1;
    int binBits = 0;
    while ((1u << binBits) < (uint32_t)par->binSize) binBits++;
    cfo.binMask = binMask;
    cfo.binBits = binBits;
    long nOut = 0;

    for (uint32_t q = 0; q < nQ; q++) {
        const int L = (int)(qOff[q + 1] - qOff[q]);
        const unsigned char *pcol = isProfile ? profData + (size_t)qOff[q] * 25 : NULL;
        st.queryResidues += (uint64_t)L;
        if (isProfile) {
            /* Sequence::mapProfile: query letters, no composition bias; diagonal profile = score / 4, X column 0 */
            for (int p = 0; p < L; p++) {
                qres[p] = pcol[(size_t)p * 25 + 20];
                bias[p] = 0.0f;
                for (int a = 0; a < 20; a++) prof[p * 21 + a] = (signed char)((short)(signed char)pcol[(size_t)p * 25 + a] / 4);
                prof[p * 21 + 20] = 0;
            }
        } else {
        orc_map_sequence(qSeq + qOff[q], L, qres);
        /* QueryMatcher::matchQuery (QueryMatcher.cpp:85-107): bias from the SEED matrix */
        if (par->compBias) orc_comp_bias(&kmat, qres, L, bias, 1.0f);
        else memset(bias, 0, sizeof(float) * (size_t)L);
        /* UngappedAlignment::createProfile (UngappedAlignment.cpp:385-414) */
        for (int p = 0; p < L; p++) {
            float b = bias[p];
            b = (b < 0.0) ? b / 4 - 0.5 : b / 4 + 0.5;
            signed char corr = (signed char)b;
            for (int a = 0; a < 21; a++)
                prof[p * 21 + a] = (signed char)(umat.sub[qres[p] * 21 + a] + corr);
        }
        }
        /* QueryMatcher::match (QueryMatcher.cpp:213-346) + KmerGenerator (KmerGenerator.cpp:107-216) */
        match_t m;
        m.hits = hits; m.nHit = 0; m.maxDbMatches = maxDbMatches;
        m.overflowHitCount = 0; m.overflowNumMatches = 0; m.nOverflows = 0;
        m.found = cand; m.foundSize = foundDiagonalsSize; m.cfo = &cfo;
        m.prof = prof; m.L = L; m.tres = tres; m.tOff = tOff; m.residues = &st.diagResidues;
        int stop = 0;
        for (int p = 0; p + SPAN <= L && !stop; p++) {
            unsigned char km[KMER];
            int hasX = 0;
            float bsum = 0;
            for (int k = 0; k < KMER; k++) {
                km[k] = qres[p + SEED_POS[k]];
                if (km[k] == ORC_X) hasX = 1;
                bsum += bias[p + SEED_POS[k]];
            }
            if (hasX) continue;
            short b = (short)((bsum < 0.0) ? bsum - 0.5 : bsum + 0.5);
            short thr = (short)(kmerThr - b > 0 ? kmerThr - b : 0);
            if (isProfile) {
                short sc[KMER][20];
                unsigned int ixs[KMER][20];
                for (int k = 0; k < KMER; k++) {
                    const unsigned char *col = pcol + (size_t)(p + SEED_POS[k]) * 25;
                    for (int a = 0; a < 20; a++) { sc[k][a] = (short)(signed char)col[a]; ixs[k][a] = (unsigned int)a; }
                    ranked_desc_sort20(sc[k], ixs[k]);
                }
                const size_t *list;
                const size_t nList = profile_kmer_list(sc, ixs, thr, pkS, pkI, &list);
                for (size_t c = 0; c < nList && !stop; c++) {
                    st.probes++;
                    stop = match_bucket(&m, &ix, (uint32_t)list[c], p);
                }
                continue;
            }
            uint32_t i1 = km[0] + 20u * km[1] + 400u * km[2];
            uint32_t i2 = km[3] + 20u * km[4] + 400u * km[5];
            const short *s1 = e3.score + (size_t)i1 * N3, *s2 = e3.score + (size_t)i2 * N3;
            const unsigned short *x1 = e3.index + (size_t)i1 * N3, *x2 = e3.index + (size_t)i2 * N3;
            short cutoff1 = (short)(thr - s2[0]);
            for (int a = 0; a < N3 && s1[a] >= cutoff1 && !stop; a++) {
                short cutoff2 = (short)(thr - s1[a]);
                for (int c = 0; c < N3 && s2[c] >= cutoff2 && !stop; c++) {
                    uint32_t kmer = x1[a] + (uint32_t)N3 * x2[c];
                    st.probes++;
                    stop = match_bucket(&m, &ix, kmer, p);
                }
            }
        }
        /* QueryMatcher.cpp:322-333: the last buffered part; after an overflow the lists are merged by mergeDiagonalDuplicates.
         * An empty last part returns NO hits at all (hitCount stays 0) -- literal. */
        nCand = 0;
        if (m.nHit > 0) {
            nCand = cfo_find_duplicates(&cfo, hits, m.nHit, cand + m.overflowHitCount, foundDiagonalsSize - m.overflowHitCount);
            if (m.overflowHitCount != 0) nCand = cfo_merge_diagonal(&cfo, cand, m.overflowHitCount + nCand);
        }
        st.matches += m.overflowNumMatches + m.nHit;
        if (m.nOverflows) st.overflowQueries++;
        if (m.nOverflows > 1) st.multiOverflowQueries++;
        st.candidates += nCand;
        /* UngappedAlignment::align (UngappedAlignment.cpp:25-27, 331-362): score every candidate that has no score yet, cap 255 */
        align_cands(cand, nCand, prof, L, tres, tOff, &st.diagResidues);
        /* keepMaxScoreElementOnly (CacheFriendlyOperations.cpp:350-380) */
        size_t nKeep = cfo_keep_max(&cfo, cand, nCand);
        cfo_reserve(&cfo, nCand > nKeep ? nCand : nKeep);
        cand_t *tmp = cfo.tmp;
        memcpy(tmp, cand, nKeep * sizeof(cand_t));
        /* score histogram -> threshold (QueryMatcher.h:206-216), stable order by score desc
         * (radixSortByScoreSize QueryMatcher.cpp:498-523), getResult (QueryMatcher.cpp:363-420) */
        unsigned int histo[256];
        memset(histo, 0, sizeof histo);
        for (size_t c = 0; c < nKeep; c++) histo[tmp[c].count]++;
        size_t found = 0;
        unsigned int thr;
        for (thr = 255; thr > 0; thr--) {
            found += histo[thr];
            if (found >= maxHits) break;
        }
        if (thr < (unsigned int)par->minDiagScore) thr = (unsigned int)par->minDiagScore;
        size_t nAbove = 0;
        for (int s = 255; s >= (int)thr; s--)
            for (size_t c = 0; c < nKeep; c++)
                if (tmp[c].count == s) cand[nAbove++] = tmp[c];
        /* scoreIsTruncated (QueryMatcher.cpp:157-171): at least --max-seqs elements sit at the 8-bit cap, so the cap no longer
         * orders them.  rescoreHits (QueryMatcher.cpp:525-544) replaces the capped score by the uncapped diagonal score,
         * rescaled to 0..255 relative to the query's self score, then the list is re-sorted (stable) by that value and
         * getResult maps it back with rescaleScore = maxSelfScore - 255 (QueryMatcher.cpp:407-410). */
        int rescale = 0;
        if (thr >= 255) {
            int maxSelf = diag_score(prof, L, qres, L, 0, NULL) - 255;              /* scoreSingleSequence(query, 0, 0) */
            if (maxSelf < 1) maxSelf = 1;
            if (maxSelf > 65535) maxSelf = 65535;
            const float fltMaxSelf = (float)maxSelf;
            unsigned int histo2[256];
            memset(histo2, 0, sizeof histo2);
            size_t nNew = 0;
            for (size_t c = 0; c < nAbove && cand[c].count >= 255; c++) {
                uint32_t t = cand[c].id;
                unsigned int newScore = (unsigned int)diag_score(prof, L, tres + tOff[t], (int)(tOff[t + 1] - tOff[t]),
                                                                 (short)cand[c].diag, NULL);
                newScore -= 255u;
                const float score = (float)(newScore < 65535u ? newScore : 65535u);
                cand[c].count = (unsigned char)(int)((score / fltMaxSelf) * (float)255 + 0.5);
                histo2[cand[c].count]++;
                nNew++;
            }
            /* radixSortByScoreSize(threshold 0): stable, new score descending */
            size_t w = 0;
            for (int s = 255; s >= 0; s--)
                for (size_t c = 0; c < nNew; c++)
                    if (cand[c].count == s) tmp[w++] = cand[c];
            memcpy(cand, tmp, nNew * sizeof(cand_t));
            nAbove = nNew;
            rescale = maxSelf;
        }
        size_t nRes = 0;
        uint32_t identity = par->sameDB ? q : UINT_MAX;
        if (identity != UINT_MAX) { res[0].seqId = identity; res[0].score = 65535; res[0].diag = 0; nRes = 1; }
        size_t eligible = 0;
        for (size_t c = 0; c < nAbove; c++) {
            if (cand[c].id == identity) continue;
            eligible++;
            if (nRes >= maxHits) continue;
            res[nRes].seqId = cand[c].id;
            res[nRes].score = cand[c].count;
            res[nRes].diag = cand[c].diag;
            if (rescale != 0) {
                res[nRes].score = (int)(255u + ((unsigned int)cand[c].count * (unsigned int)rescale / 255u));
            } else if (cand[c].count >= 255) {
                uint32_t t = cand[c].id;
                res[nRes].score = diag_score(prof, L, tres + tOff[t], (int)(tOff[t + 1] - tOff[t]),
                                             (short)cand[c].diag, NULL);
            }
            nRes++;
        }
        if (eligible + (identity != UINT_MAX ? 1 : 0) > nRes) st.truncated++;
        if (nRes > 1) {
            if (identity != UINT_MAX) qsort(res + 1, nRes - 1, sizeof(res_t), cmp_res);
            else qsort(res, nRes, sizeof(res_t), cmp_res);
        }
        st.hits += nRes;
        /* Prefiltering::runSplit (Prefiltering.cpp:845-869): coverage pre-filter, then write */
        for (size_t r = 0; r < nRes; r++) {
            uint32_t t = res[r].seqId;
            if (par->covThr > 0.0 && (par->covMode == 0 || par->covMode == 2 || par->covMode == 5)) {
                if (!can_be_covered(par->covThr, par->covMode, (float)L, (float)(tOff[t + 1] - tOff[t]))) continue;
            }
            if (nOut >= outCap) return -1;
            out[nOut].query = q;
            out[nOut].target = t;
            out[nOut].score = res[r].score;
            out[nOut].diag = (int16_t)res[r].diag;
            nOut++;
        }
    }
    if (stats) *stats = st;
    free(tres); free(ix.off); free(ix.ent); free(e3.score); free(e3.index);
    free(qres); free(bias); free(prof); free(hits); free(cand); free(cfo.tmp); free(cfo.last); free(res);
    for (int b = 0; b < 2; b++) { free(pkS[b]); free(pkI[b]); }
    return nOut;
}
