/* oracle (test infrastructure): scalar restatement of MMseqs2 `align` as Spacedust's clustersearch
 * runs it (-a 1 --alignment-mode 2, BLOSUM62 x2, gap open 11 / extend 1, composition bias on).
 * See orc_common.h for the rules that govern this directory. */
#include "orc_align.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------------------------------------
 * E-value: EvalueComputation (alignment/EvalueComputation.h:18-92) with the hard-coded
 * BLOSUM62 (11,1) Gumbel parameters (:64-69) and ALP's finite-size-corrected area
 * (lib/alp/sls_pvalues.cpp:366-532, compute_tmp_values :340-361; Phi(x) = 0.5 erfc(-sqrt(.5) x)
 * lib/alp/sls_basic.hpp:195-198).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    double lambda, K, a_I, a_J, b_I, b_J, alpha_I, alpha_J, beta_I, beta_J, sigma, tau;
    double vi_thr, vj_thr, c_thr, logK;
} gumbel_t;

static void gumbel_blosum62_11_1(gumbel_t *g) {
    /* {lambda, K, a1, b1, a2, b2, alpha1, beta1, alpha2, beta2, sigma, tau} */
    const double lambda = 0.27359865037097330642, K = 0.044620920658722244834;
    const double a1 = 1.5938724404943873658, b1 = -19.959867650284412122;
    const double a2 = 1.5938724404943873658, b2 = -19.959867650284412122;
    const double alpha1 = 30.455610143099914211, beta1 = -622.28684628915891608;
    const double alpha2 = 30.455610143099914211, beta2 = -622.28684628915891608;
    const double sigma = 29.602444874818868215, tau = -601.81087985041381216;
    g->lambda = lambda; g->K = K;
    g->a_I = a2; g->a_J = a1; g->b_I = b2; g->b_J = b1;            /* sls_alignment_evaluer.cpp:679-721 */
    g->alpha_I = alpha2; g->alpha_J = alpha1; g->beta_I = beta2; g->beta_J = beta1;
    g->sigma = sigma; g->tau = tau;
    double v;
    v = 2.0 * g->alpha_I / g->lambda; g->vi_thr = v > 0.0 ? v : 0.0;
    v = 2.0 * g->alpha_J / g->lambda; g->vj_thr = v > 0.0 ? v : 0.0;
    v = 2.0 * g->sigma / g->lambda;   g->c_thr = v > 0.0 ? v : 0.0;
    g->logK = log(K);
}

static double phi(double x) { return 0.5 * erfc(-sqrt(0.5) * x); }

/* area(score, seqlen1 = query length, seqlen2 = DB residues): m_ = seqlen2, n_ = seqlen1 */
static double alp_area(const gumbel_t *g, double y, double seqlen1, double seqlen2) {
    const double pi = 3.1415926535897932384626433832795;
    const double const_val = 1 / sqrt(2.0 * pi);
    double m_ = seqlen2, n_ = seqlen1;
    double m_li_y = m_ - (g->a_I * y + g->b_I);
    double vi_y = g->alpha_I * y + g->beta_I;
    if (vi_y < g->vi_thr) vi_y = g->vi_thr;
    double sqrt_vi_y = sqrt(vi_y);
    double m_F = (sqrt_vi_y == 0.0) ? 1e100 : m_li_y / sqrt_vi_y;
    double P_m_F = phi(m_F);
    double E_m_F = -const_val * exp(-0.5 * m_F * m_F);
    double p1 = m_li_y * P_m_F - sqrt_vi_y * E_m_F;
    double n_lj_y = n_ - (g->a_J * y + g->b_J);
    double vj_y = g->alpha_J * y + g->beta_J;
    if (vj_y < g->vj_thr) vj_y = g->vj_thr;
    double sqrt_vj_y = sqrt(vj_y);
    double n_F = (sqrt_vj_y == 0.0) ? 1e100 : n_lj_y / sqrt_vj_y;
    double P_n_F = phi(n_F);
    double E_n_F = -const_val * exp(-0.5 * n_F * n_F);
    double p2 = n_lj_y * P_n_F - sqrt_vj_y * E_n_F;
    double c_y = g->sigma * y + g->tau;
    if (c_y < g->c_thr) c_y = g->c_thr;
    return p1 * p2 + c_y * (P_m_F * P_n_F);
}

double orc_evalue(double score, double qLen, double dbResidues) {
    gumbel_t g;
    gumbel_blosum62_11_1(&g);
    double epa = g.K * exp(-g.lambda * score);
    return epa * alp_area(&g, score, qLen, dbResidues);
}

int orc_bitscore(double score) {
    gumbel_t g;
    gumbel_blosum62_11_1(&g);
    return (int)((g.lambda * score - g.logK) / log(2.0) + 0.5);
}

/* ---------------------------------------------------------------------------------------------
 * Striped Smith-Waterman, restated row-sequentially.  sw_sse2_byte / sw_sse2_word
 * (alignment/StripedSmithWaterman.cpp:639-940, 943-1214) hold the query in `width` interleaved
 * segments of segLen rows.  Their result equals exact Gotoh H except for one layout-dependent rule:
 * the E (gap in target direction) update sees H before the lazy-F correction, i.e. with vertical
 * gaps that start inside the same segment only.  Both tracks are carried here: Ft/H (true vertical
 * chain) and Fm/H1 (chain reset at every segment start), E is fed from H1.
 * ------------------------------------------------------------------------------------------- */
typedef struct { int score, tEnd, qEnd, overflow; uint64_t cells; } sw_end;

static sw_end sw_striped(const signed char *rowScore /* [qLen][21]: score of query row vs target letter */,
                         int qLen, const unsigned char *t, int tLen, int reverseT, int width,
                         int go, int ge, int limit /* first H >= limit overflows (byte mode) */,
                         int terminate /* stop after the first column whose max equals this; <0: never */,
                         int *H, int *E) {
    sw_end r;
    r.score = 0; r.tEnd = -1; r.qEnd = 0x7fffffff; r.overflow = 0; r.cells = 0;
    const int segLen = (qLen + width - 1) / width;
    for (int i = 0; i < qLen; i++) { H[i] = 0; E[i] = 0; }
    int max = 0;
    for (int c = 0; c < tLen; c++) {
        const int col = reverseT ? (tLen - 1 - c) : c;
        const int letter = t[col];
        int diag = 0, Ft = 0, Fm = 0, colMax = 0, colMaxRow = -1;
        for (int i = 0; i < qLen; i++) {
            if (i % segLen == 0) Fm = 0;
            int h = diag + rowScore[i * 21 + letter];
            const int e = E[i];
            if (h < e) h = e;                 /* e >= 0 doubles as the local-alignment floor */
            int h1 = h > Fm ? h : Fm;
            int hs = h > Ft ? h : Ft;
            diag = H[i];
            H[i] = hs;
            if (hs > colMax) { colMax = hs; colMaxRow = i; }
            int open1 = h1 - go; if (open1 < 0) open1 = 0;
            int ext = e - ge; if (ext < 0) ext = 0;
            E[i] = ext > open1 ? ext : open1;
            int fm = Fm - ge; if (fm < 0) fm = 0;
            Fm = fm > open1 ? fm : open1;
            int openS = hs - go; if (openS < 0) openS = 0;
            int ft = Ft - ge; if (ft < 0) ft = 0;
            Ft = ft > openS ? ft : openS;
        }
        r.cells += (uint64_t)qLen;
        if (colMax > max) {
            max = colMax;
            if (max >= limit) { r.overflow = 1; break; }
            r.tEnd = col;
            r.qEnd = colMaxRow;              /* smallest row holding the column maximum */
        }
        if (colMax == terminate) break;
    }
    r.score = r.overflow ? 255 : max;
    return r;
}

/* banded_sw + traceback (StripedSmithWaterman.cpp:1348-1600): plain Gotoh on the sub-rectangle
 * inside a band that doubles until the score is reproduced; direction bytes as in the reference
 * (1 diag; 2/3 query-gap extend/open; 4/5 target-gap extend/open); ties: extend over open, diagonal
 * over gaps, F over E.  Returns backtrace length written to bt (forward order) or -1. */
static int banded_traceback(const signed char *rowScore, int qStart, const unsigned char *t, int dbLen,
                            int qLen, int score, int go, int ge, char *bt, int btCap, uint64_t *cells) {
    int band = abs(dbLen - qLen) + 1;
    int max = 0;
    signed char *dir = NULL;
    int *hb = NULL, *eb = NULL, *hc = NULL;
    long widthD = 0;
    for (;;) {
        const long width = (long)band * 2 + 3;
        widthD = (long)band * 2 + 1;
        free(dir); free(hb); free(eb); free(hc);
        dir = (signed char *)malloc((size_t)(widthD * qLen * 3 + 16));
        hb = (int *)calloc((size_t)width + 2, sizeof(int));
        eb = (int *)calloc((size_t)width + 2, sizeof(int));
        hc = (int *)calloc((size_t)width + 2, sizeof(int));
        for (int i = 0; i < qLen; i++) {
            int beg = i - band > 0 ? i - band : 0;
            int end = i + band < dbLen - 1 ? i + band : dbLen - 1;
            int edge = end + 1 < width - 1 ? end + 1 : (int)(width - 1);
            int f = 0, u = 0;
            hb[0] = eb[0] = hb[edge] = eb[edge] = hc[0] = 0;
            signed char *line = dir + widthD * i * 3;
            const int xi = i - band > 0 ? i - band : 0;          /* first column of this row's band */
            const int xim = i - 1 - band > 0 ? i - 1 - band : 0; /* ... of the previous row's band  */
            for (int j = beg; j <= end; j++) {
                u = j - xi + 1;
                const int up = j - xim + 1;      /* (i-1, j)   in h_b/e_b coordinates */
                const int left = j - 1 - xi + 1; /* (i, j-1)   in h_c coordinates     */
                const int dg = j - 1 - xim + 1;  /* (i-1, j-1)                        */
                const int cell = (j - xi) * 3;
                int t1 = i == 0 ? -go : hb[up] - go;
                int t2 = i == 0 ? -ge : eb[up] - ge;
                eb[u] = t1 > t2 ? t1 : t2;
                line[cell + 0] = t1 > t2 ? 3 : 2;
                t1 = hc[left] - go;
                t2 = f - ge;
                f = t1 > t2 ? t1 : t2;
                line[cell + 1] = t1 > t2 ? 5 : 4;
                const int f1 = f > 0 ? f : 0;
                const int e1 = eb[u] > 0 ? eb[u] : 0;
                t1 = e1 > f1 ? e1 : f1;
                t2 = hb[dg] + rowScore[(qStart + i) * 21 + t[j]];
                hc[u] = t1 > t2 ? t1 : t2;
                if (hc[u] > max) max = hc[u];
                if (t1 <= t2) line[cell + 2] = 1;
                else line[cell + 2] = e1 > f1 ? line[cell + 0] : line[cell + 1];
                if (cells) (*cells)++;
            }
            for (int j = 1; j <= u; j++) hb[j] = hc[j];
        }
        if (max >= score) break;
        band *= 2;
        if (band > 4 * (qLen + dbLen) + 8) break;   /* cannot happen for consistent inputs */
    }
    /* trace from the lower-right corner in state H */
    int i = qLen - 1, j = dbLen - 1, state = 2, n = 0, ok = 1;
    char *rev = (char *)malloc((size_t)(qLen + dbLen + 4));
    while (i > 0 || j > 0) {
        if (i < 0 || j < 0) { ok = 0; break; }
        const int xi = i - band > 0 ? i - band : 0;
        if (j - xi < 0 || j - xi >= widthD) { ok = 0; break; }
        const signed char d = dir[widthD * i * 3 + (j - xi) * 3 + state];
        char op;
        switch (d) {
            case 1: --i; --j; state = 2; op = 'M'; break;
            case 2: --i; state = 0; op = 'I'; break;
            case 3: --i; state = 2; op = 'I'; break;
            case 4: --j; state = 1; op = 'D'; break;
            case 5: --j; state = 2; op = 'D'; break;
            default: ok = 0; op = 'M'; break;
        }
        if (!ok) break;
        rev[n++] = op;
    }
    rev[n++] = 'M';        /* the start cell */
    int len = -1;
    if (ok && n <= btCap) {
        for (int k = 0; k < n; k++) bt[k] = rev[n - 1 - k];
        len = n;
    }
    free(rev); free(dir); free(hb); free(eb); free(hc);
    return len;
}

static float cov(unsigned int start, unsigned int end, unsigned int len) {   /* SmithWaterman::computeCov :1671-1673 */
    unsigned int mx = start > end ? start : end, mn = start < end ? start : end;
    unsigned int a = len < mx ? len : mx;
    return (a - mn + 1) / (float)len;
}

static int has_coverage(float covThr, int covMode, float q, float t) {   /* Util::hasCoverage Util.cpp:496-511 */
    switch (covMode) {
        case 0: return (q >= covThr) && (t >= covThr);
        case 2: return q >= covThr;
        case 1: return t >= covThr;
        default: return 1;
    }
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

static int cmp_result(const void *a, const void *b) {   /* Matcher::compareHits Matcher.h:157-168 */
    const orc_aln *x = (const orc_aln *)a, *y = (const orc_aln *)b;
    if (x->evalue != y->evalue) return x->evalue < y->evalue ? -1 : 1;
    if (x->bits != y->bits) return x->bits > y->bits ? -1 : 1;
    if (x->tLen != y->tLen) return x->tLen < y->tLen ? -1 : 1;
    if (x->targetKey != y->targetKey) return x->targetKey < y->targetKey ? -1 : 1;
    return 0;
}

static long align_core(const char *tSeq, const uint64_t *tOff, const uint32_t *tKeys, uint32_t nT,
                       const char *qSeq, const uint64_t *qOff, const uint32_t *qKeys, uint32_t nQ,
                       const uint32_t *pairQ, const uint32_t *pairT, uint64_t nPairs, uint64_t dbResidues,
                       const orc_align_params *par, orc_aln *out, long outCap, char *btPool, uint64_t btCap,
                       orc_align_stats *stats, const unsigned char *profData);

long orc_align(const char *tSeq, const uint64_t *tOff, const uint32_t *tKeys, uint32_t nT,
               const char *qSeq, const uint64_t *qOff, const uint32_t *qKeys, uint32_t nQ,
               const uint32_t *pairQ, const uint32_t *pairT, uint64_t nPairs, uint64_t dbResidues,
               const orc_align_params *par, orc_aln *out, long outCap, char *btPool, uint64_t btCap,
               orc_align_stats *stats) {
    return align_core(tSeq, tOff, tKeys, nT, qSeq, qOff, qKeys, nQ, pairQ, pairT, nPairs, dbResidues, par, out, outCap, btPool, btCap,
                      stats, NULL);
}

/* Profile queries (row a22): profData = profile entries concatenated, 25 bytes per column (Sequence.h:458-471), qOff in columns.
 * ssw_init for profiles (StripedSmithWaterman.cpp:1229-1282): score(row i, letter a) = profile score / 4 (Sequence.cpp:270-274),
 * X scores 0, no composition bias, bias = |min over the 20 x L scores|; identities are counted against the profile's query
 * letter (computerBacktrace :548-581). */
long orc_align_profile(const char *tSeq, const uint64_t *tOff, const uint32_t *tKeys, uint32_t nT,
                       const unsigned char *profData, const uint64_t *qOff, const uint32_t *qKeys, uint32_t nQ,
                       const uint32_t *pairQ, const uint32_t *pairT, uint64_t nPairs, uint64_t dbResidues,
                       const orc_align_params *par, orc_aln *out, long outCap, char *btPool, uint64_t btCap,
                       orc_align_stats *stats) {
    return align_core(tSeq, tOff, tKeys, nT, NULL, qOff, qKeys, nQ, pairQ, pairT, nPairs, dbResidues, par, out, outCap, btPool, btCap,
                      stats, profData);
}

static long align_core(const char *tSeq, const uint64_t *tOff, const uint32_t *tKeys, uint32_t nT,
                       const char *qSeq, const uint64_t *qOff, const uint32_t *qKeys, uint32_t nQ,
                       const uint32_t *pairQ, const uint32_t *pairT, uint64_t nPairs, uint64_t dbResidues,
                       const orc_align_params *par, orc_aln *out, long outCap, char *btPool, uint64_t btCap,
                       orc_align_stats *stats, const unsigned char *profData) {
    orc_submat m;
    orc_build_submat(0, 2.0f, par->scoreBias, &m);      /* Alignment.cpp:152 */
    signed char mat[21 * 21];
    int minMat = 0;
    for (int i = 0; i < 21 * 21; i++) { mat[i] = (signed char)m.sub[i]; if (mat[i] < minMat) minMat = mat[i]; }
    orc_align_stats st;
    memset(&st, 0, sizeof st);
    int maxL = 1;
    for (uint32_t q = 0; q < nQ; q++) if ((int)(qOff[q + 1] - qOff[q]) > maxL) maxL = (int)(qOff[q + 1] - qOff[q]);
    for (uint32_t t = 0; t < nT; t++) if ((int)(tOff[t + 1] - tOff[t]) > maxL) maxL = (int)(tOff[t + 1] - tOff[t]);
    unsigned char *qres = (unsigned char *)malloc((size_t)maxL + 1), *tres = (unsigned char *)malloc((size_t)maxL + 1);
    float *biasF = (float *)malloc(sizeof(float) * (size_t)(maxL + 1));
    signed char *cb = (signed char *)malloc((size_t)maxL + 1);
    signed char *rowScore = (signed char *)malloc((size_t)21 * (size_t)(maxL + 1));
    signed char *revScore = (signed char *)malloc((size_t)21 * (size_t)(maxL + 1));
    int *H = (int *)malloc(sizeof(int) * (size_t)(maxL + 1)), *E = (int *)malloc(sizeof(int) * (size_t)(maxL + 1));
    char *bt = (char *)malloc((size_t)2 * maxL + 8);
    long nOut = 0;
    uint64_t btUsed = 0;
    uint64_t p = 0;
    while (p < nPairs) {
        const uint32_t q = pairQ[p];
        uint64_t pe = p;
        while (pe < nPairs && pairQ[pe] == q) pe++;
        const int Lq = (int)(qOff[q + 1] - qOff[q]);
        int bias;
        if (profData) {
            const unsigned char *pcol = profData + (size_t)qOff[q] * 25;
            int minProf = 0;
            for (int i = 0; i < Lq; i++) {
                qres[i] = pcol[(size_t)i * 25 + 20];
                for (int a = 0; a < 20; a++) {
                    const signed char v = (signed char)((short)(signed char)pcol[(size_t)i * 25 + a] / 4);
                    rowScore[i * 21 + a] = v;
                    if (v < minProf) minProf = v;
                }
                rowScore[i * 21 + 20] = 0;
            }
            bias = abs(minProf);
        } else {
        orc_map_sequence(qSeq + qOff[q], Lq, qres);
        /* SmithWaterman::ssw_init (StripedSmithWaterman.cpp:1216-1345) */
        int minCb = 0;
        if (par->compBias) {
            orc_comp_bias(&m, qres, Lq, biasF, 1.0f);
            for (int i = 0; i < Lq; i++) {
                cb[i] = (signed char)((biasF[i] < 0.0) ? biasF[i] - 0.5 : biasF[i] + 0.5);
                if (cb[i] < minCb) minCb = cb[i];
            }
        } else memset(cb, 0, (size_t)Lq);
        bias = abs(minMat) + abs(minCb);
        for (int i = 0; i < Lq; i++)
            for (int a = 0; a < 21; a++) rowScore[i * 21 + a] = (signed char)(mat[a * 21 + qres[i]] + cb[i]);
        }
        const long firstOut = nOut;
        for (uint64_t k = p; k < pe; k++) {
            const uint32_t t = pairT[k];
            const int Lt = (int)(tOff[t + 1] - tOff[t]);
            if (!can_be_covered(par->covThr, par->covMode, (float)Lq, (float)Lt)) continue;      /* Alignment.cpp:370 */
            orc_map_sequence(tSeq + tOff[t], Lt, tres);
            st.alignments++;
            orc_aln r;
            memset(&r, 0, sizeof r);
            r.query = q; r.target = t; r.targetKey = tKeys[t]; r.qLen = Lq; r.tLen = Lt;
            const int identity = par->sameDB && qKeys[q] == tKeys[t];
            int pass = 0;
            int btLen = 0;
            if (identity) {
                /* SmithWaterman::scoreIdentical (:1675-1710); note the 16-bit accumulator */
                short sc = 0;
                for (int i = 0; i < Lt && i < Lq; i++) sc = (short)(sc + rowScore[i * 21 + tres[i]]);
                r.score = sc;
                r.evalue = orc_evalue(r.score, Lq, (double)dbResidues);
                r.qStart = 0; r.tStart = 0; r.qEnd = Lt - 1; r.tEnd = Lt - 1;
                r.qcov = 1.0f; r.tcov = 1.0f; r.seqId = 1.0f;
                btLen = Lt;
                for (int i = 0; i < Lt; i++) bt[i] = 'M';
                r.alnLen = Lt;
                pass = 1;
            } else {
                sw_end f = sw_striped(rowScore, Lq, tres, Lt, 0, 32, par->gapOpen, par->gapExtend, 255 - bias, 255, H, E);
                st.cellsForward += f.cells;
                int word = 0;
                if (f.overflow) {
                    f = sw_striped(rowScore, Lq, tres, Lt, 0, 16, par->gapOpen, par->gapExtend, 0x7fffffff, 65535, H, E);
                    st.cellsForward += f.cells;
                    st.wordMode++;
                    word = 1;
                }
                r.score = f.score; r.qEnd = f.qEnd; r.tEnd = f.tEnd; r.qStart = -1; r.tStart = -1;
                if (f.tEnd < 0) { r.evalue = 0; continue; }       /* no residue could be aligned: never passes */
                r.evalue = orc_evalue(r.score, Lq, (double)dbResidues);
                r.qcov = cov(0, (unsigned)r.qEnd, (unsigned)Lq);
                r.tcov = cov(0, (unsigned)r.tEnd, (unsigned)Lt);
                if (r.evalue > par->evalThr || !has_coverage(par->covThr, par->covMode, r.qcov, r.tcov)) continue;
                /* reverse pass on the prefixes, stops at the first column whose maximum is the score */
                const int qn = r.qEnd + 1, tn = r.tEnd + 1;
                for (int i = 0; i < qn; i++) memcpy(revScore + i * 21, rowScore + (r.qEnd - i) * 21, 21);
                sw_end b = sw_striped(revScore, qn, tres, tn, 1, word ? 16 : 32, par->gapOpen, par->gapExtend,
                                      word ? 0x7fffffff : 255 - bias, r.score, H, E);
                st.cellsReverse += b.cells;
                st.reversePasses++;
                if (b.score != r.score) { st.scoreMismatch++; continue; }   /* the reference aborts here */
                r.tStart = b.tEnd;
                r.qStart = r.qEnd - b.qEnd;
                r.qcov = cov((unsigned)r.qStart, (unsigned)r.qEnd, (unsigned)Lq);
                r.tcov = cov((unsigned)r.tStart, (unsigned)r.tEnd, (unsigned)Lt);
                if (!has_coverage(par->covThr, par->covMode, r.qcov, r.tcov)) continue;
                btLen = banded_traceback(rowScore, r.qStart, tres + r.tStart, r.tEnd - r.tStart + 1,
                                         r.qEnd - r.qStart + 1, r.score, par->gapOpen, par->gapExtend, bt, 2 * maxL + 4,
                                         &st.cellsTraceback);
                st.tracebacks++;
                if (btLen < 0) { st.tracebackFailed++; continue; }
                int ids = 0, qi = r.qStart, ti = r.tStart;
                for (int k2 = 0; k2 < btLen; k2++) {
                    if (bt[k2] == 'M') { ids += (tres[ti] == qres[qi]); qi++; ti++; }
                    else if (bt[k2] == 'I') qi++;
                    else ti++;
                }
                r.alnLen = btLen;
                r.seqId = (float)ids / (float)btLen;             /* Util::computeSeqId, SEQ_ID_ALN_LEN */
                /* Alignment::checkCriteria (Alignment.cpp:548-567) */
                pass = (r.evalue <= par->evalThr) && (r.seqId >= par->seqIdThr) &&
                       has_coverage(par->covThr, par->covMode, r.qcov, r.tcov) && (r.alnLen >= par->minAlnLen);
            }
            if (!pass) continue;
            r.bits = orc_bitscore(r.score);
            if (nOut >= outCap || btUsed + (uint64_t)btLen > btCap) return -1;
            r.btOff = btUsed; r.btLen = (uint32_t)btLen;
            memcpy(btPool + btUsed, bt, (size_t)btLen);
            btUsed += (uint64_t)btLen;
            out[nOut++] = r;
            st.passed++;
        }
        if (nOut - firstOut > 1) qsort(out + firstOut, (size_t)(nOut - firstOut), sizeof(orc_aln), cmp_result);
        p = pe;
    }
    if (stats) *stats = st;
    free(qres); free(tres); free(biasF); free(cb); free(rowScore); free(revScore); free(H); free(E); free(bt);
    return nOut;
}

/* Matcher::resultToBuffer (Matcher.cpp:280-327) with the run-length backtrace of compressAlignment (:166-185). */
int orc_format_alignment(const orc_aln *r, const char *btPool, char *buf, int cap) {
    int n = snprintf(buf, (size_t)cap, "%u\t%d\t", r->targetKey, r->bits);
    if (r->seqId == 1.0f) n += snprintf(buf + n, (size_t)(cap - n), "1.00");
    else {
        n += snprintf(buf + n, (size_t)(cap - n), "0.");
        if (r->seqId < 0.10) buf[n++] = '0';
        if (r->seqId < 0.01) buf[n++] = '0';
        n += snprintf(buf + n, (size_t)(cap - n), "%d", (int)(r->seqId * 1000));
    }
    n += snprintf(buf + n, (size_t)(cap - n), "\t%.3E\t%d\t%d\t%d\t%d\t%d\t%d\t", r->evalue, r->qStart, r->qEnd, r->qLen,
                  r->tStart, r->tEnd, r->tLen);
    const char *bt = btPool + r->btOff;
    char state = 'M';
    unsigned count = 0;
    for (uint32_t i = 0; i < r->btLen; i++) {
        if (bt[i] != state) {
            n += snprintf(buf + n, (size_t)(cap - n), "%u%c", count, state);
            state = bt[i];
            count = 1;
        } else count++;
    }
    n += snprintf(buf + n, (size_t)(cap - n), "%u%c\n", count, state);
    return n;
}
