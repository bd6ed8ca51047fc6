/* oracle (test infrastructure): substitution matrices, residue mapping, composition bias, tantan
 * masking.  See orc_common.h for the rules that govern this directory. */
#include "orc_common.h"
#include "matrix_data.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>

/* Follows SubstitutionMatrix::readProbMatrix (lib/mmseqs/src/commons/SubstitutionMatrix.cpp:315-397)
 * and BaseMatrix::generateSubMatrix (lib/mmseqs/src/commons/BaseMatrix.cpp:98-159): the .out file
 * carries half-bit scores S, a background and lambda; joint P_ab = exp(lambda*S_ab) p_a p_b with the
 * 20 real letters' background scaled by (1 - p_X); the integer matrix is
 * round_half_away(bitFactor * log2(P_ab / (r_a r_b)) + bias) with r = row sums of P (r_X = 1e-5). */
void orc_build_submat(int which, float bitFactor, float scoreBias, orc_submat *out) {
    const double (*S)[21] = which == 0 ? sd_blosum62_scores : sd_vtml80_scores;
    const double *bg = which == 0 ? sd_blosum62_background : sd_vtml80_background;
    const double lambda = which == 0 ? sd_blosum62_lambda : sd_vtml80_lambda;
    double pBack[ORC_AA];
    for (int i = 0; i < ORC_AA; i++) pBack[i] = bg[i];
    /* X scores are all negative in both tables => "xIsPositive == false" branch */
    for (int i = 0; i < ORC_AA - 1; i++) pBack[i] = pBack[i] * (1.0 - pBack[ORC_X]);
    for (int i = 0; i < ORC_AA; i++)
        for (int j = 0; j < ORC_AA; j++)
            out->prob[i * ORC_AA + j] = exp(lambda * S[i][j]) * pBack[i] * pBack[j];
    for (int i = 0; i < ORC_AA; i++) out->pBack[i] = pBack[i];
    double r[ORC_AA];
    for (int i = 0; i < ORC_AA; i++) {
        r[i] = 0;
        for (int j = 0; j < ORC_AA; j++) r[i] += out->prob[i * ORC_AA + j];
    }
    r[ORC_AA - 1] = 1E-5;
    const double bf = (double)bitFactor, sb = (double)scoreBias;
    for (int i = 0; i < ORC_AA; i++)
        for (int j = 0; j < ORC_AA; j++) {
            double s = log2(out->prob[i * ORC_AA + j] / (r[i] * r[j]));
            double v = bf * s + sb;
            out->sub[i * ORC_AA + j] = (short)((v < 0.0) ? v - 0.5 : v + 0.5);
        }
}

/* Letter -> code, SubstitutionMatrix::setupLetterMapping (SubstitutionMatrix.cpp:257-298):
 * J->L, B->D, Z->E, U/O/anything else -> X; case-insensitive. */
static unsigned char g_aa2num[256];
static int g_aa2num_ready = 0;
static void init_aa2num(void) {
    const char *letters = sd_blosum62_letters;
    unsigned char base[256];
    memset(base, 255, sizeof base);
    for (int i = 0; i < ORC_AA; i++) base[(unsigned char)letters[i]] = (unsigned char)i;
    for (int c = 0; c < 256; c++) {
        int u = toupper(c);
        unsigned char code;
        if (u == 'J') code = base['L'];
        else if (u == 'Z') code = base['E'];
        else if (u == 'B') code = base['D'];
        else if (u < 256 && base[u] != 255) code = base[u];
        else code = ORC_X;
        g_aa2num[c] = code;
    }
    g_aa2num_ready = 1;
}

void orc_map_sequence(const char *s, int len, unsigned char *out) {
    if (!g_aa2num_ready) init_aa2num();
    for (int i = 0; i < len; i++) out[i] = g_aa2num[(unsigned char)s[i]];
}

/* SubstitutionMatrix::calcLocalAaBiasCorrection (SubstitutionMatrix.cpp:79-109): window +-20,
 * float accumulator updated through double intermediates exactly as the mixed-type expression does. */
void orc_comp_bias(const orc_submat *m, const unsigned char *seq, int L, float *bias, float scale) {
    for (int i = 0; i < L; i++) {
        int lo = i - 20 < 0 ? 0 : i - 20;
        int hi = i + 20 > L ? L : i + 20;
        int wlen = hi - lo;
        const short *row = m->sub + seq[i] * ORC_AA;
        int sum = 0;
        for (int j = lo; j < hi; j++) sum += row[seq[j]];
        sum -= row[seq[i]];
        float d = (float)sum;
        d = (float)((double)d / (-1.0 * (double)(float)wlen));
        for (int a = 0; a < ORC_AA; a++) d = (float)((double)d + m->pBack[a] * (double)(float)row[a]);
        bias[i] = scale * d;
    }
}

/* tantan repeat masker as MMseqs2 calls it (Masker.cpp:52-92 -> lib/tantan/tantan.cpp:calcRepeatProbs):
 * HMM with a background state and 50 repeat-offset states, no indel states; forward pass with
 * rescaling every 16 letters, backward pass, posterior >= minMaskProb => X.  The summation order
 * of the AVX2 build (4 interleaved partial sums, (s0+s2)+(s1+s3), then the scalar tail) is kept. */
int orc_tantan_mask(unsigned char *seq, int L, const orc_submat *m, double minMaskProb) {
    enum { W = 50 };
    if (L <= 0) return 0;
    double lr[ORC_AA][ORC_AA];
    for (int i = 0; i < ORC_AA; i++)
        for (int j = 0; j < ORC_AA; j++)
            lr[i][j] = m->prob[i * ORC_AA + j] / (m->pBack[i] * m->pBack[j]);
    const double repeatProb = 0.005, repeatEndProb = 0.05, decay = 0.9;
    const double b2b = 1 - repeatProb, f2b = repeatEndProb, f2f0 = 1 - repeatEndProb;
    const double b2fFirst = repeatProb * ((1 - decay) / (1 - pow(decay, W)));
    double b2f[W], fg[W];
    {
        double p = b2fFirst;
        for (int i = 0; i < W; i++) { b2f[i] = p; p *= decay; }
    }
    float *lp = (float *)malloc(sizeof(float) * (size_t)L);
    double *scales = (double *)malloc(sizeof(double) * (size_t)(L / 16 + 1));
    double bgp = 1.0;
    for (int i = 0; i < W; i++) fg[i] = 0.0;
    for (int p = 0; p < L; p++) {
        const double *row = lr[seq[p]];
        int maxOff = p < W ? p : W;
        double b = bgp, s[4] = {0, 0, 0, 0};
        int i = 0;
        for (; i <= maxOff - 4; i += 4)
            for (int k = 0; k < 4; k++) {
                double f = fg[i + k];
                s[k] = s[k] + f;
                fg[i + k] = (b * b2f[i + k] + f * f2f0) * row[seq[p - i - k - 1]];
            }
        double fromFg = (s[0] + s[2]) + (s[1] + s[3]);
        for (; i < maxOff; i++) {
            double f = fg[i];
            fromFg += f;
            fg[i] = (b * b2f[i] + f * f2f0) * row[seq[p - i - 1]];
        }
        bgp = b * b2b + fromFg * f2b;
        if (p % 16 == 15) {
            double sc = 1 / bgp;
            scales[p / 16] = sc;
            bgp *= sc;
            for (int k = 0; k < W; k++) fg[k] *= sc;
        }
        lp[p] = (float)bgp;
    }
    double tot = 0.0;
    for (int i = 0; i < W; i++) tot += fg[i];
    double z = bgp * b2b + tot * f2b;
    bgp = b2b;
    for (int i = 0; i < W; i++) fg[i] = f2b;
    for (int p = L - 1; p >= 0; p--) {
        double nonRepeat = lp[p] * bgp / z;
        lp[p] = 1 - (float)nonRepeat;
        if (p % 16 == 15) {
            double sc = scales[p / 16];
            bgp *= sc;
            for (int k = 0; k < W; k++) fg[k] *= sc;
        }
        const double *row = lr[seq[p]];
        int maxOff = p < W ? p : W;
        double toBg = f2b * bgp, s[4] = {0, 0, 0, 0};
        int i = 0;
        for (; i <= maxOff - 4; i += 4)
            for (int k = 0; k < 4; k++) {
                double f = fg[i + k] * row[seq[p - i - k - 1]];
                s[k] = s[k] + b2f[i + k] * f;
                fg[i + k] = toBg + f2f0 * f;
            }
        double toFg = (s[0] + s[2]) + (s[1] + s[3]);
        for (; i < maxOff; i++) {
            double f = fg[i] * row[seq[p - i - 1]];
            toFg += b2f[i] * f;
            fg[i] = toBg + f2f0 * f;
        }
        bgp = b2b * bgp + toFg;
    }
    int masked = 0;
    for (int p = 0; p < L; p++)
        if (lp[p] >= minMaskProb) { seq[p] = ORC_X; masked++; }
    free(lp);
    free(scales);
    return masked;
}

/* Prefiltering::getKmerThreshold (Prefiltering.cpp:1005-1065), sequence (non-profile) branch. */
int orc_kmer_threshold(float sensitivity, int kmerSize) {
    float thr;
    if (kmerSize == 5) { float base = 160.75; thr = base - (sensitivity * 12.75); }
    else if (kmerSize == 6) { float base = 163.2; thr = base - (sensitivity * 8.917); }
    else { float base = 186.15; thr = base - (sensitivity * 11.22); }
    return (int)thr;
}
