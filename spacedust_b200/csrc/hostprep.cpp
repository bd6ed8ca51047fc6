#include "hostprep.h"
#include "matrix_data.h"

#include <algorithm>
#include <atomic>
#include <cctype>
#include <cmath>
#include <cstring>
#include <thread>
#include <unistd.h>

namespace sd {

void buildSubMat(MatrixId which, float bitFactor, float scoreBias, SubMat &out) {
    const double (*S)[21] = which == BLOSUM62 ? sd_blosum62_scores : sd_vtml80_scores;
    const double *bg = which == BLOSUM62 ? sd_blosum62_background : sd_vtml80_background;
    const double lambda = which == BLOSUM62 ? sd_blosum62_lambda : sd_vtml80_lambda;
    // readProbMatrix (SubstitutionMatrix.cpp:315-397): background of the 20 letters is scaled by
    // (1 - p_X) because no X score is positive; joint P_ab = exp(lambda S_ab) p_a p_b
    for (int a = 0; a < AA; a++) out.pBack[a] = bg[a];
    for (int a = 0; a < AA - 1; a++) out.pBack[a] = out.pBack[a] * (1.0 - out.pBack[XCODE]);
    for (int a = 0; a < AA; a++)
        for (int b = 0; b < AA; b++)
            out.prob[a * AA + b] = std::exp(lambda * S[a][b]) * out.pBack[a] * out.pBack[b];
    // generateSubMatrix (BaseMatrix.cpp:110-159): marginals are re-derived as row sums (X pinned to
    // 1e-5), scores = round-half-away(bitFactor*log2(P_ab/(r_a r_b)) + bias)
    double marg[AA];
    for (int a = 0; a < AA; a++) {
        double s = 0;
        for (int b = 0; b < AA; b++) s += out.prob[a * AA + b];
        marg[a] = s;
    }
    marg[XCODE] = 1E-5;
    const double bf = bitFactor, sb = scoreBias;
    for (int a = 0; a < AA; a++)
        for (int b = 0; b < AA; b++) {
            const double v = bf * std::log2(out.prob[a * AA + b] / (marg[a] * marg[b])) + sb;
            out.sub[a * AA + b] = static_cast<short>((v < 0.0) ? v - 0.5 : v + 0.5);
        }
}

static uint8_t g_map[256];
static std::atomic<bool> g_mapReady{false};

const uint8_t *aa2num() {
    if (!g_mapReady.load(std::memory_order_acquire)) {
        uint8_t direct[256];
        std::memset(direct, 0xFF, sizeof direct);
        for (int i = 0; i < AA; i++) direct[static_cast<unsigned char>(sd_blosum62_letters[i])] = static_cast<uint8_t>(i);
        uint8_t tmp[256];
        for (int c = 0; c < 256; c++) {
            const int u = std::toupper(c);
            uint8_t code = XCODE;
            if (u == 'J') code = direct['L'];
            else if (u == 'B') code = direct['D'];
            else if (u == 'Z') code = direct['E'];
            else if (u >= 0 && u < 256 && direct[u] != 0xFF) code = direct[u];
            tmp[c] = code;
        }
        std::memcpy(g_map, tmp, sizeof tmp);
        g_mapReady.store(true, std::memory_order_release);
    }
    return g_map;
}

void mapSequence(const char *s, uint64_t len, uint8_t *out) {
    const uint8_t *m = aa2num();
    for (uint64_t i = 0; i < len; i++) out[i] = m[static_cast<unsigned char>(s[i])];
}

void likelihoodRatios(const SubMat &m, double *lr) {
    // ProbabilityMatrix ctor (lib/mmseqs/src/commons/BaseMatrix.h:83-96)
    for (int a = 0; a < AA; a++)
        for (int b = 0; b < AA; b++) lr[a * AA + b] = m.prob[a * AA + b] / (m.pBack[a] * m.pBack[b]);
}


int kmerThreshold(float sensitivity, int kmerSize) {
    float best;
    if (kmerSize == 5) { float base = 160.75; best = base - (sensitivity * 12.75); }
    else if (kmerSize == 6) { float base = 163.2; best = base - (sensitivity * 8.917); }
    else { float base = 186.15; best = base - (sensitivity * 11.22); }
    return static_cast<int>(best);
}

void buildExt3(const SubMat &seedMat, std::vector<int16_t> &score, std::vector<uint16_t> &index, std::vector<uint16_t> &cntGE) {
    constexpr int N3 = 8000;
    cntGE.resize(static_cast<size_t>(N3) * 512);
    score.resize(static_cast<size_t>(N3) * N3);
    index.resize(static_cast<size_t>(N3) * N3);
    // pair tables first: s2[x0,x1][y0,y1] would be 160k entries; the triple loop below is cheap enough
    unsigned nThreads = hostThreads();
    std::atomic<int> nextRow{0};
    auto work = [&]() {
        std::vector<int16_t> rowScore(N3);
        for (;;) {
            const int x = nextRow.fetch_add(16);
            if (x >= N3) break;
            for (int xr = x; xr < std::min(N3, x + 16); xr++) {
                const int x0 = xr % 20, x1 = (xr / 20) % 20, x2 = xr / 400;   // stored index = x0 + 20 x1 + 400 x2
                const short *r0 = seedMat.sub + x0 * AA, *r1 = seedMat.sub + x1 * AA, *r2 = seedMat.sub + x2 * AA;
                int histo[512] = {0};
                int e = 0;
                for (int y0 = 0; y0 < 20; y0++)
                    for (int y1 = 0; y1 < 20; y1++) {
                        const int s01 = r0[y0] + r1[y1];
                        for (int y2 = 0; y2 < 20; y2++) {
                            const int16_t s = static_cast<int16_t>(s01 + r2[y2]);
                            rowScore[e++] = s;
                            histo[255 - s]++;
                        }
                    }
                int base = 0;
                uint16_t *outC = cntGE.data() + static_cast<size_t>(xr) * 512;    // outC[255 - v] = #neighbours with score >= v
                for (int h = 0; h < 512; h++) { const int c = histo[h]; histo[h] = base; base += c; outC[h] = static_cast<uint16_t>(base); }
                int16_t *outS = score.data() + static_cast<size_t>(xr) * N3;
                uint16_t *outI = index.data() + static_cast<size_t>(xr) * N3;
                e = 0;
                for (int y0 = 0; y0 < 20; y0++)
                    for (int y1 = 0; y1 < 20; y1++)
                        for (int y2 = 0; y2 < 20; y2++) {
                            const int16_t s = rowScore[e++];
                            const int slot = histo[255 - s]++;
                            outS[slot] = s;
                            outI[slot] = static_cast<uint16_t>(y0 + 20 * y1 + 400 * y2);
                        }
            }
        }
    };
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < nThreads; t++) pool.emplace_back(work);
    work();
    for (auto &t : pool) t.join();
}

int referenceBinSize(uint64_t nTargets) {
    uint64_t l2 = 262144;
#if defined(_SC_LEVEL2_CACHE_SIZE)
    const long v = sysconf(_SC_LEVEL2_CACHE_SIZE);
    if (v > 0) l2 = static_cast<uint64_t>(v);
#endif
    for (int b = 2; b <= 1024; b *= 2)
        if (nTargets / b < l2) return b;
    return 2048;
}

Evaluer::Evaluer(double dbRes) : dbResidues(dbRes) {
    // {lambda, K, a1, b1, a2, b2, alpha1, beta1, alpha2, beta2, sigma, tau} for blosum62.out, gap open 11, extend 1
    lambda = 0.27359865037097330642; K = 0.044620920658722244834;
    const double a1 = 1.5938724404943873658, b1 = -19.959867650284412122, a2 = 1.5938724404943873658, b2 = -19.959867650284412122;
    const double alpha1 = 30.455610143099914211, beta1 = -622.28684628915891608, alpha2 = 30.455610143099914211, beta2 = -622.28684628915891608;
    sigma = 29.602444874818868215; tau = -601.81087985041381216;
    aI = a2; aJ = a1; bI = b2; bJ = b1; alphaI = alpha2; alphaJ = alpha1; betaI = beta2; betaJ = beta1;
    viThr = std::max(2.0 * alphaI / lambda, 0.0);
    vjThr = std::max(2.0 * alphaJ / lambda, 0.0);
    cThr = std::max(2.0 * sigma / lambda, 0.0);
    logK = std::log(K);
}

static inline double normalCdf(double x) { return 0.5 * std::erfc(-std::sqrt(0.5) * x); }

double Evaluer::area(double y, double seqLen) const {
    const double pi = 3.1415926535897932384626433832795;
    const double invSqrt2Pi = 1 / std::sqrt(2.0 * pi);
    const double m = dbResidues, n = seqLen;       // AlignmentEvaluer::area passes (seqlen2_, seqlen1_)
    const double mLen = m - (aI * y + bI);
    const double vi = std::max(viThr, alphaI * y + betaI);
    const double sVi = std::sqrt(vi);
    const double mF = (sVi == 0.0) ? 1e100 : mLen / sVi;
    const double pM = normalCdf(mF);
    const double eM = -invSqrt2Pi * std::exp(-0.5 * mF * mF);
    const double p1 = mLen * pM - sVi * eM;
    const double nLen = n - (aJ * y + bJ);
    const double vj = std::max(vjThr, alphaJ * y + betaJ);
    const double sVj = std::sqrt(vj);
    const double nF = (sVj == 0.0) ? 1e100 : nLen / sVj;
    const double pN = normalCdf(nF);
    const double eN = -invSqrt2Pi * std::exp(-0.5 * nF * nF);
    const double p2 = nLen * pN - sVj * eN;
    const double c = std::max(cThr, sigma * y + tau);
    return p1 * p2 + c * (pM * pN);
}

double Evaluer::evalue(double score, double seqLen) const {
    const double perArea = K * std::exp(-lambda * score);
    return perArea * area(score, seqLen);
}

int Evaluer::bitScore(double score) const { return static_cast<int>((lambda * score - logK) / std::log(2.0) + 0.5); }

}  // namespace sd
