// Host-side preparation shared by the CUDA stages: substitution matrices, residue mapping,
// the sorted 3-mer neighbour table.  Small double-precision pieces the
// reference also runs once per process; they stay on the host in the reference's operation order
// (hard part 4).
#pragma once
#include <algorithm>
#include <cstdint>
#include <thread>
#include <vector>

namespace sd {

// Host threads of the post-processing helpers: all hardware threads (at most 64) for a single process; with one process per GPU
// (sd_shard_init) every rank takes its share of the node, otherwise N ranks x 64 threads fight for the same cores.
inline unsigned &hostThreadLimit() { static unsigned limit = 0; return limit; }
inline unsigned hostThreads() {
    const unsigned hc = std::max(1u, std::min(std::thread::hardware_concurrency(), 64u));
    return hostThreadLimit() ? std::min(hc, hostThreadLimit()) : hc;
}

constexpr int AA = 21;     // alphabet incl. X
constexpr int XCODE = 20;

struct SubMat {
    short sub[AA * AA];
    double pBack[AA];
    double prob[AA * AA];
};

enum MatrixId { BLOSUM62 = 0, VTML80 = 1 };

// SubstitutionMatrix ctor (lib/mmseqs/src/commons/SubstitutionMatrix.cpp:12-57) for the two shipped tables.
void buildSubMat(MatrixId which, float bitFactor, float scoreBias, SubMat &out);

// SubstitutionMatrix::setupLetterMapping (SubstitutionMatrix.cpp:257-298)
const uint8_t *aa2num();
void mapSequence(const char *s, uint64_t len, uint8_t *out);

// likelihood ratios of the tantan repeat masker (ProbabilityMatrix, lib/mmseqs/src/commons/BaseMatrix.h:83-96); the masking
// itself runs on the device (tantan.cu)
void likelihoodRatios(const SubMat &m, double *lr);

// Prefiltering::getKmerThreshold (Prefiltering.cpp:1005-1065), amino-acid sequence branch
int kmerThreshold(float sensitivity, int kmerSize);

// ExtendedSubstitutionMatrix::calcScoreMatrix for 3-mers (ExtendedSubstitutionMatrix.cpp:20-69):
// row x = all 8000 3-mers ordered by score desc, ties in first-letter-major enumeration order.
// cntGE[x*512 + 255 - v] = number of 3-mers whose score against x is >= v  (v in [-256,255])
void buildExt3(const SubMat &seedMat, std::vector<int16_t> &score, std::vector<uint16_t> &index, std::vector<uint16_t> &cntGE);

// QueryMatcher::initDiagonalMatcher (QueryMatcher.cpp:422-449) + Util::getL2CacheSize (Util.cpp:317-332)
int referenceBinSize(uint64_t nTargets);

// EvalueComputation (alignment/EvalueComputation.h) with the BLOSUM62 (11,1) Gumbel constants and ALP's
// finite-size-corrected area (lib/alp/sls_pvalues.cpp:366-532)
struct Evaluer {
    double lambda, K, logK, aI, aJ, bI, bJ, alphaI, alphaJ, betaI, betaJ, sigma, tau, viThr, vjThr, cThr;
    double dbResidues;
    explicit Evaluer(double dbResidues);
    double area(double score, double seqLen) const;
    double evalue(double score, double seqLen) const;
    int bitScore(double score) const;
};

}  // namespace sd
