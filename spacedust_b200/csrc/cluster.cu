// Hit aggregation and clustering on B200: besthitbyset / mergeresultsbyset / combinehits on the host
// (per-record decimal round trips that gate everything downstream), the agglomerative clustering
// of clusterhits as one CTA per genome pair on the device.
//
// Reference behaviour restated (paths relative to /root/reference):
//   src/util/besthitbyset.cpp:10-20,41-144     min-E hit per (query gene, target set), log P-value, %.3E
//   src/util/Aggregation.cpp:24-157            grouping by target set in ascending set key
//   src/util/combinehits.cpp:74-234            hits below the hard-coded 1e-6 threshold, exp(logP) %.3E
//   src/util/ClusterHits.cpp:80-181,359-454    scores, compatibility, merge loop with its stale dmin[]
//   src/util/ClusterHits.cpp:184-213,456-485   multihit P-value, output thresholds
//
// The reference keeps a dense K x K score matrix and rebuilds candidate clusters with vector copies and
// sorts (O(K^3)).  Its matrix entry is always a pure function of the two current clusters, and two
// compatible clusters never interleave in query order (query positions are unique inside a genome pair and
// ClusterHits.cpp:158 rejects overlapping ranges through unsigned wrap-around), so span, size and the number
// of conserved neighbour pairs of a union follow in O(1) from per-cluster summaries.  The kernel therefore
// keeps no matrix: it re-evaluates entries on demand and replays the reference's dmin[] updates verbatim,
// which gives the same merge sequence in O(K^2).  Transcendentals (Lanczos log-gamma, log(1 - m/k), log 2,
// log q0) are tabulated on the host with the same libm the reference links, the device only adds and
// multiplies doubles (compiled with --fmad=false), so scores compare bit-for-bit.
#include <mutex>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>

#include <chrono>
#include <cstdlib>
#include <string>

#include "sd_internal.h"

namespace sd {

constexpr int CL_THREADS = 256;
constexpr int CL_LOGTAB = 512;      // log(1 - m/k) tabulated for k < CL_LOGTAB

struct ClusterTables {
    const double *lgam;      // Lanczos log-gamma of 0..maxOrf+1 (ClusterHits.cpp:268-271)
    const double *log1m;     // [k][m] = log(1 - 1.0*m/k), k < CL_LOGTAB
    double logQ0, log2v, sMin;
    unsigned d;
};

struct Node {
    int k;                   // hits in the cluster, 0 = merged away
    unsigned iMin, iMax, jMin, jMax;
    int m;                   // conserved neighbour pairs inside the cluster
    unsigned firstT, lastT;  // target position of the first / last hit in query order
    int firstS, lastS;       // qStrand == tStrand of those hits
};

struct ClusterArgs {
    ClusterTables tab;
    const uint32_t *matchOff;  // nMatches+1
    const uint32_t *order;     // matches, largest first
    uint32_t nMatches;
    const uint32_t *qPos, *tPos;
    const uint8_t *sameStrand;
    Node *nodes;
    int *dmin;
    double *val;
    uint32_t *root;            // out: node each hit ended up in
    uint32_t *next;
    unsigned long long *stats; // [0] merges, [1] inexact log uses
};

__device__ __forceinline__ double clusterScore(const ClusterTables &T, int k, int span, int m, unsigned long long &inexact) {
    const double logpClu = 2 * T.lgam[span + 1] - 2 * T.lgam[span - k + 1] - T.lgam[k + 1] + k * T.logQ0;
    double l1m;
    if (k < CL_LOGTAB) l1m = T.log1m[k * CL_LOGTAB + m];
    else { l1m = log(1 - 1.0 * m / k); inexact++; }
    const double logpOrd = l1m - m * T.log2v - T.lgam[m + 1];
    return -0.5 * logpClu - 0.5 * logpOrd;
}

// DistMat[a][b] of the reference for the current clusters (0 for a == b, an emptied cluster, or incompatibility)
__device__ __forceinline__ double pairScore(const ClusterTables &T, const Node *nodes, int a, int b, unsigned long long &inexact) {
    if (a == b) return 0.0;
    const Node A = nodes[a], B = nodes[b];
    if (A.k == 0 || B.k == 0) return 0.0;
    const unsigned dj = min(A.jMin - B.jMax, B.jMin - A.jMax);     // unsigned on purpose (ClusterHits.cpp:158)
    const unsigned di = min(A.iMin - B.iMax, B.iMin - A.iMax);
    if (!(dj <= T.d && di <= T.d)) return 0.0;
    const bool aFirst = A.iMax < B.iMin;
    const Node &lo = aFirst ? A : B;
    const Node &hi = aFirst ? B : A;
    const bool sameOrder = hi.firstT > lo.lastT;
    const int junction = ((lo.lastS != 0) == sameOrder && (hi.firstS != 0) == sameOrder) ? 1 : 0;
    const int k = A.k + B.k;
    const int m = A.m + B.m + junction;
    const unsigned iMin = min(A.iMin, B.iMin), iMax = max(A.iMax, B.iMax), jMin = min(A.jMin, B.jMin), jMax = max(A.jMax, B.jMax);
    const int span = max((int)(iMax - iMin + 1), (int)(jMax - jMin + 1));
    return clusterScore(T, k, span, m, inexact);
}

// block-wide "first maximum": largest value, smallest index among equals
__device__ void blockArgMax(double &v, int &idx, double *sV, int *sI) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const double v2 = __shfl_down_sync(0xffffffffu, v, d);
        const int i2 = __shfl_down_sync(0xffffffffu, idx, d);
        if (v2 > v || (v2 == v && i2 < idx)) { v = v2; idx = i2; }
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sV[w] = v; sI[w] = idx; }
    __syncthreads();
    if (w == 0) {
        v = (l < CL_THREADS / 32) ? sV[l] : -DBL_MAX;
        idx = (l < CL_THREADS / 32) ? sI[l] : INT_MAX;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const double v2 = __shfl_down_sync(0xffffffffu, v, d);
            const int i2 = __shfl_down_sync(0xffffffffu, idx, d);
            if (v2 > v || (v2 == v && i2 < idx)) { v = v2; idx = i2; }
        }
        if (l == 0) { sV[0] = v; sI[0] = idx; }
    }
    __syncthreads();
    v = sV[0];
    idx = sI[0];
    __syncthreads();
}

__global__ void __launch_bounds__(CL_THREADS) clusterKernel(ClusterArgs A) {
    __shared__ double sV[CL_THREADS / 32];
    __shared__ int sI[CL_THREADS / 32];
    __shared__ uint32_t sMatch;
    unsigned long long inexact = 0, merges = 0;
    const ClusterTables T = A.tab;
    for (;;) {
        if (threadIdx.x == 0) sMatch = atomicAdd(A.next, 1u);
        __syncthreads();
        const uint32_t mi = sMatch;
        __syncthreads();
        if (mi >= A.nMatches) break;
        const uint32_t b = A.order[mi];
        const uint32_t base = A.matchOff[b];
        const int K = (int)(A.matchOff[b + 1] - base);
        Node *nodes = A.nodes + base;
        int *dmin = A.dmin + base;
        double *val = A.val + base;
        uint32_t *root = A.root + base;
        for (int i = threadIdx.x; i < K; i += CL_THREADS) {
            Node n;
            n.k = 1;
            n.iMin = n.iMax = A.qPos[base + i];
            n.jMin = n.jMax = A.tPos[base + i];
            n.m = 0;
            n.firstT = n.lastT = n.jMin;
            n.firstS = n.lastS = A.sameStrand[base + i];
            nodes[i] = n;
            root[i] = (uint32_t)i;
        }
        __syncthreads();
        // initial dmin: first maximum of every row (ClusterHits.cpp:377-388)
        for (int i = threadIdx.x; i < K; i += CL_THREADS) {
            int best = 0;
            double bestV = pairScore(T, nodes, i, 0, inexact);
            for (int j = 1; j < K; j++) {
                const double v = pairScore(T, nodes, i, j, inexact);
                if (v > bestV) { bestV = v; best = j; }
            }
            dmin[i] = best;
            val[i] = bestV;
        }
        __syncthreads();
        for (;;) {
            // closest pair: first row whose best entry is maximal (ClusterHits.cpp:399-404)
            double v = -DBL_MAX;
            int idx = INT_MAX;
            for (int i = threadIdx.x; i < K; i += CL_THREADS) {
                const double x = val[i];
                if (x > v) { v = x; idx = i; }
            }
            blockArgMax(v, idx, sV, sI);
            const double maxScore = v;
            const int i1 = idx;
            if (maxScore == 0) break;
            const int i2 = dmin[i1];
            __syncthreads();
            if (threadIdx.x == 0) {
                // append i2 to i1 (ClusterHits.cpp:416-419); summaries of the union
                const Node X = nodes[i1], Y = nodes[i2];
                const bool xFirst = X.iMax < Y.iMin;
                const Node &lo = xFirst ? X : Y;
                const Node &hi = xFirst ? Y : X;
                const bool sameOrder = hi.firstT > lo.lastT;
                Node n;
                n.k = X.k + Y.k;
                n.iMin = min(X.iMin, Y.iMin); n.iMax = max(X.iMax, Y.iMax);
                n.jMin = min(X.jMin, Y.jMin); n.jMax = max(X.jMax, Y.jMax);
                n.m = X.m + Y.m + (((lo.lastS != 0) == sameOrder && (hi.firstS != 0) == sameOrder) ? 1 : 0);
                n.firstT = lo.firstT; n.firstS = lo.firstS; n.lastT = hi.lastT; n.lastS = hi.lastS;
                nodes[i1] = n;
                Node e = Y;
                e.k = 0;
                nodes[i2] = e;
                root[i2] = (uint32_t)i1;
                merges++;
            }
            __syncthreads();
            // new row/column i1, stale-pointer updates of every other row (ClusterHits.cpp:423-451)
            double bx = -DBL_MAX;
            int bj = INT_MAX;
            for (int j = threadIdx.x; j < K; j += CL_THREADS) {
                const double x = (j == i1 || j == i2) ? 0.0 : pairScore(T, nodes, i1, j, inexact);
                if (x > bx) { bx = x; bj = j; }
                if (j != i1 && j != i2) {
                    int dj = dmin[j];
                    double cur = pairScore(T, nodes, j, dj, inexact);
                    if (x > cur) { dj = i1; cur = x; dmin[j] = dj; }
                    val[j] = cur;
                } else if (j == i2) {
                    val[j] = 0.0;
                }
            }
            blockArgMax(bx, bj, sV, sI);
            if (threadIdx.x == 0) { dmin[i1] = bj; val[i1] = bx; }
            __syncthreads();
            if (!(maxScore >= T.sMin)) break;     // do-while: the merge below sMin has been applied
        }
        __syncthreads();
        // resolve the merge forest
        for (int i = threadIdx.x; i < K; i += CL_THREADS) {
            uint32_t r = (uint32_t)i;
            while (nodes[r].k == 0) r = root[r];
            root[i] = r;
        }
        __syncthreads();
    }
    if (merges) atomicAdd(&A.stats[0], merges);
    if (inexact) atomicAdd(&A.stats[1], inexact);
}

// ---------------------------------------------------------------------------------------------- host side
static double roundTrip3E(double x) {
    char buf[48];
    snprintf(buf, sizeof buf, "%.3E", x);
    return strtod(buf, nullptr);
}

static double computeLogPval(double eval) {   // besthitbyset.cpp:10-20 with logCalibration = log(1)
    const double logCalibration = std::log(1);
    if (eval == 0) return std::log(DBL_MIN) - logCalibration;
    if (eval > 0 && eval < 10e-4) return std::log(eval) - logCalibration;
    return std::log(1 - std::exp(-eval)) - logCalibration;
}

static double lanczosLogGamma(double x) {     // ClusterHits.cpp:42-63, arguments >= 1 only
    static const double dk[11] = {2.48574089138753565546e-5, 1.05142378581721974210, -3.45687097222016235469,
                                  4.51227709466894823700, -2.98285225323576655721, 1.05639711577126713077,
                                  -1.95428773191645869583e-1, 1.70970543404441224307e-2, -5.71926117404305781283e-4,
                                  4.63399473359905636708e-6, -2.71994908488607703910e-9};
    const double r10 = 10.900511;
    const double c = 2 * std::sqrt(std::exp(1.0) / M_PI);
    if (x == 1) return 0.0;
    double sum = dk[0];
    for (int i = 1; i <= 10; i++) sum += dk[i] / (x + (i - 1));
    return std::log(c) + (x - 0.5) * std::log(x + r10 - 0.5) - (x - 0.5) + std::log(sum);
}

struct BestHit {
    uint32_t aln, q, qSet, tSet;
    double logP;          // value besthitbyset prints (after %.3E)
};

}  // namespace sd

using namespace sd;

// ClusterHits.cpp:359-485 for a list of matches (genome pairs): pairs with at least two hits go to the device, clusters come back in
// match order, then node order; member.aln = index of the hit in the caller's hit arrays.  Shared by sd_cluster_run (hits = kept best
// hits of the fused path) and sd_clusterhits_run (hits = lines of a `matches` DB).
struct Match { uint32_t qSet, tSet; uint64_t first; uint32_t n, nq; };
struct ClusterCoreOut {
    std::vector<sd_cluster> clusters;
    std::vector<sd_cluster_member> members;
    uint64_t maxK = 0;
    uint32_t nDev = 0;
    unsigned long long hStat[2] = {0, 0};
    double kernelMs = 0;
};

template <typename Tick>
static int clusterMatchesCore(sd_context *ctx, cudaStream_t st, cudaEvent_t evA, cudaEvent_t evB, const std::vector<Match> &matches,
                              const uint32_t *hitQPos, const uint32_t *hitTPos, const uint8_t *hitSame, const double *mPval, uint32_t maxOrf,
                              const sd_cluster_params *par, ClusterCoreOut &out, Tick tick) {
    // ---- clusterhits: genome pairs with at least two hits go to the device
    std::vector<double> lgam(maxOrf + 2);
    lgam[0] = INFINITY;
    for (uint32_t i = 1; i < maxOrf + 2; i++) lgam[i] = lanczosLogGamma((double)i);
    const unsigned d = (unsigned)par->max_gene_gap;
    auto hostScore = [&](int k, int span, int m) {
        const double logpClu = 2 * lgam[span + 1] - 2 * lgam[span - k + 1] - lgam[k + 1] + k * std::log(0.001);
        const double logpOrd = std::log(1 - 1.0 * m / k) - m * std::log(2) - lgam[m + 1];
        return -0.5 * logpClu - 0.5 * logpOrd;
    };
    if (d + 2 >= lgam.size()) { setError("sd_cluster_run: --max-gene-gap larger than the biggest set"); return SD_EINVAL; }
    const double sMin = hostScore(2, (int)d + 1, 1);

    std::vector<uint32_t> dMatchOff(1, 0), hQ, hT, devMatch;
    std::vector<uint8_t> hS;
    uint64_t maxK = 0;
    for (size_t mi = 0; mi < matches.size(); mi++) {
        const Match &m = matches[mi];
        maxK = std::max<uint64_t>(maxK, m.n);
        if (m.n < 2) continue;
        std::vector<uint32_t> seen;
        seen.reserve(m.n);
        for (uint32_t h = 0; h < m.n; h++) {
            const uint64_t g = m.first + h;
            if (hitQPos[g] + 1 >= lgam.size() || hitTPos[g] + 1 >= lgam.size()) { setError("sd_cluster_run: gene position beyond the set size"); return SD_EINVAL; }
            hQ.push_back(hitQPos[g]);
            hT.push_back(hitTPos[g]);
            hS.push_back(hitSame[g] ? 1 : 0);
            seen.push_back(hitQPos[g]);
        }
        std::sort(seen.begin(), seen.end());
        if (std::adjacent_find(seen.begin(), seen.end()) != seen.end()) {
            setError("sd_cluster_run: query positions repeat inside genome pair (%u,%u): --suboptimal-hits > 0 style input is not supported", m.qSet, m.tSet);
            return SD_ELIMIT;
        }
        devMatch.push_back((uint32_t)mi);
        dMatchOff.push_back((uint32_t)hQ.size());
    }
    tick("matches");
    const uint32_t nDev = (uint32_t)devMatch.size();
    std::vector<uint32_t> hRoot(hQ.size());
    unsigned long long hStat[2] = {0, 0};
    double kernelMs = 0;
    if (nDev) {
        std::vector<uint32_t> order(nDev);
        for (uint32_t i = 0; i < nDev; i++) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return dMatchOff[a + 1] - dMatchOff[a] > dMatchOff[b + 1] - dMatchOff[b]; });
        std::vector<double> log1m((size_t)CL_LOGTAB * CL_LOGTAB, 0.0);
        for (int k = 1; k < CL_LOGTAB; k++)
            for (int m = 0; m < k; m++) log1m[(size_t)k * CL_LOGTAB + m] = std::log(1 - 1.0 * m / k);
        DevBuf<double> dLgam, dLog1m, dVal;
        DevBuf<uint32_t> dOff, dOrder, dQ, dT, dRoot, dNext;
        DevBuf<uint8_t> dS;
        DevBuf<Node> dNodes;
        DevBuf<int> dDmin;
        DevBuf<unsigned long long> dStats;
        int rc;
        const size_t nH = hQ.size();
        if ((rc = dLgam.alloc(lgam.size())) || (rc = dLog1m.alloc(log1m.size())) || (rc = dVal.alloc(nH)) || (rc = dOff.alloc(nDev + 1)) ||
            (rc = dOrder.alloc(nDev)) || (rc = dQ.alloc(nH)) || (rc = dT.alloc(nH)) || (rc = dRoot.alloc(nH)) || (rc = dNext.alloc(1)) ||
            (rc = dS.alloc(nH)) || (rc = dNodes.alloc(nH)) || (rc = dDmin.alloc(nH)) || (rc = dStats.alloc(2)))
            return rc;
        SD_CUDA(cudaMemcpyAsync(dLgam.p, lgam.data(), lgam.size() * sizeof(double), cudaMemcpyHostToDevice, st));
        SD_CUDA(cudaMemcpyAsync(dLog1m.p, log1m.data(), log1m.size() * sizeof(double), cudaMemcpyHostToDevice, st));
        SD_CUDA(cudaMemcpyAsync(dOff.p, dMatchOff.data(), (nDev + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        SD_CUDA(cudaMemcpyAsync(dOrder.p, order.data(), nDev * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        SD_CUDA(cudaMemcpyAsync(dQ.p, hQ.data(), nH * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        SD_CUDA(cudaMemcpyAsync(dT.p, hT.data(), nH * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        SD_CUDA(cudaMemcpyAsync(dS.p, hS.data(), nH, cudaMemcpyHostToDevice, st));
        SD_CUDA(cudaMemsetAsync(dNext.p, 0, sizeof(uint32_t), st));
        SD_CUDA(cudaMemsetAsync(dStats.p, 0, 2 * sizeof(unsigned long long), st));
        ClusterArgs A;
        A.tab.lgam = dLgam.p; A.tab.log1m = dLog1m.p; A.tab.logQ0 = std::log(0.001); A.tab.log2v = std::log(2); A.tab.sMin = sMin; A.tab.d = d;
        A.matchOff = dOff.p; A.order = dOrder.p; A.nMatches = nDev; A.qPos = dQ.p; A.tPos = dT.p; A.sameStrand = dS.p;
        A.nodes = dNodes.p; A.dmin = dDmin.p; A.val = dVal.p; A.root = dRoot.p; A.next = dNext.p; A.stats = dStats.p;
        const uint32_t nCta = std::min<uint32_t>(nDev, (uint32_t)ctx->smCount * 4u);
        // fused clustersearch: two host threads share the GPU; take the turn so that the event pair brackets this kernel and not
        // the other chunk's Smith-Waterman launches queued before it (the kernel is a few ms)
        std::unique_lock<std::mutex> gpuTurn;
        if (ctx->gpuToken) gpuTurn = std::unique_lock<std::mutex>(*ctx->gpuToken);
        SD_CUDA(cudaEventRecord(evA, st));
        clusterKernel<<<nCta, CL_THREADS, 0, st>>>(A);
        ctx->launches++;
        SD_CUDA(cudaGetLastError());
        SD_CUDA(cudaEventRecord(evB, st));
        SD_CUDA(cudaMemcpyAsync(hRoot.data(), dRoot.p, nH * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        SD_CUDA(cudaMemcpyAsync(hStat, dStats.p, sizeof hStat, cudaMemcpyDeviceToHost, st));
        SD_CUDA(cudaStreamSynchronize(st));
        if (gpuTurn.owns_lock()) gpuTurn.unlock();
        float ms = 0;
        cudaEventElapsedTime(&ms, evA, evB);
        kernelMs = ms;
    }
    tick("kernel");
    // ---- clusters of each genome pair in node order; thresholds (ClusterHits.cpp:456-485)
    const uint32_t cls = (uint32_t)std::max(par->cluster_size, 0);
    std::vector<std::vector<sd_cluster>> partC(nDev);
    std::vector<std::vector<sd_cluster_member>> partM(nDev);
    parallelItems(nDev, [&](uint64_t di64) {
        const uint32_t di = (uint32_t)di64;
        std::vector<sd_cluster> &outC = partC[di];
        std::vector<sd_cluster_member> &outM = partM[di];
        const Match &m = matches[devMatch[di]];
        const uint32_t base = dMatchOff[di];
        std::map<uint32_t, std::vector<uint32_t>> byRoot;
        for (uint32_t h = 0; h < m.n; h++) byRoot[hRoot[base + h]].push_back(h);
        for (auto &kv : byRoot) {
            std::vector<uint32_t> &mem = kv.second;
            if (mem.size() < cls) continue;
            std::sort(mem.begin(), mem.end(), [&](uint32_t a, uint32_t b) { return hQ[base + a] < hQ[base + b]; });   // findConservedPairs sorts by qPos
            unsigned iMin = UINT_MAX, iMax = 0, jMin = UINT_MAX, jMax = 0;
            for (uint32_t h : mem) {
                iMin = std::min(iMin, hQ[base + h]); iMax = std::max(iMax, hQ[base + h]);
                jMin = std::min(jMin, hT[base + h]); jMax = std::max(jMax, hT[base + h]);
            }
            const int span = std::max((int)(iMax - iMin + 1), (int)(jMax - jMin + 1));
            int cons = 0;
            for (size_t l = 0; l + 1 < mem.size(); l++) {
                const bool sameOrder = hT[base + mem[l + 1]] > hT[base + mem[l]];
                if ((hS[base + mem[l]] != 0) == sameOrder && (hS[base + mem[l + 1]] != 0) == sameOrder) cons++;
            }
            const int k = (int)mem.size();
            const double pCO = std::exp(-hostScore(k, span, cons));
            // multihitPval (ClusterHits.cpp:184-213)
            const uint32_t Nq = m.nq;
            double pMH;
            {
                size_t kk = 0;
                double r = 0;
                const double thr = std::log((double)par->alpha / (Nq + 1));
                for (uint32_t h : mem) {
                    const double lp = std::log(mPval[m.first + h]);
                    if (lp < thr) { kk++; r -= lp - thr; }
                }
                if (r == 0) pMH = 1.0;
                else if (std::isinf(r)) pMH = 0.0;
                else {
                    const double emr = std::exp(-r);
                    if (emr == 0) pMH = 0.0;
                    else {
                        double sum = 0;
                        for (size_t i = 0; i + 1 < kk; i++) sum += std::pow(r, (double)i) / std::exp(lgam[i + 1]);
                        pMH = emr * sum;
                    }
                }
            }
            if (!(pCO <= par->cluster_pval && pMH <= par->multihit_pval)) continue;
            sd_cluster c;
            c.q_set = m.qSet; c.t_set = m.tSet; c.p_co = pCO; c.p_mh = pMH; c.n = (uint32_t)k; c.first = (uint32_t)outM.size();
            for (uint32_t h : mem) {
                sd_cluster_member mm;
                mm.aln = (uint32_t)(m.first + h); mm.pad = 0; mm.pval = mPval[m.first + h];
                outM.push_back(mm);
            }
            outC.push_back(c);
        }
    });
    std::vector<sd_cluster> outC;
    std::vector<sd_cluster_member> outM;
    for (uint32_t di = 0; di < nDev; di++) {
        for (sd_cluster c : partC[di]) { c.first += (uint32_t)outM.size(); outC.push_back(c); }
        outM.insert(outM.end(), partM[di].begin(), partM[di].end());
    }
    out.clusters.swap(outC);
    out.members.swap(outM);
    out.maxK = maxK; out.nDev = nDev; out.hStat[0] = hStat[0]; out.hStat[1] = hStat[1]; out.kernelMs = kernelMs;
    return SD_OK;
}

extern "C" void sd_cluster_params_default(sd_cluster_params *p) {
    p->alpha = 1.0f;
    p->filter_self_match = 0;
    p->max_gene_gap = 3;
    p->cluster_size = 2;
    p->cluster_pval = 0.01;
    p->multihit_pval = 0.01;
}

extern "C" int sd_cluster_run(sd_context *ctx, const sd_alignment *alns, uint64_t n_alns,
                              const uint32_t *q_set_of, const uint32_t *q_pos, const uint8_t *q_strand, uint32_t nq,
                              const uint32_t *t_set_of, const uint32_t *t_pos, const uint8_t *t_strand, uint32_t nt,
                              const uint32_t *q_set_size, uint32_t n_q_sets, const uint32_t *t_set_size, uint32_t n_t_sets,
                              const sd_cluster_params *par, sd_cluster *clusters, uint64_t cap, uint64_t *n_clusters,
                              sd_cluster_member *members, uint64_t mcap, uint64_t *n_members, sd_cluster_stats *stats) {
    if (!ctx || !par || !n_clusters || !n_members || !q_set_of || !q_pos || !q_strand || !t_set_of || !t_pos || !t_strand ||
        !q_set_size || !t_set_size || (n_alns && !alns)) { setError("sd_cluster_run: bad argument"); return SD_EINVAL; }
    *n_clusters = 0; *n_members = 0;
    if (stats) memset(stats, 0, sizeof *stats);
    if (n_alns == 0) return SD_OK;
    if (n_alns >= 0xffffffffull) { setError("sd_cluster_run: too many alignments in one call"); return SD_ELIMIT; }
    SD_CUDA(cudaSetDevice(ctx->device));
    sd::StreamScope allocScope(ctx->stream);
    cudaStream_t st = ctx->stream;
    sd::EventQuad events;
    { const int erc = events.create(); if (erc) return erc; }
    cudaEvent_t ev0 = events.e[0], ev1 = events.e[1], evA = events.e[2], evB = events.e[3];
    SD_CUDA(cudaEventRecord(ev0, st));

    const bool timing = getenv("SD_TIMING") != nullptr;
    auto tPrev = std::chrono::steady_clock::now();
    std::string tLog;
    auto tick = [&](const char *name) {
        if (!timing) return;
        auto now = std::chrono::steady_clock::now();
        char b[96];
        snprintf(b, sizeof b, " %s=%.1f", name, std::chrono::duration<double, std::milli>(now - tPrev).count());
        tLog += b;
        tPrev = now;
    };
    // ---- besthitbyset: per query gene and target set the first alignment with the smallest printed E-value
    std::vector<BestHit> best;
    best.reserve(n_alns);
    std::vector<double> evPrinted(n_alns);        // align prints %.3E, besthitbyset parses it
    parallelFor(n_alns, [&](uint64_t i) { evPrinted[i] = roundTrip3E(alns[i].evalue); });
    tick("evprint");
    {
        // query groups (alignments arrive grouped by ascending query), validated once
        std::vector<uint64_t> gStart;
        gStart.reserve(nq + 1);
        for (uint64_t i = 0; i < n_alns; i++) {
            if (i == 0 || alns[i].query != alns[i - 1].query) {
                if (i && alns[i].query < alns[i - 1].query) { setError("sd_cluster_run: alignments must be grouped by ascending query"); return SD_EINVAL; }
                if (alns[i].query >= nq) { setError("sd_cluster_run: alignment %llu has query %u out of range", (unsigned long long)i, alns[i].query); return SD_EINVAL; }
                gStart.push_back(i);
            }
            if (alns[i].target >= nt) { setError("sd_cluster_run: alignment %llu has target out of range", (unsigned long long)i); return SD_EINVAL; }
        }
        const uint64_t nGroups = gStart.size();
        gStart.push_back(n_alns);
        const unsigned nChunks = 256;
        std::vector<std::vector<BestHit>> part(nChunks);
        std::atomic<int> badSet{0};
        parallelChunks(nGroups, nChunks, [&](unsigned c, uint64_t lo, uint64_t hi) {
            std::vector<std::pair<uint32_t, uint32_t>> grp;   // (target set, alignment index)
            std::vector<BestHit> &out = part[c];
            for (uint64_t gi = lo; gi < hi; gi++) {
                const uint32_t q = alns[gStart[gi]].query;
                grp.clear();
                for (uint64_t j = gStart[gi]; j < gStart[gi + 1]; j++) grp.emplace_back(t_set_of[alns[j].target], (uint32_t)j);
                std::stable_sort(grp.begin(), grp.end(), [](const std::pair<uint32_t, uint32_t> &a, const std::pair<uint32_t, uint32_t> &b) { return a.first < b.first; });
                size_t g = 0;
                while (g < grp.size()) {
                    size_t h = g;
                    double bestEval = DBL_MAX;
                    int64_t bestIdx = -1;
                    while (h < grp.size() && grp[h].first == grp[g].first) {
                        const double ev = evPrinted[grp[h].second];
                        if (ev < bestEval) { bestEval = ev; bestIdx = grp[h].second; }
                        h++;
                    }
                    if (bestIdx >= 0) {
                        BestHit bh;
                        bh.aln = (uint32_t)bestIdx; bh.q = q; bh.qSet = q_set_of[q]; bh.tSet = grp[g].first;
                        bh.logP = bestEval;      // turned into the printed log P-value below (parallel)
                        if (bh.qSet >= n_q_sets || bh.tSet >= n_t_sets) badSet = 1;
                        out.push_back(bh);
                    }
                    g = h;
                }
            }
        });
        if (badSet) { setError("sd_cluster_run: set index out of range"); return SD_EINVAL; }
        for (unsigned c = 0; c < nChunks; c++) best.insert(best.end(), part[c].begin(), part[c].end());
    }
    tick("besthit");
    parallelFor(best.size(), [&](uint64_t i) { best[i].logP = roundTrip3E(computeLogPval(best[i].logP)); });
    // ---- mergeresultsbyset + combinehits: group by (query set, target set), members in gene order
    if ((uint64_t)n_q_sets * n_t_sets <= (1ull << 24)) {
        // stable counting sort on the genome pair
        std::vector<uint32_t> start((size_t)n_q_sets * n_t_sets + 1, 0);
        for (const BestHit &b : best) start[(size_t)b.qSet * n_t_sets + b.tSet + 1]++;
        for (size_t i = 1; i < start.size(); i++) start[i] += start[i - 1];
        std::vector<BestHit> sorted(best.size());
        for (const BestHit &b : best) sorted[start[(size_t)b.qSet * n_t_sets + b.tSet]++] = b;
        best.swap(sorted);
    } else {
        std::stable_sort(best.begin(), best.end(), [](const BestHit &a, const BestHit &b) {
            if (a.qSet != b.qSet) return a.qSet < b.qSet;
            return a.tSet < b.tSet;
        });
    }
    tick("pairsort");
    std::vector<Match> matches;
    std::vector<uint32_t> mAln;       // alignment index per kept hit
    std::vector<double> mPval;        // P-value column as combinehits prints it
    const double logPvalThr = std::log(10e-7);         // combinehits.cpp:101 (hard-coded)
    {
        size_t i = 0;
        while (i < best.size()) {
            size_t j = i;
            while (j < best.size() && best[j].qSet == best[i].qSet && best[j].tSet == best[i].tSet) j++;
            if (!(par->filter_self_match && best[i].qSet == best[i].tSet)) {
                size_t k = 0;
                double r = 0;
                const uint32_t first = (uint32_t)mAln.size();
                for (size_t h = i; h < j; h++) {
                    if (best[h].logP < logPvalThr) {
                        k++;
                        r -= best[h].logP - logPvalThr;
                        mAln.push_back(best[h].aln);
                        mPval.push_back(best[h].logP);     // exp + %.3E round trip below (parallel)
                    }
                }
                if (r == 0 || k == 0) { mAln.resize(first); mPval.resize(first); }
                else matches.push_back(Match{best[i].qSet, best[i].tSet, first, (uint32_t)k, q_set_size[best[i].qSet]});
            }
            i = j;
        }
    }
    parallelFor(mPval.size(), [&](uint64_t i) { mPval[i] = roundTrip3E(std::exp(mPval[i])); });
    tick("combine");
    // ---- clusterhits on the kept hits (clusterMatchesCore): positions / strands / printed P-values per kept hit
    uint32_t maxOrf = 0;
    for (uint32_t s = 0; s < n_q_sets; s++) maxOrf = std::max(maxOrf, q_set_size[s]);
    for (uint32_t s = 0; s < n_t_sets; s++) maxOrf = std::max(maxOrf, t_set_size[s]);
    std::vector<uint32_t> hitQPos(mAln.size()), hitTPos(mAln.size());
    std::vector<uint8_t> hitSame(mAln.size());
    parallelFor(mAln.size(), [&](uint64_t i) {
        const sd_alignment &a = alns[mAln[i]];
        hitQPos[i] = q_pos[a.query];
        hitTPos[i] = t_pos[a.target];
        hitSame[i] = (q_strand[a.query] != 0) == (t_strand[a.target] != 0) ? 1 : 0;
    });
    ClusterCoreOut core;
    {
        const int crc = clusterMatchesCore(ctx, st, evA, evB, matches, hitQPos.data(), hitTPos.data(), hitSame.data(), mPval.data(), maxOrf, par, core, tick);
        if (crc) return crc;
    }
    std::vector<sd_cluster> &outC = core.clusters;
    std::vector<sd_cluster_member> &outM = core.members;
    parallelFor(outM.size(), [&](uint64_t i) { outM[i].aln = mAln[outM[i].aln]; });
    const uint64_t maxK = core.maxK;
    const uint32_t nDev = core.nDev;
    const unsigned long long *hStat = core.hStat;
    const double kernelMs = core.kernelMs;
    *n_clusters = outC.size();
    *n_members = outM.size();
    if (outC.size() > cap || outM.size() > mcap || (outC.size() && (!clusters || !members))) {
        setError("sd_cluster_run: needs room for %zu clusters / %zu members", outC.size(), outM.size());
        return SD_ESMALL;
    }
    if (!outC.empty()) memcpy(clusters, outC.data(), outC.size() * sizeof(sd_cluster));
    if (!outM.empty()) memcpy(members, outM.data(), outM.size() * sizeof(sd_cluster_member));
    tick("assemble");
    if (timing) fprintf(stderr, "[sd_cluster]%s\n", tLog.c_str());
    SD_CUDA(cudaEventRecord(ev1, st));
    SD_CUDA(cudaStreamSynchronize(st));
    if (stats) {
        float ms = 0;
        cudaEventElapsedTime(&ms, ev0, ev1);
        stats->best_hits = best.size(); stats->genome_pairs = matches.size(); stats->pairs_clustered = nDev;
        stats->max_hits_per_pair = maxK; stats->merges = hStat[0]; stats->inexact_log = hStat[1];
        stats->kernel_ms = kernelMs; stats->total_ms = ms;
    }
    return SD_OK;
}

// The clusterhits module body on its own (src/util/ClusterHits.cpp:294-485) for matches that come from a `matches` DB instead of
// the fused path: match i owns hits first[i] .. first[i+1]-1 (query-gene position, target-gene position, same-strand flag and the
// P-value column of every hit), nq[i] = genes of its query set (header column 3).  members[].aln = index of the hit.
extern "C" int sd_clusterhits_run(sd_context *ctx, uint32_t n_matches, const uint64_t *first, const uint32_t *q_set, const uint32_t *t_set,
                                  const uint32_t *nq, const uint32_t *hit_q_pos, const uint32_t *hit_t_pos, const uint8_t *hit_same_strand,
                                  const double *hit_pval, uint32_t max_set_size, const sd_cluster_params *par, sd_cluster *clusters, uint64_t cap,
                                  uint64_t *n_clusters, sd_cluster_member *members, uint64_t mcap, uint64_t *n_members, sd_cluster_stats *stats) {
    if (!ctx || !par || !n_clusters || !n_members || !first || (n_matches && (!q_set || !t_set || !nq))) { setError("sd_clusterhits_run: bad argument"); return SD_EINVAL; }
    *n_clusters = 0; *n_members = 0;
    if (stats) memset(stats, 0, sizeof *stats);
    if (n_matches == 0) return SD_OK;
    if (first[n_matches] >= 0xffffffffull) { setError("sd_clusterhits_run: too many hits in one call"); return SD_ELIMIT; }
    if (first[n_matches] && (!hit_q_pos || !hit_t_pos || !hit_same_strand || !hit_pval)) { setError("sd_clusterhits_run: bad argument"); return SD_EINVAL; }
    SD_CUDA(cudaSetDevice(ctx->device));
    sd::StreamScope allocScope(ctx->stream);
    cudaStream_t st = ctx->stream;
    sd::EventQuad events;
    { const int erc = events.create(); if (erc) return erc; }
    SD_CUDA(cudaEventRecord(events.e[0], st));
    std::vector<Match> matches(n_matches);
    for (uint32_t i = 0; i < n_matches; i++) {
        if (first[i + 1] < first[i]) { setError("sd_clusterhits_run: match offsets must ascend"); return SD_EINVAL; }
        matches[i] = Match{q_set[i], t_set[i], first[i], (uint32_t)(first[i + 1] - first[i]), nq[i]};
    }
    ClusterCoreOut core;
    const int crc = clusterMatchesCore(ctx, st, events.e[2], events.e[3], matches, hit_q_pos, hit_t_pos, hit_same_strand, hit_pval, max_set_size, par, core,
                                       [](const char *) {});
    if (crc) return crc;
    *n_clusters = core.clusters.size();
    *n_members = core.members.size();
    if (core.clusters.size() > cap || core.members.size() > mcap || (core.clusters.size() && (!clusters || !members))) {
        setError("sd_clusterhits_run: needs room for %zu clusters / %zu members", core.clusters.size(), core.members.size());
        return SD_ESMALL;
    }
    if (!core.clusters.empty()) memcpy(clusters, core.clusters.data(), core.clusters.size() * sizeof(sd_cluster));
    if (!core.members.empty()) memcpy(members, core.members.data(), core.members.size() * sizeof(sd_cluster_member));
    SD_CUDA(cudaEventRecord(events.e[1], st));
    SD_CUDA(cudaStreamSynchronize(st));
    if (stats) {
        float ms = 0;
        cudaEventElapsedTime(&ms, events.e[0], events.e[1]);
        stats->genome_pairs = n_matches; stats->pairs_clustered = core.nDev; stats->max_hits_per_pair = core.maxK;
        stats->merges = core.hStat[0]; stats->inexact_log = core.hStat[1]; stats->kernel_ms = core.kernelMs; stats->total_ms = ms;
    }
    return SD_OK;
}
