// k-mer prefilter on B200 (sm_100a): target index build, query preparation, and the fused
// similar-k-mer / index probe / double-diagonal-hit / ungapped-diagonal-score / top-N kernel.
//
// Reference behaviour restated here (all paths relative to /root/reference/lib/mmseqs/src):
//   prefiltering/IndexBuilder.cpp:55-239, IndexTable.h:133-173,348-401   target index
//   prefiltering/QueryMatcher.cpp:85-346                                  matchQuery / match
//   prefiltering/KmerGenerator.cpp:107-216                                similar k-mer list (order!)
//   prefiltering/CacheFriendlyOperations.cpp:185-274,350-380              double hit / best diagonal
//   prefiltering/UngappedAlignment.cpp:30-43,385-430                      diagonal score
//   prefiltering/QueryMatcher.cpp:363-420, QueryMatcher.h:38-48,206-216   threshold, top-N, sort
//   prefiltering/Prefiltering.cpp:845-869                                 coverage pre-filter
//
// Design: one warp owns one query at a time and walks its similar-k-mer probes in the
// reference's stream order, 32 probes / 32 index entries per step, so the order-dependent double-hit
// rule ("same 8-bit diagonal as the previous hit of this target") is evaluated with __match_any_sync
// inside the step and a per-warp last-diagonal table in HBM between steps.  No sort, no atomics on
// the hit stream; the stream itself is never materialised.
#include <cub/cub.cuh>

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <memory>
#include <numeric>

#include "sd_internal.h"
#include "prep.cuh"

namespace sd {

__constant__ int c_seedPos[KMER] = {0, 1, 3, 5, 8, 9};

// ------------------------------------------------------------------------------------------------
// small warp helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t laneId() { return threadIdx.x & 31; }

__device__ __forceinline__ uint32_t warpExclusiveScan(uint32_t v, uint32_t &total) {
    uint32_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
        if (laneId() >= (uint32_t)d) x += y;
    }
    total = __shfl_sync(0xffffffffu, x, 31);
    return x - v;
}

// largest lane j with excl[j] <= x  (x < total); every lane must call
__device__ __forceinline__ uint32_t warpFindOwner(uint32_t excl, uint32_t x) {
    uint32_t lo = 0, hi = 31;
#pragma unroll
    for (int it = 0; it < 5; it++) {
        uint32_t mid = (lo + hi + 1) >> 1;
        uint32_t v = __shfl_sync(0xffffffffu, excl, mid);
        if (v <= x) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// Owner of slot (base + lane) in a warp-wide segmented range: lane j owns slots [excl[j], excl[j] + len[j]).
// Requires the lanes with len != 0 to form a prefix of the warp (then rank among them == lane index): one ballot for
// the segments that start at or before `base`, one warp OR-reduction (REDUX) for the heads inside the 32-slot window.
__device__ __forceinline__ uint32_t warpOwnerOfSlot(uint32_t excl, bool nonEmpty, uint32_t base, uint32_t lane) {
    const uint32_t before = __ballot_sync(0xffffffffu, nonEmpty && excl <= base);
    const uint32_t rel = excl - base;
    const uint32_t heads = __reduce_or_sync(0xffffffffu, (nonEmpty && rel - 1u < 31u) ? (1u << rel) : 0u);
    return (uint32_t)__popc(before) - 1u + (uint32_t)__popc(heads & (0xffffffffu >> (31u - lane)));
}

// number of neighbours of 3-mer x with score >= v (cumulative table built with the sorted rows)
__device__ __forceinline__ uint32_t countGE(const uint16_t *__restrict__ cnt, uint32_t x, int v) {
    if (v > 255) return 0u;
    if (v < -256) return N3;
    return (uint32_t)__ldg(&cnt[(size_t)x * 512 + (uint32_t)(255 - v)]);
}

// ------------------------------------------------------------------------------------------------
// index build
// ------------------------------------------------------------------------------------------------
// key = seq<<42 | kmer<<16 | pos for indexable k-mers, ~0 otherwise (IndexTable::addKmerCount)
__global__ void extractTargetKmers(const uint8_t *__restrict__ res, const uint32_t *__restrict__ off, uint32_t nT,
                                   uint32_t totalRes, const int8_t *__restrict__ selfScore, int kmerThr,
                                   uint64_t *__restrict__ keys) {
    uint32_t gp = blockIdx.x * blockDim.x + threadIdx.x;
    if (gp >= totalRes) return;
    uint32_t s = seqOfPosition(off, nT, gp);
    uint32_t p = gp - off[s];
    uint32_t L = off[s + 1] - off[s];
    uint64_t key = ~0ull;
    if (p + SPAN <= L && p < 65536u) {
        uint32_t idx = 0, mul = 1;
        int self = 0;
        bool ok = true;
#pragma unroll
        for (int k = 0; k < KMER; k++) {
            uint8_t c = res[gp + c_seedPos[k]];
            ok = ok && (c < 20);
            self += selfScore[c];
            idx += (uint32_t)c * mul;
            mul *= 20;
        }
        if (ok && !(kmerThr > 0 && self < kmerThr)) key = ((uint64_t)s << 42) | ((uint64_t)idx << 16) | p;
    }
    keys[gp] = key;
}

// after sorting by (seq,kmer,pos): the first key of each (seq,kmer) run is the entry the reference keeps
__global__ void firstOccurrence(const uint64_t *__restrict__ sorted, uint32_t n, uint32_t *__restrict__ kmerKey,
                                uint64_t *__restrict__ payload, uint32_t *__restrict__ kCount) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t k = sorted[i];
    bool keep = (k != ~0ull) && (i == 0 || (sorted[i - 1] >> 16) != (k >> 16));
    uint32_t kmer = (uint32_t)((k >> 16) & 0x3ffffffu);
    kmerKey[i] = keep ? kmer : 0xffffffffu;
    payload[i] = ((k >> 42) << 16) | (k & 0xffffu);   // seq<<16 | pos
    if (keep) atomicAdd(&kCount[kmer], 1u);
}

__global__ void packEntries(const uint64_t *__restrict__ payload, uint64_t n, uint2 *__restrict__ ent) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t v = payload[i];
    ent[i] = make_uint2((uint32_t)(v >> 16), (uint32_t)(v & 0xffffu));
}

// Occupancy bitmap of the k-mer space: one warp turns 32x32 bucket sizes into 32 words (coalesced reads of kOff).
// bits[w] = k-mers 32w..32w+31 with a non-empty bucket (8 MB for 20^6 k-mers: meant to stay in the 126 MB L2, so an
// empty probe never reaches HBM); rank[w] (scan below) = number of non-empty k-mers before 32w = row of the first one.
__global__ void occupancyBits(const uint32_t *__restrict__ kOff, uint32_t nWords, uint32_t *__restrict__ bits,
                              uint32_t *__restrict__ cnt) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t w0 = warp * 32;
    if (w0 >= nWords) return;
    uint32_t mine = 0;
    for (uint32_t j = 0; j < 32 && w0 + j < nWords; j++) {
        const uint32_t k = (w0 + j) * 32 + lane;
        const bool full = (k < NKMER) && (kOff[k + 1] != kOff[k]);
        const uint32_t b = __ballot_sync(0xffffffffu, full);
        if (j == lane) mine = b;
    }
    if (w0 + lane < nWords) { bits[w0 + lane] = mine; cnt[w0 + lane] = __popc(mine); }
}

// One 8-byte row per non-empty k-mer: a bucket with a single entry is stored inline {target, 1<<31 | pos} (no second
// random access), longer buckets as {start, length} into ent[].
__global__ void compactRows(const uint32_t *__restrict__ kOff, const uint32_t *__restrict__ bits,
                            const uint32_t *__restrict__ rank, uint32_t nWords, const uint2 *__restrict__ ent,
                            uint2 *__restrict__ rows) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t w = k >> 5;
    if (w >= nWords || k >= NKMER) return;
    const uint32_t b = bits[w];
    if (!((b >> (k & 31)) & 1u)) return;
    const uint32_t start = kOff[k], len = kOff[k + 1] - start;
    uint2 row = make_uint2(start, len);
    if (len == 1) { const uint2 e = ent[start]; row = make_uint2(e.x, 0x80000000u | e.y); }
    rows[rank[w] + __popc(b & ((1u << (k & 31)) - 1u))] = row;
}

// ------------------------------------------------------------------------------------------------
// query preparation
// ------------------------------------------------------------------------------------------------
struct QueryBatchDev {
    const uint8_t *code;     // residue codes, concatenated
    const uint32_t *off;     // nQ+1
    const int8_t *corr;      // per-position profile correction  (UngappedAlignment::createProfile)
    const int16_t *thr;      // per-position similar-k-mer threshold, -1 = no k-mer here
};

// UngappedAlignment::createProfile (UngappedAlignment.cpp:393-396): aaCorrBias/4 -+ 0.5 evaluated in double,
// stored to float, truncated to char
__global__ void profileCorrection(const float *__restrict__ bias, uint32_t totalRes, int8_t *__restrict__ corr) {
    uint32_t gp = blockIdx.x * blockDim.x + threadIdx.x;
    if (gp >= totalRes) return;
    const float b = bias[gp];
    const float q = __fdiv_rn(b, 4.0f);
    const float r = (b < 0.0f) ? __double2float_rn(__dadd_rn((double)q, -0.5)) : __double2float_rn(__dadd_rn((double)q, 0.5));
    corr[gp] = (int8_t)(int)r;
}

// QueryMatcher::match (QueryMatcher.cpp:226-244): per k-mer start, threshold = max(kmerThr - round(sum bias), 0)
__global__ void kmerThresholds(const uint8_t *__restrict__ code, const uint32_t *__restrict__ off, uint32_t nQ,
                               uint32_t totalRes, const float *__restrict__ bias, int kmerThr,
                               int16_t *__restrict__ thr) {
    uint32_t gp = blockIdx.x * blockDim.x + threadIdx.x;
    if (gp >= totalRes) return;
    uint32_t s = seqOfPosition(off, nQ, gp);
    uint32_t p = gp - off[s];
    uint32_t L = off[s + 1] - off[s];
    int16_t out = -1;
    if (p + SPAN <= L) {
        bool ok = true;
        float acc = 0.0f;
#pragma unroll
        for (int k = 0; k < KMER; k++) {
            ok = ok && (code[gp + c_seedPos[k]] < 20);
            acc = __fadd_rn(acc, bias[gp + c_seedPos[k]]);
        }
        if (ok) {
            double adj = (acc < 0.0f) ? __dadd_rn((double)acc, -0.5) : __dadd_rn((double)acc, 0.5);
            short rb = (short)(int)adj;
            int t = kmerThr - rb;
            out = (int16_t)(t > 0 ? t : 0);
        }
    }
    thr[gp] = out;
}

// ------------------------------------------------------------------------------------------------
// the prefilter kernel
// ------------------------------------------------------------------------------------------------
constexpr int PF_WARPS = 8;              // warps per CTA
constexpr int PF_CAP2 = 1024;            // staged prefix of the second 3-mer's neighbour row
constexpr int PF_UNROLL = 4;             // probe steps in flight per warp (independent bitmap loads)
constexpr uint32_t PF_Q1CAP = 256;       // ring 1 (non-empty buckets): up to 32*PF_UNROLL pushes between drains
constexpr uint32_t PF_Q2CAP = 128;       // ring 2 (index hits)
constexpr uint32_t PF_QDRAIN = 96;       // a stage runs when its ring holds this many: popped items were prefetched >= 64 pushes ago
constexpr size_t PF_DYN_SMEM = (size_t)PF_WARPS * (PF_Q1CAP + PF_Q2CAP) * (sizeof(uint32_t) + sizeof(uint16_t));
static_assert(PF_QDRAIN - 1 + 32 * PF_UNROLL <= PF_Q1CAP && PF_QDRAIN - 1 + 32 <= PF_Q2CAP, "ring capacity");

__device__ __forceinline__ void prefetchL2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// L2 cache policies: the streamed, never re-used data (bucket rows, entries, the per-warp tables) is marked evict-first so
// that it does not push the re-used data (occupancy bitmap, count table, target residues) out of the L2.
__device__ __forceinline__ uint64_t l2Policy(int evictFirst) {
    uint64_t pol;
    if (evictFirst) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2PolicyKeep() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint32_t ldKeep(const uint32_t *p, uint64_t pol) {
    uint32_t v;
    asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ uint32_t ldStream(const uint32_t *p, uint64_t pol) {
    uint32_t v;
    asm volatile("ld.global.cg.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol) : "memory");
    return v;
}
__device__ __forceinline__ uint2 ldStream2(const uint2 *p, uint64_t pol) {
    uint2 v;
    asm volatile("ld.global.cg.L2::cache_hint.v2.u32 {%0, %1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(pol) : "memory");
    return v;
}
__device__ __forceinline__ void stStream(uint32_t *p, uint32_t v, uint64_t pol) {
    asm volatile("st.global.cg.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(pol) : "memory");
}

struct PrefilterArgs {
    // target side
    const uint8_t *tRes;
    const uint32_t *tOff;
    uint32_t nT;
    const uint32_t *bmBits;    // per 32 k-mers: occupancy bits (8 MB, L2 resident)
    const uint32_t *bmRank;    // per 32 k-mers: row of the first non-empty one
    const uint2 *rows;         // per non-empty k-mer: {target, 1<<31 | pos} (single entry) or {start, length}
    const uint2 *ent;          // bucket entries: (target id, first position of the k-mer in that target)
    const int16_t *e3score;
    const uint16_t *e3index;
    const uint16_t *e3cnt;     // [8000][512] cumulative counts: e3cnt[x*512 + 255 - v] = #{y : score(x,y) >= v}
    const int8_t *diagSub;     // 21x21 BLOSUM62 x2 (-0.2)
    // query side
    QueryBatchDev q;
    uint32_t nQ;
    const uint32_t *order;     // queries, longest first
    uint32_t queryFirst;       // identity = queryFirst + q when sameDb
    int sameDb;
    // parameters
    uint32_t maxHits;          // min(max_seqs, nT)
    uint32_t minDiagScore;
    float covThr;
    int covMode;
    uint32_t binMask;
    // scratch
    uint32_t *table;           // per warp: nT words
    uint32_t candCap;
    uint64_t *cand;            // per warp: candCap   (t<<32 | diag<<16 | score<<8)
    uint64_t *keep;            // per warp: candCap   (t<<32 | diag<<16 | score<<8), emission order kept
    uint32_t *resT;            // per warp: maxHits
    int32_t *resScore;
    uint16_t *resDiag;
    // output
    sd_prefilter_hit *out;     // nQ * maxHits
    uint32_t *outCount;        // nQ
    unsigned long long *stats; // [0]=probes [1]=matches [2]=diagResidues [3]=hits [4]=candidates [5]=truncated [6]=errors
    uint32_t *nextQuery;
    uint32_t *errQuery;
    uint32_t *failList;        // queries whose candidate list did not fit candCap (re-run with a larger list)
    uint32_t *nFail;
    int evictFirst;            // L2 evict-first policy for the streamed data
    int prefetch;              // bit 0: table words, bit 1: bucket rows
    int debugSkip;             // profiling only (SD_PF_SKIP): bit 0 drops stage C, bit 1 drops stage B -- results are then wrong
};

enum { ST_PROBES = 0, ST_MATCHES, ST_DIAGRES, ST_HITS, ST_CANDS, ST_TRUNC, ST_ERR_CAND, ST_ERR_RESCORE, ST_N };

// UngappedAlignment::computeSingelSequenceScores + scalarDiagonalScoring (UngappedAlignment.cpp:30-43,416-430)
__device__ int diagonalScore(const PrefilterArgs &a, const int8_t *sSub, uint32_t qBase, int Lq, uint32_t t, int diag,
                             uint32_t &residues) {
    const uint32_t tb = a.tOff[t];
    const int Lt = (int)(a.tOff[t + 1] - tb);
    int qs, ts, len;
    if (diag >= 0) {
        if (diag >= Lq) return 0;
        qs = diag; ts = 0; len = min(Lt, Lq - diag);
    } else {
        if (-diag >= Lt) return 0;
        qs = 0; ts = -diag; len = min(Lt + diag, Lq);
    }
    const uint8_t *qc = a.q.code + qBase + qs;
    const int8_t *qk = a.q.corr + qBase + qs;
    const uint8_t *tc = a.tRes + tb + ts;
    int best = 0, run = 0;
    for (int p = 0; p < len; p++) {
        run += (int)sSub[qc[p] * AA + tc[p]] + (int)qk[p];
        run = max(run, 0);
        best = max(best, run);
    }
    residues += (uint32_t)len;
    return best;
}

// same score, computed by a whole warp: lane l takes residues l, l+32, ...; with S the prefix sums of the per-residue scores,
// run_i = S_i - min(0, min_{j<=i} S_j) and the result is max_i run_i
__device__ __forceinline__ int diagonalScoreWarp(const PrefilterArgs &a, const int8_t *sSub, uint32_t qBase, int Lq, uint32_t t,
                                                 int diag, uint32_t lane, unsigned long long &residues) {
    const uint32_t tb = __ldg(&a.tOff[t]);
    const int Lt = (int)(__ldg(&a.tOff[t + 1]) - tb);
    int qs, ts, len;
    if (diag >= 0) {
        if (diag >= Lq) return 0;
        qs = diag; ts = 0; len = min(Lt, Lq - diag);
    } else {
        if (-diag >= Lt) return 0;
        qs = 0; ts = -diag; len = min(Lt + diag, Lq);
    }
    const uint8_t *qc = a.q.code + qBase + qs;
    const int8_t *qk = a.q.corr + qBase + qs;
    const uint8_t *tc = a.tRes + tb + ts;
    int best = 0, carrySum = 0, carryMin = 0;
    for (int p0 = 0; p0 < len; p0 += 32) {
        const int p = p0 + (int)lane;
        int x = 0;
        if (p < len) x = (int)sSub[__ldg(&qc[p]) * AA + __ldg(&tc[p])] + (int)__ldg(&qk[p]);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= (uint32_t)d) x += y;
        }
        x += carrySum;
        int m = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, m, d);
            if (lane >= (uint32_t)d) m = min(m, y);
        }
        m = min(m, carryMin);
        best = max(best, x - m);
        carrySum = __shfl_sync(0xffffffffu, x, 31);
        carryMin = __shfl_sync(0xffffffffu, m, 31);
    }
    if (lane == 0) residues += (unsigned long long)len;
    return __reduce_max_sync(0xffffffffu, best);
}

__global__ void __launch_bounds__(PF_WARPS * 32, 4) prefilterKernel(PrefilterArgs a) {
    __shared__ uint16_t sX2[PF_WARPS][PF_CAP2];
    __shared__ uint32_t sHist[PF_WARPS][256];
    __shared__ int8_t sSub[AA * AA];
    extern __shared__ __align__(16) uint8_t sDyn[];    // stage rings: [warp][Q1c | Q2t | Q1p | Q2d]

    for (int i = threadIdx.x; i < AA * AA; i += blockDim.x) sSub[i] = a.diagSub[i];
    __syncthreads();

    const uint32_t lane = laneId();
    const uint32_t wInBlock = threadIdx.x >> 5;
    const uint32_t warp = blockIdx.x * PF_WARPS + wInBlock;
    const uint32_t ltMask = (1u << lane) - 1u;
    uint32_t *tab = a.table + (size_t)warp * a.nT;
    uint64_t *cand = a.cand + (size_t)warp * a.candCap;
    uint64_t *keep = a.keep + (size_t)warp * a.candCap;
    uint32_t *resT = a.resT + (size_t)warp * a.maxHits;
    int32_t *resScore = a.resScore + (size_t)warp * a.maxHits;
    uint16_t *resDiag = a.resDiag + (size_t)warp * a.maxHits;
    uint16_t *X2 = sX2[wInBlock];
    uint32_t *hist = sHist[wInBlock];
    uint32_t *Q1c = reinterpret_cast<uint32_t *>(sDyn) + (size_t)wInBlock * (PF_Q1CAP + PF_Q2CAP);
    uint32_t *Q2t = Q1c + PF_Q1CAP;
    uint16_t *Q1p = reinterpret_cast<uint16_t *>(sDyn + (size_t)PF_WARPS * (PF_Q1CAP + PF_Q2CAP) * sizeof(uint32_t)) + (size_t)wInBlock * (PF_Q1CAP + PF_Q2CAP);
    uint16_t *Q2d = Q1p + PF_Q1CAP;

    const uint64_t pol = l2Policy(a.evictFirst);
    const uint64_t polKeep = a.evictFirst ? l2PolicyKeep() : l2Policy(0);
    uint32_t epoch = 0;   // table words start zeroed; live epochs are 1..255

    for (;;) {
        uint32_t qi = 0;
        if (lane == 0) qi = atomicAdd(a.nextQuery, 1u);
        qi = __shfl_sync(0xffffffffu, qi, 0);
        if (qi >= a.nQ) break;
        // per-query counters: added to the global statistics only when the query completes (a query whose candidate
        // list overflows is re-run by a second launch and counted there)
        unsigned long long stProbes = 0, stMatches = 0, stDiagRes = 0, stHits = 0, stCands = 0, stTrunc = 0;
        const uint32_t q = a.order[qi];
        const uint32_t qBase = a.q.off[q];
        const int Lq = (int)(a.q.off[q + 1] - qBase);
        const uint32_t identity = a.sameDb ? (a.queryFirst + q) : 0xffffffffu;

        // two epochs per query: E for the hit stream, E+1 for the best-diagonal pass
        if (epoch + 2 > 255) {
            for (uint32_t i = lane; i < a.nT; i += 32) __stcg(&tab[i], 0u);
            __syncwarp();
            epoch = 0;
        }
        const uint32_t E = epoch + 1;
        epoch += 2;

        uint32_t nCand = 0;
        bool candOverflow = false;

        // ---------------- hit stream: positions -> similar k-mers -> buckets -> double-hit rule
        // Three stages, decoupled by two small per-warp FIFO rings in shared memory so that every stage runs with full
        // lanes and every HBM access was prefetched into L2 one stage earlier (stream order is the FIFO order):
        //   A  32 similar k-mers per step -> occupancy bitmap + rank (L2 resident) -> non-empty ones to ring 1
        //   B  32 buckets per step: compact offset row (prefetched by A) -> entries -> (target, diagonal) to ring 2
        //   C  32 index hits per step: double-diagonal rule on the per-warp table (prefetched by B)
        uint32_t q1h = 0, q1n = 0, q2h = 0, q2n = 0;

        auto stageC = [&](const uint32_t n) {
            __syncwarp();
            const bool validE = lane < n;
            uint32_t t = 0x80000000u | lane;     // unique dummy for idle lanes
            uint32_t diag16 = 0;
            if (validE) {
                const uint32_t slot = (q2h + lane) & (PF_Q2CAP - 1);
                t = Q2t[slot];
                diag16 = Q2d[slot];
            }
            q2h += n;
            q2n -= n;
            if (a.debugSkip & 1) return;
            const uint32_t d8 = diag16 & 0xffu;
            // --- CacheFriendlyOperations::findDuplicates, restated per target in stream order
            const uint32_t grp = __match_any_sync(0xffffffffu, t);
            uint32_t old = 0;
            if (validE) old = ldStream(&tab[t], pol);
            if ((old >> 24) != E) old = 0;
            const uint32_t lowerG = grp & ltMask;
            const int prevLane = lowerG ? (31 - __clz(lowerG)) : -1;
            uint32_t prevD8 = __shfl_sync(0xffffffffu, d8, prevLane >= 0 ? prevLane : (int)lane);
            if (prevLane < 0) prevD8 = (old >> 16) & 0xffu;
            const bool emit = validE && (d8 == prevD8);
            const uint32_t emitBallot = __ballot_sync(0xffffffffu, emit);
            const uint32_t grpE = grp & emitBallot;
            const uint32_t lowerE = grpE & ltMask;
            const int prevELane = lowerE ? (31 - __clz(lowerE)) : -1;
            uint32_t prevED8 = __shfl_sync(0xffffffffu, d8, prevELane >= 0 ? prevELane : (int)lane);
            bool prevEExists = true;
            if (prevELane < 0) { prevED8 = old & 0xffu; prevEExists = ((old >> 8) & 1u) != 0; }
            const bool isCand = emit && !(prevEExists && prevED8 == d8);
            const int lastELane = grpE ? (31 - __clz(grpE)) : -1;
            const uint32_t lastED8 = __shfl_sync(0xffffffffu, d8, lastELane >= 0 ? lastELane : (int)lane);
            if (validE && (grp >> lane) == 1u) {
                uint32_t w = (E << 24) | (d8 << 16);
                if (lastELane >= 0) w |= (1u << 8) | lastED8;
                else w |= old & 0x1ffu;
                stStream(&tab[t], w, pol);
            }
            const uint32_t candBallot = __ballot_sync(0xffffffffu, isCand);
            if (isCand) {
                const uint32_t pos = nCand + __popc(candBallot & ltMask);
                if (pos < a.candCap) cand[pos] = ((uint64_t)t << 32) | ((uint64_t)diag16 << 16);
            }
            nCand += __popc(candBallot);
        };

        auto stageB = [&](const uint32_t n) {
            __syncwarp();
            uint32_t rx = 0, ry = 0, m = 0, pq = 0;
            if (a.debugSkip & 2) { q1h += n; q1n -= n; return; }
            if (lane < n) {
                const uint32_t slot = (q1h + lane) & (PF_Q1CAP - 1);
                const uint2 row = ldStream2(&a.rows[Q1c[slot]], pol);
                pq = Q1p[slot];
                rx = row.x;
                ry = row.y;
                m = (ry >> 31) ? 1u : ry;                // >= 1: only non-empty buckets are queued
            }
            q1h += n;
            q1n -= n;
            uint32_t totalM;
            const uint32_t exclM = warpExclusiveScan(m, totalM);
            stMatches += (lane == 0) ? totalM : 0;
            for (uint32_t ec = 0; ec < totalM; ec += 32) {
                const uint32_t eidx = ec + lane;
                const bool validE = eidx < totalM;
                const uint32_t ownerM = warpOwnerOfSlot(exclM, m != 0, ec, lane);
                const uint32_t oExcl = __shfl_sync(0xffffffffu, exclM, ownerM);
                const uint32_t oX = __shfl_sync(0xffffffffu, rx, ownerM);
                const uint32_t oY = __shfl_sync(0xffffffffu, ry, ownerM);
                const uint32_t oP = __shfl_sync(0xffffffffu, pq, ownerM);
                if (validE) {
                    uint2 en = make_uint2(oX, oY & 0xffffu);
                    if (!(oY >> 31)) en = ldStream2(&a.ent[oX + (eidx - oExcl)], pol);
                    if (a.prefetch & 1) prefetchL2(&tab[en.x]);
                    const uint32_t slot = (q2h + q2n + lane) & (PF_Q2CAP - 1);
                    Q2t[slot] = en.x;
                    Q2d[slot] = (uint16_t)(oP - en.y);
                }
                q2n += min(32u, totalM - ec);
                while (q2n >= PF_QDRAIN) stageC(32u);
            }
        };

        for (int p0 = 0; p0 + SPAN <= Lq; p0 += 32) {
            // ---- set-up of 32 positions at once (one per lane): thresholds, 3-mer rows, reachable prefix of row 2
            int thrL = -1, top2L = 0;
            uint32_t i1L = 0, i2L = 0, nb2L = 0;
            {
                const int pl = p0 + (int)lane;
                if (pl + SPAN <= Lq) thrL = a.q.thr[qBase + pl];
                if (thrL >= 0) {
                    const uint8_t *qc = a.q.code + qBase + pl;
                    i1L = qc[0] + 20u * qc[1] + 400u * qc[3];
                    i2L = qc[5] + 20u * qc[8] + 400u * qc[9];
                    const int top1 = (int)__ldg(&a.e3score[(size_t)i1L * N3]);
                    top2L = (int)__ldg(&a.e3score[(size_t)i2L * N3]);
                    if (top1 < thrL - top2L) thrL = -1;           // KmerGenerator.cpp:129-137: nothing reaches the threshold
                    else {
                        nb2L = countGE(a.e3cnt, i2L, thrL - top1);
                        if (nb2L == 0) thrL = -1;
                        else {
                            prefetchL2(&a.e3index[(size_t)i1L * N3]);
                            const uint16_t *x2 = a.e3index + (size_t)i2L * N3;
                            for (uint32_t c = 0; c < nb2L && c < (uint32_t)PF_CAP2; c += 64) prefetchL2(&x2[c]);
                        }
                    }
                }
            }
            uint32_t active = __ballot_sync(0xffffffffu, thrL >= 0);
            while (active) {
                const int src = __ffs(active) - 1;
                active &= active - 1;
                const int p = p0 + src;
                const int thr = __shfl_sync(0xffffffffu, thrL, src);
                const uint32_t i1 = __shfl_sync(0xffffffffu, i1L, src);
                const uint32_t i2 = __shfl_sync(0xffffffffu, i2L, src);
                const uint32_t nb2 = __shfl_sync(0xffffffffu, nb2L, src);
                const int cutoff1 = thr - __shfl_sync(0xffffffffu, top2L, src);
                const int16_t *s1 = a.e3score + (size_t)i1 * N3;
                const uint16_t *x1 = a.e3index + (size_t)i1 * N3;
                const uint16_t *x2 = a.e3index + (size_t)i2 * N3;
                // stage the prefix of row 2 that any first-row element can reach
                __syncwarp();
                for (uint32_t c = lane; c < nb2 && c < (uint32_t)PF_CAP2; c += 32) X2[c] = __ldg(&x2[c]);
                __syncwarp();

                for (uint32_t a0 = 0; a0 < N3; a0 += 32) {
                    const uint32_t ai = a0 + lane;
                    const int sa = (ai < N3) ? (int)__ldg(&s1[ai]) : -32768;
                    const bool okA = sa >= cutoff1;
                    const uint32_t okBallot = __ballot_sync(0xffffffffu, okA);
                    if (okBallot == 0) break;
                    uint32_t nb = 0;
                    uint32_t xa = 0;
                    if (okA) {
                        xa = __ldg(&x1[ai]);
                        nb = countGE(a.e3cnt, i2, thr - sa);      // KmerGenerator.cpp:201: second-row elements with s2 >= thr - s1
                    }
                    // nb is non-increasing over the lanes (row 1 is sorted), zeros only at the tail
                    uint32_t totalProbes;
                    const uint32_t exclA = warpExclusiveScan(nb, totalProbes);
                    stProbes += (lane == 0) ? totalProbes : 0;

                    for (uint32_t pc = 0; pc < totalProbes; pc += 32 * PF_UNROLL) {
                        uint32_t w[PF_UNROLL], kmerW[PF_UNROLL];
#pragma unroll
                        for (int j = 0; j < PF_UNROLL; j++) {
                            const uint32_t base = pc + 32u * j;
                            w[j] = 0;
                            kmerW[j] = 0;
                            if (base < totalProbes) {
                                const uint32_t ownerA = warpOwnerOfSlot(exclA, nb != 0, base, lane);
                                const uint32_t ownerExcl = __shfl_sync(0xffffffffu, exclA, ownerA);
                                const uint32_t ownerX = __shfl_sync(0xffffffffu, xa, ownerA);
                                if (base + lane < totalProbes) {
                                    const uint32_t c = base + lane - ownerExcl;
                                    const uint32_t xc = (c < (uint32_t)PF_CAP2) ? (uint32_t)X2[c] : (uint32_t)__ldg(&x2[c]);
                                    kmerW[j] = ownerX + N3 * xc;            // KmerGenerator.cpp:207
                                    w[j] = ldKeep(&a.bmBits[kmerW[j] >> 5], polKeep);
                                }
                            }
                        }
                        uint32_t rk[PF_UNROLL];
#pragma unroll
                        for (int j = 0; j < PF_UNROLL; j++) {
                            rk[j] = 0;
                            if ((w[j] >> (kmerW[j] & 31u)) & 1u) rk[j] = __ldg(&a.bmRank[kmerW[j] >> 5]);   // idle lanes carry w = 0
                        }
#pragma unroll
                        for (int j = 0; j < PF_UNROLL; j++) {
                            if (pc + 32u * j < totalProbes) {
                                const uint32_t bit = kmerW[j] & 31u;
                                const bool has = ((w[j] >> bit) & 1u) != 0;
                                const uint32_t ci = rk[j] + __popc(w[j] & ((1u << bit) - 1u));
                                if (has && (a.prefetch & 2)) prefetchL2(&a.rows[ci]);
                                const uint32_t hb = __ballot_sync(0xffffffffu, has);
                                if (has) {
                                    const uint32_t slot = (q1h + q1n + __popc(hb & ltMask)) & (PF_Q1CAP - 1);
                                    Q1c[slot] = ci;
                                    Q1p[slot] = (uint16_t)p;
                                }
                                q1n += __popc(hb);
                            }
                        }
                        while (q1n >= PF_QDRAIN) stageB(32u);
                    }
                    if (okBallot != 0xffffffffu) break;
                }
            }
        }
        while (q1n > 0) stageB(min(q1n, 32u));
        while (q2n > 0) stageC(min(q2n, 32u));
        __syncwarp();
        if (nCand > a.candCap) { candOverflow = true; nCand = a.candCap; }
        stCands += (lane == 0) ? nCand : 0;
        __syncwarp();

        // ---------------- diagonal scores (UngappedAlignment::align), capped at 255
        // The whole warp scores one diagonal at a time: lanes take consecutive residues (coalesced reads of query and
        // target), the running max(0, run + s) becomes prefix sums minus their running minimum (warp scans).
        for (uint32_t c0 = 0; c0 < nCand; c0 += 32) {
            const uint32_t c = c0 + lane;
            const uint64_t v = (c < nCand) ? cand[c] : 0ull;
            const uint32_t nHere = min(32u, nCand - c0);
            int myScore = 0;
            for (uint32_t j = 0; j < nHere; j++) {
                const uint64_t vj = __shfl_sync(0xffffffffu, v, j);
                const int sc = diagonalScoreWarp(a, sSub, qBase, Lq, (uint32_t)(vj >> 32), (int)(int16_t)(uint16_t)(vj >> 16), lane, stDiagRes);
                if (lane == j) myScore = sc;
            }
            if (c < nCand) {
                const int sc = min(myScore, 255);
                cand[c] = v | ((uint64_t)sc << 8);
                // best score per target: epoch E+1 words order above every epoch-E word
                atomicMax(&tab[(uint32_t)(v >> 32)], ((E + 1) << 24) | (uint32_t)sc);
            }
        }
        __threadfence_block();
        __syncwarp();

        // ---------------- keepMaxScoreElementOnly: first candidate (stream order) carrying the target's max
        for (uint32_t i = lane; i < 256; i += 32) hist[i] = 0;
        __syncwarp();
        uint32_t nKeep = 0;
        for (uint32_t c0 = 0; c0 < nCand; c0 += 32) {
            const uint32_t c = c0 + lane;
            const bool valid = c < nCand;
            uint64_t v = 0;
            uint32_t t = 0x80000000u | lane;
            if (valid) { v = cand[c]; t = (uint32_t)(v >> 32); }
            const uint32_t s = (uint32_t)(v >> 8) & 0xffu;
            uint32_t w = 0;
            if (valid) w = __ldcg(&tab[t]);
            const bool isMax = valid && ((w & 0x100u) == 0) && ((w & 0xffu) == s);
            const uint32_t grp = __match_any_sync(0xffffffffu, t);
            const uint32_t maxBallot = __ballot_sync(0xffffffffu, isMax);
            const uint32_t grpMax = grp & maxBallot;
            const bool winner = isMax && ((grpMax & ltMask) == 0);
            if (winner) __stcg(&tab[t], w | 0x100u);
            const uint32_t winBallot = __ballot_sync(0xffffffffu, winner);
            if (winner) {
                keep[nKeep + __popc(winBallot & ltMask)] = v;
                atomicAdd(&hist[s], 1u);
            }
            nKeep += __popc(winBallot);
            __syncwarp();
        }

        // ---------------- score threshold (QueryMatcher::computeScoreThreshold) and selection
        uint32_t thrScore = 0;
        if (lane == 0) {
            uint32_t found = 0, sc;
            for (sc = 255; sc > 0; sc--) {
                found += hist[sc];
                if (found >= a.maxHits) break;
            }
            thrScore = max(sc, a.minDiagScore);
        }
        thrScore = __shfl_sync(0xffffffffu, thrScore, 0);
        const bool needRescale = thrScore >= 255u;   // >= max-seqs elements at the 8-bit cap: rescoreHits path below
        uint32_t nRes = 0;
        if (identity != 0xffffffffu) {
            if (lane == 0) { resT[0] = identity; resScore[0] = 65535; resDiag[0] = 0; }
            nRes = 1;
        }
        bool dropped = false;
        if (needRescale) {
            // scoreIsTruncated (QueryMatcher.cpp:157-171): the cap no longer orders the top class.  rescoreHits (:525-544):
            // uncapped diagonal score minus 255, rescaled to 0..255 against the query's self score; stable re-sort by that
            // value; getResult maps it back with rescaleScore = maxSelf (:407-410).  Rare, so plain lane-parallel code.
            // 1. the capped class in the reference's order (bin-major, arrival order inside a bin) -> cand[] (free by now)
            uint32_t nL = 0;
            for (uint32_t bin = 0; bin <= a.binMask; bin++) {
                for (uint32_t c0 = 0; c0 < nKeep; c0 += 32) {
                    const uint32_t c = c0 + lane;
                    uint64_t v = 0;
                    bool sel = false;
                    if (c < nKeep) {
                        v = keep[c];
                        sel = ((uint32_t)(v >> 8) & 0xffu) == 255u && ((uint32_t)(v >> 32) & a.binMask) == bin;
                    }
                    const uint32_t b = __ballot_sync(0xffffffffu, sel);
                    if (sel) cand[nL + __popc(b & ltMask)] = v & ~0xffffull;      // (target, diagonal), score byte cleared
                    nL += __popc(b);
                }
            }
            __syncwarp();
            // 2. self score of the query on the main diagonal (scoreSingleSequence(query, 0, 0)), uncapped
            int maxSelf = 0;
            if (lane == 0) {
                const uint8_t *qc = a.q.code + qBase;
                const int8_t *qk = a.q.corr + qBase;
                int best = 0, run = 0;
                for (int p = 0; p < Lq; p++) {
                    run += (int)sSub[qc[p] * AA + qc[p]] + (int)qk[p];
                    run = max(run, 0);
                    best = max(best, run);
                }
                maxSelf = min(max(best - 255, 1), 65535);
            }
            maxSelf = __shfl_sync(0xffffffffu, maxSelf, 0);
            const float fltMaxSelf = (float)maxSelf;
            // 3. rescaled 8-bit scores + their histogram
            for (uint32_t i = lane; i < 256; i += 32) hist[i] = 0;
            __syncwarp();
            for (uint32_t c = lane; c < nL; c += 32) {
                const uint64_t v = cand[c];
                uint32_t dummy = 0;
                const uint32_t full = (uint32_t)diagonalScore(a, sSub, qBase, Lq, (uint32_t)(v >> 32), (int)(int16_t)(uint16_t)(v >> 16), dummy);
                const uint32_t newScore = full - 255u;
                const float sc = (float)min(newScore, 65535u);
                const float scaled = __fmul_rn(__fdiv_rn(sc, fltMaxSelf), 255.0f);
                const uint32_t cnt = (uint32_t)(int)__dadd_rn((double)scaled, 0.5) & 0xffu;
                cand[c] = v | (uint64_t)cnt;
                atomicAdd(&hist[cnt], 1u);
            }
            __syncwarp();
            // 4. stable selection by rescaled score, highest first
            for (int s2 = 255; s2 >= 0; s2--) {
                if (hist[s2] == 0) continue;
                for (uint32_t c0 = 0; c0 < nL; c0 += 32) {
                    const uint32_t c = c0 + lane;
                    uint64_t v = 0;
                    bool sel = false;
                    if (c < nL) {
                        v = cand[c];
                        sel = (uint32_t)(v & 0xffu) == (uint32_t)s2 && (uint32_t)(v >> 32) != identity;
                    }
                    const uint32_t b = __ballot_sync(0xffffffffu, sel);
                    if (sel) {
                        const uint32_t pos = nRes + __popc(b & ltMask);
                        if (pos < a.maxHits) {
                            resT[pos] = (uint32_t)(v >> 32);
                            resScore[pos] = (int)(255u + ((uint32_t)s2 * (uint32_t)maxSelf) / 255u);
                            resDiag[pos] = (uint16_t)(v >> 16);
                        }
                    }
                    nRes += __popc(b);
                    if (nRes > a.maxHits) { dropped = true; nRes = a.maxHits; }
                }
            }
            __syncwarp();
        } else {
        // everything strictly above the threshold survives
        for (uint32_t c0 = 0; c0 < nKeep; c0 += 32) {
            const uint32_t c = c0 + lane;
            uint64_t v = 0;
            bool sel = false;
            if (c < nKeep) {
                v = keep[c];
                const uint32_t s = (uint32_t)(v >> 8) & 0xffu;
                sel = s > thrScore && (uint32_t)(v >> 32) != identity;
            }
            const uint32_t b = __ballot_sync(0xffffffffu, sel);
            if (sel) {
                const uint32_t pos = nRes + __popc(b & ltMask);
                if (pos < a.maxHits) {
                    resT[pos] = (uint32_t)(v >> 32);
                    resScore[pos] = (int)((v >> 8) & 0xffu);
                    resDiag[pos] = (uint16_t)(v >> 16);
                }
            }
            nRes += __popc(b);
        }
        nRes = min(nRes, a.maxHits);
        // the threshold class fills what is left, in (id & binMask, stream order) order
        for (uint32_t bin = 0; bin <= a.binMask; bin++) {
            for (uint32_t c0 = 0; c0 < nKeep; c0 += 32) {
                const uint32_t c = c0 + lane;
                uint64_t v = 0;
                bool sel = false;
                if (c < nKeep) {
                    v = keep[c];
                    const uint32_t s = (uint32_t)(v >> 8) & 0xffu;
                    const uint32_t t = (uint32_t)(v >> 32);
                    sel = s == thrScore && t != identity && (t & a.binMask) == bin;
                }
                const uint32_t b = __ballot_sync(0xffffffffu, sel);
                if (sel) {
                    const uint32_t pos = nRes + __popc(b & ltMask);
                    if (pos < a.maxHits) {
                        resT[pos] = (uint32_t)(v >> 32);
                        resScore[pos] = (int)((v >> 8) & 0xffu);
                        resDiag[pos] = (uint16_t)(v >> 16);
                    }
                }
                nRes += __popc(b);
                if (nRes > a.maxHits) { dropped = true; nRes = a.maxHits; }
            }
        }
        }
        __syncwarp();
        if (dropped) stTrunc += (lane == 0) ? 1 : 0;
        // scores that hit the 8-bit cap are replaced by the uncapped diagonal score (QueryMatcher.cpp:410-413)
        for (uint32_t r = lane; r < nRes && !needRescale; r += 32) {
            if (resScore[r] >= 255 && resScore[r] != 65535) {
                uint32_t dummy = 0;
                resScore[r] = diagonalScore(a, sSub, qBase, Lq, resT[r], (int)(int16_t)resDiag[r], dummy);
            }
        }
        __syncwarp();
        stHits += (lane == 0) ? nRes : 0;

        // ---------------- final order (score desc, id asc; identity stays first) + coverage pre-filter
        const uint32_t first = (identity != 0xffffffffu) ? 1u : 0u;
        sd_prefilter_hit *outQ = a.out + (size_t)q * a.maxHits;
        uint32_t nOut = 0;
        // rank sort: rank = number of elements that come before
        for (uint32_t r0 = 0; r0 < nRes; r0 += 32) {
            const uint32_t r = r0 + lane;
            uint32_t rank = 0xffffffffu;
            uint32_t t = 0;
            int s = 0;
            uint16_t dg = 0;
            if (r < nRes) {
                t = resT[r]; s = resScore[r]; dg = resDiag[r];
                if (r < first) rank = 0;
                else {
                    rank = first;
                    for (uint32_t j = first; j < nRes; j++) {
                        const int sj = resScore[j];
                        const uint32_t tj = resT[j];
                        rank += (sj > s || (sj == s && tj < t)) ? 1u : 0u;
                    }
                }
            }
            // coverage pre-filter (Prefiltering.cpp:856-863, Util::canBeCovered Util.cpp:477-494)
            bool pass = r < nRes;
            if (pass && a.covThr > 0.0f && (a.covMode == 0 || a.covMode == 2 || a.covMode == 5)) {
                const float ql = (float)Lq;
                const float tl = (float)(a.tOff[t + 1] - a.tOff[t]);
                if (a.covMode == 0) pass = (__fdiv_rn(ql, tl) >= a.covThr) && (__fdiv_rn(tl, ql) >= a.covThr);
                else if (a.covMode == 2) pass = __fdiv_rn(tl, ql) >= a.covThr;
                else pass = __fdiv_rn(fminf(tl, ql), fmaxf(tl, ql)) >= a.covThr;
            }
            if (r < nRes) {
                // park (rank, pass) in the scratch result arrays' tail: write straight to the output slot by rank
                outQ[rank].query = q;
                outQ[rank].target = t;
                outQ[rank].score = pass ? s : -1;      // -1 marks "filtered by coverage"; compacted below
                outQ[rank].diag = (int16_t)dg;
                outQ[rank].pad = 0;
            }
        }
        __syncwarp();
        __threadfence_block();
        // compact the coverage-filtered entries in place, keeping order
        for (uint32_t r0 = 0; r0 < nRes; r0 += 32) {
            const uint32_t r = r0 + lane;
            sd_prefilter_hit h;
            bool ok = false;
            if (r < nRes) { h = outQ[r]; ok = h.score >= 0; }
            const uint32_t b = __ballot_sync(0xffffffffu, ok);
            __syncwarp();
            if (ok) outQ[nOut + __popc(b & ltMask)] = h;
            nOut += __popc(b);
            __syncwarp();
        }
        // diagResidues is accumulated per lane
        for (int d = 16; d > 0; d >>= 1) stDiagRes += __shfl_down_sync(0xffffffffu, stDiagRes, d);
        if (lane == 0) {
            if (candOverflow) {
                a.outCount[q] = 0;
                a.failList[atomicAdd(a.nFail, 1u)] = q;
                atomicAdd(&a.stats[ST_ERR_CAND], 1ull);
                atomicMin(a.errQuery, q);
            } else {
                a.outCount[q] = nOut;
                atomicAdd(&a.stats[ST_PROBES], stProbes);
                atomicAdd(&a.stats[ST_MATCHES], stMatches);
                atomicAdd(&a.stats[ST_DIAGRES], stDiagRes);
                atomicAdd(&a.stats[ST_HITS], stHits);
                atomicAdd(&a.stats[ST_CANDS], stCands);
                if (stTrunc) atomicAdd(&a.stats[ST_TRUNC], stTrunc);
            }
        }
        __syncwarp();
    }
}

// dense[offs[q] + i] = out[q*maxHits + i]; one warp per query
__global__ void compactHits(const sd_prefilter_hit *__restrict__ out, const uint32_t *__restrict__ count,
                            const uint32_t *__restrict__ offs, uint32_t nQ, uint32_t maxHits,
                            sd_prefilter_hit *__restrict__ dense) {
    const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= nQ) return;
    const uint32_t n = count[q];
    const sd_prefilter_hit *src = out + (size_t)q * maxHits;
    sd_prefilter_hit *dst = dense + offs[q];
    for (uint32_t i = laneId(); i < n; i += 32) dst[i] = src[i];
}

}  // namespace sd

// =================================================================================================
// host side
// =================================================================================================
using namespace sd;

static inline unsigned gridFor(uint64_t n, unsigned block) { return (unsigned)((n + block - 1) / block); }

static int ensureExt3(sd_context *ctx) {
    if (ctx->haveExt3) return SD_OK;
    std::vector<int16_t> sc;
    std::vector<uint16_t> ix, cn;
    buildExt3(ctx->seedMat, sc, ix, cn);
    int rc;
    if ((rc = ctx->e3score.alloc(sc.size())) != SD_OK) return rc;
    if ((rc = ctx->e3index.alloc(ix.size())) != SD_OK) return rc;
    if ((rc = ctx->e3cnt.alloc(cn.size())) != SD_OK) return rc;
    SD_CUDA(cudaMemcpyAsync(ctx->e3cnt.p, cn.data(), cn.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, ctx->stream));
    SD_CUDA(cudaMemcpyAsync(ctx->e3score.p, sc.data(), sc.size() * sizeof(int16_t), cudaMemcpyHostToDevice, ctx->stream));
    SD_CUDA(cudaMemcpyAsync(ctx->e3index.p, ix.data(), ix.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, ctx->stream));
    SD_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->haveExt3 = true;
    return SD_OK;
}

extern "C" void sd_prefilter_params_default(sd_prefilter_params *p) {
    p->sensitivity = 5.7f;
    p->max_seqs = 300;
    p->min_ungapped_score = 15;
    p->cov_thr = 0.8f;
    p->cov_mode = 2;
    p->mask = 1;
    p->mask_prob = 0.9f;
    p->comp_bias_corr = 1;
    p->bin_size = 0;
    p->same_db = 0;
}

extern "C" int sd_target_index_build(sd_context *ctx, const char *seqs, const uint64_t *off, uint32_t n,
                                     const sd_prefilter_params *par, sd_target_index **out) {
    if (!ctx || !seqs || !off || !par || !out || n == 0) { setError("sd_target_index_build: bad argument"); return SD_EINVAL; }
    const uint64_t total = off[n];
    if (total >= 0xfff00000ull) { setError("target DB too large for 32-bit residue offsets (%llu residues)", (unsigned long long)total); return SD_ELIMIT; }
    if (n >= (1u << 22)) { setError("more than 4M targets in one index shard (%u)", n); return SD_ELIMIT; }
    for (uint32_t i = 0; i < n; i++)
        if (off[i + 1] - off[i] >= 32768) { setError("target %u longer than 32767 residues: long-sequence path not implemented", i); return SD_ELIMIT; }
    SD_CUDA(cudaSetDevice(ctx->device));
    sd::StreamScope allocScope(ctx->stream);
    const bool timing = getenv("SD_TIMING") != nullptr;
    auto tPrev = std::chrono::steady_clock::now();
    std::string tLog;
    auto tick = [&](const char *name) {
        // phase boundary: always drain the stream.  The build frees and re-allocates several 100 MB stream-ordered buffers per
        // phase; without the drain the pool sometimes grows by fresh mappings instead of reusing them (measured: 25 ms
        // -> 300-800 ms per build)
        cudaStreamSynchronize(ctx->stream);
        if (!timing) return;
        auto now = std::chrono::steady_clock::now();
        char b[96];
        snprintf(b, sizeof b, " %s=%.1f", name, std::chrono::duration<double, std::milli>(now - tPrev).count());
        tLog += b;
        tPrev = now;
    };
    std::unique_ptr<sd_target_index> idx(new sd_target_index());
    idx->ctx = ctx;
    idx->nT = n;
    idx->totalRes = total;
    idx->kmerThr = kmerThreshold(par->sensitivity, KMER);
    idx->hostOff.resize(n + 1);
    for (uint32_t i = 0; i <= n; i++) idx->hostOff[i] = (uint32_t)off[i];

    // residues: mapped on the host, uploaded once; the masked copy is made on the device (tantan.cu)
    std::vector<uint8_t> codes(total + SPAN);
    mapSequence(seqs, total, codes.data());
    std::fill(codes.begin() + total, codes.end(), (uint8_t)XCODE);
    int rc;
    cudaStream_t st = ctx->stream;
    if ((rc = idx->raw.alloc(total + SPAN)) != SD_OK) return rc;
    if ((rc = idx->res.alloc(total + SPAN)) != SD_OK) return rc;
    if ((rc = idx->off.alloc(n + 1)) != SD_OK) return rc;
    SD_CUDA(cudaMemcpyAsync(idx->raw.p, codes.data(), total + SPAN, cudaMemcpyHostToDevice, st));
    SD_CUDA(cudaMemcpyAsync(idx->res.p, idx->raw.p, total + SPAN, cudaMemcpyDeviceToDevice, st));
    SD_CUDA(cudaMemcpyAsync(idx->off.p, idx->hostOff.data(), (n + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    for (uint32_t i = 0; i < n; i++) idx->maxLen = std::max<uint32_t>(idx->maxLen, (uint32_t)(off[i + 1] - off[i]));
    tick("map+upload");
    if (par->mask && (rc = launchTantan(ctx, idx->res.p, idx->off.p, idx->hostOff, n, ctx->seedMat, (double)par->mask_prob, &idx->masked)) != SD_OK) return rc;
    tick("tantan");

    DevBuf<uint32_t> kOff;                 // dense bucket starts: build-time only
    if ((rc = kOff.alloc((size_t)NKMER + 1)) != SD_OK) return rc;

    int8_t selfHost[AA];
    for (int a = 0; a < AA; a++) selfHost[a] = (int8_t)ctx->seedMat.sub[a * AA + a];   // getScoreLookup IndexBuilder.cpp:10-22
    DevBuf<int8_t> self;
    if ((rc = self.alloc(AA)) != SD_OK) return rc;
    SD_CUDA(cudaMemcpyAsync(self.p, selfHost, AA, cudaMemcpyHostToDevice, st));

    const uint32_t nPos = (uint32_t)total;
    DevBuf<uint64_t> keysA, keysB, payloadA, payloadB;
    DevBuf<uint32_t> kmerA, kmerB;
    if ((rc = keysA.alloc(nPos)) != SD_OK || (rc = keysB.alloc(nPos)) != SD_OK) return rc;
    extractTargetKmers<<<gridFor(nPos, 256), 256, 0, st>>>(idx->res.p, idx->off.p, n, nPos, self.p, idx->kmerThr, keysA.p);
    ctx->launches++;
    SD_CUDA(cudaGetLastError());
    // sort by (seq, kmer, pos)
    {
        size_t tmpBytes = 0;
        SD_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmpBytes, keysA.p, keysB.p, (int64_t)nPos, 0, 64, st));
        DevBuf<uint8_t> tmp;
        if ((rc = tmp.alloc(tmpBytes)) != SD_OK) return rc;
        SD_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, tmpBytes, keysA.p, keysB.p, (int64_t)nPos, 0, 64, st));
        SD_CUDA(cudaStreamSynchronize(st));
    }
    tick("extract+sort64");
    keysA.release();
    if ((rc = kmerA.alloc(nPos)) != SD_OK || (rc = kmerB.alloc(nPos)) != SD_OK) return rc;
    if ((rc = payloadA.alloc(nPos)) != SD_OK || (rc = payloadB.alloc(nPos)) != SD_OK) return rc;
    SD_CUDA(cudaMemsetAsync(kOff.p, 0, ((size_t)NKMER + 1) * sizeof(uint32_t), st));
    firstOccurrence<<<gridFor(nPos, 256), 256, 0, st>>>(keysB.p, nPos, kmerA.p, payloadA.p, kOff.p);
    ctx->launches++;
    SD_CUDA(cudaGetLastError());
    // stable sort by k-mer: buckets come out ordered by (seq, pos); dropped entries (0xffffffff) go last
    {
        size_t tmpBytes = 0;
        SD_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, kmerA.p, kmerB.p, payloadA.p, payloadB.p, (int64_t)nPos, 0, 32, st));
        DevBuf<uint8_t> tmp;
        if ((rc = tmp.alloc(tmpBytes)) != SD_OK) return rc;
        SD_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmpBytes, kmerA.p, kmerB.p, payloadA.p, payloadB.p, (int64_t)nPos, 0, 32, st));
        SD_CUDA(cudaStreamSynchronize(st));
    }
    tick("first+sort32");
    // bucket starts = exclusive scan of the counts
    {
        size_t tmpBytes = 0;
        SD_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, kOff.p, kOff.p, (int64_t)NKMER + 1, st));
        DevBuf<uint8_t> tmp;
        if ((rc = tmp.alloc(tmpBytes)) != SD_OK) return rc;
        SD_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmpBytes, kOff.p, kOff.p, (int64_t)NKMER + 1, st));
    }
    uint32_t nEnt = 0;
    SD_CUDA(cudaMemcpyAsync(&nEnt, kOff.p + NKMER, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    SD_CUDA(cudaStreamSynchronize(st));
    idx->nEntries = nEnt;
    tick("scan");
    if ((rc = idx->ent.alloc(nEnt)) != SD_OK) return rc;
    if (nEnt > 0) {
        packEntries<<<gridFor(nEnt, 256), 256, 0, st>>>(payloadB.p, nEnt, idx->ent.p);
        ctx->launches++;
        SD_CUDA(cudaGetLastError());
    }
    tick("pack");
    // occupancy bitmap + rank, compact bucket starts
    {
        const uint32_t nWords = (NKMER + 31) / 32;
        DevBuf<uint32_t> cnt;
        if ((rc = idx->bmBits.alloc(2 * (size_t)nWords + 2)) || (rc = cnt.alloc(nWords + 1))) return rc;
        idx->bmRank = idx->bmBits.p + nWords;      // bits and rank share one allocation: one persisting-L2 window covers both
        SD_CUDA(cudaMemsetAsync(cnt.p + nWords, 0, sizeof(uint32_t), st));
        occupancyBits<<<gridFor((uint64_t)((nWords + 31) / 32) * 32, 256), 256, 0, st>>>(kOff.p, nWords, idx->bmBits.p, cnt.p);
        size_t tmpBytes = 0;
        SD_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, cnt.p, idx->bmRank, (int64_t)nWords + 1, st));
        DevBuf<uint8_t> tmp;
        if ((rc = tmp.alloc(tmpBytes)) != SD_OK) return rc;
        SD_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmpBytes, cnt.p, idx->bmRank, (int64_t)nWords + 1, st));
        uint32_t nDistinct = 0;
        SD_CUDA(cudaMemcpyAsync(&nDistinct, idx->bmRank + nWords, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        SD_CUDA(cudaStreamSynchronize(st));
        idx->nDistinct = nDistinct;
        if ((rc = idx->rows.alloc((size_t)nDistinct + 1)) != SD_OK) return rc;
        compactRows<<<gridFor((uint64_t)nWords * 32, 256), 256, 0, st>>>(kOff.p, idx->bmBits.p, idx->bmRank, nWords, idx->ent.p, idx->rows.p);
        ctx->launches += 2;
        SD_CUDA(cudaGetLastError());
        SD_CUDA(cudaStreamSynchronize(st));
    }
    SD_CUDA(cudaStreamSynchronize(st));
    tick("bitmap");
    if ((rc = ensureExt3(ctx)) != SD_OK) return rc;
    tick("ext3");
    if (timing) fprintf(stderr, "[sd_index]%s\n", tLog.c_str());
    *out = idx.release();
    return SD_OK;
}

extern "C" void sd_target_index_destroy(sd_target_index *idx) { delete idx; }

extern "C" uint64_t sd_target_index_entries(const sd_target_index *idx) { return idx ? idx->nEntries : 0; }

extern "C" int sd_target_index_masked(const sd_target_index *idx, uint8_t *out, uint64_t cap) {
    if (!idx || !out || cap < idx->totalRes) { setError("sd_target_index_masked: bad argument"); return SD_EINVAL; }
    SD_CUDA(cudaMemcpyAsync(out, idx->res.p, idx->totalRes, cudaMemcpyDeviceToHost, idx->ctx->stream));
    SD_CUDA(cudaStreamSynchronize(idx->ctx->stream));
    return SD_OK;
}

extern "C" int sd_query_batch_create(sd_context *ctx, const char *seqs, const uint64_t *off, uint32_t n, sd_query_batch **out) {
    if (!ctx || !seqs || !off || !out) { setError("sd_query_batch_create: bad argument"); return SD_EINVAL; }
    const uint64_t total = off[n];
    if (total >= 0xfff00000ull) { setError("query batch too large (%llu residues): split the call", (unsigned long long)total); return SD_ELIMIT; }
    for (uint32_t i = 0; i < n; i++)
        if (off[i + 1] - off[i] >= 32768) { setError("query %u longer than 32767 residues: long-sequence path not implemented", i); return SD_ELIMIT; }
    SD_CUDA(cudaSetDevice(ctx->device));
    sd::StreamScope allocScope(ctx->stream);
    std::unique_ptr<sd_query_batch> b(new sd_query_batch());
    b->ctx = ctx; b->nq = n; b->total = total;
    b->hostOff.resize(n + 1);
    for (uint32_t i = 0; i <= n; i++) b->hostOff[i] = (uint32_t)off[i];
    b->hostCode.resize(total + 16, (uint8_t)XCODE);
    mapSequence(seqs, total, b->hostCode.data());
    int rc;
    if ((rc = b->code.alloc(total + 16)) || (rc = b->off.alloc(n + 1))) return rc;
    SD_CUDA(cudaMemcpyAsync(b->code.p, b->hostCode.data(), total + 16, cudaMemcpyHostToDevice, ctx->stream));
    SD_CUDA(cudaMemcpyAsync(b->off.p, b->hostOff.data(), (n + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    SD_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = b.release();
    return SD_OK;
}

extern "C" void sd_query_batch_destroy(sd_query_batch *b) { delete b; }

extern "C" int sd_prefilter_run(sd_context *ctx, const sd_target_index *idx, const char *qseqs, const uint64_t *qoff,
                                uint32_t nq, uint32_t query_first, const sd_prefilter_params *par,
                                sd_prefilter_hit *hits, uint64_t cap, uint64_t *n_hits, sd_prefilter_stats *stats) {
    if (!ctx || !idx || !qseqs || !qoff || !par || !n_hits || (!hits && cap)) { setError("sd_prefilter_run: bad argument"); return SD_EINVAL; }
    *n_hits = 0;
    if (nq == 0) return SD_OK;
    sd_query_batch *b = nullptr;
    int rc = sd_query_batch_create(ctx, qseqs, qoff, nq, &b);
    if (rc) return rc;
    rc = sd_prefilter_run_batch(ctx, idx, b, query_first, par, hits, cap, n_hits, stats);
    sd_query_batch_destroy(b);
    return rc;
}

// keepOnDevice != nullptr: the dense hit array stays in HBM (fused clustersearch path), nothing is copied to `hits`
int sd_prefilter_core(sd_context *ctx, const sd_target_index *idx, const sd_query_batch *queries, uint32_t query_first,
                      const sd_prefilter_params *par, sd_prefilter_hit *hits, uint64_t cap, uint64_t *n_hits,
                      sd_prefilter_stats *stats, sd::DevBuf<sd_prefilter_hit> *keepOnDevice, std::vector<uint64_t> *hitOffsets) {
    if (hitOffsets) hitOffsets->assign((size_t)(queries ? queries->nq : 0) + 1, 0);
    if (!ctx || !idx || !queries || !par || !n_hits || (!hits && cap && !keepOnDevice)) { setError("sd_prefilter_run_batch: bad argument"); return SD_EINVAL; }
    *n_hits = 0;
    const uint32_t nq = queries->nq;
    if (nq == 0) return SD_OK;
    const uint64_t total = queries->total;
    SD_CUDA(cudaSetDevice(ctx->device));
    sd::StreamScope allocScope(ctx->stream);
    cudaStream_t st = ctx->stream;
    int rc;
    cudaEvent_t ev0, ev1, evk0, evk1;
    SD_CUDA(cudaEventCreate(&ev0)); SD_CUDA(cudaEventCreate(&ev1));
    SD_CUDA(cudaEventCreate(&evk0)); SD_CUDA(cudaEventCreate(&evk1));
    SD_CUDA(cudaEventRecord(ev0, st));

    const uint32_t nPos = (uint32_t)total;
    const std::vector<uint32_t> &hOff = queries->hostOff;
    std::vector<uint32_t> order(nq);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return hOff[x + 1] - hOff[x] > hOff[y + 1] - hOff[y]; });

    struct { const uint8_t *p; } dCode{queries->code.p};
    struct { const uint32_t *p; } dOff{queries->off.p};
    DevBuf<uint32_t> dOrder;
    DevBuf<float> dBias;
    DevBuf<int8_t> dCorr, dDiagSub;
    DevBuf<int16_t> dThr, dSeedSub;
    DevBuf<double> dPBack;
    if ((rc = dOrder.alloc(nq)) || (rc = dBias.alloc(total + SPAN)) || (rc = dCorr.alloc(total + SPAN)) || (rc = dThr.alloc(total + SPAN)) ||
        (rc = dSeedSub.alloc(AA * AA)) || (rc = dDiagSub.alloc(AA * AA)) || (rc = dPBack.alloc(AA)))
        return rc;
    int16_t seedSub[AA * AA];
    int8_t diagSub[AA * AA];
    for (int i = 0; i < AA * AA; i++) { seedSub[i] = ctx->seedMat.sub[i]; diagSub[i] = (int8_t)ctx->diagMat.sub[i]; }
    SD_CUDA(cudaMemcpyAsync(dOrder.p, order.data(), nq * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    SD_CUDA(cudaMemcpyAsync(dSeedSub.p, seedSub, sizeof seedSub, cudaMemcpyHostToDevice, st));
    SD_CUDA(cudaMemcpyAsync(dDiagSub.p, diagSub, sizeof diagSub, cudaMemcpyHostToDevice, st));
    SD_CUDA(cudaMemcpyAsync(dPBack.p, ctx->seedMat.pBack, AA * sizeof(double), cudaMemcpyHostToDevice, st));
    SD_CUDA(cudaMemsetAsync(dBias.p + total, 0, SPAN * sizeof(float), st));
    compositionBias<<<gridFor(nPos, 128), 128, 0, st>>>(dCode.p, dOff.p, nq, nPos, dSeedSub.p, dPBack.p, par->comp_bias_corr, dBias.p);
    profileCorrection<<<gridFor(nPos, 256), 256, 0, st>>>(dBias.p, nPos, dCorr.p);
    kmerThresholds<<<gridFor(nPos, 256), 256, 0, st>>>(dCode.p, dOff.p, nq, nPos, dBias.p, idx->kmerThr, dThr.p);
    ctx->launches += 3;
    SD_CUDA(cudaGetLastError());

    const uint32_t maxHits = std::min<uint32_t>((uint32_t)std::max(par->max_seqs, 1), idx->nT);
    const int binSize = par->bin_size > 0 ? par->bin_size : referenceBinSize(idx->nT);
    static const int ctasPerSm = getenv("SD_PF_CTAS") ? atoi(getenv("SD_PF_CTAS")) : 4;
    const uint32_t nCta = std::max(1u, std::min<uint32_t>((uint32_t)(ctx->smCount * ctasPerSm), (nq + PF_WARPS - 1) / PF_WARPS));
    const uint32_t nWarps = nCta * PF_WARPS;
    // first-pass capacity of the per-warp candidate lists (SD_PF_CANDCAP: test hook that forces the second pass)
    // grows with the number of targets (long queries against 800 k targets have > 32 k candidates; re-running them in the
    // few-warp second pass costs seconds): 32 k up to 256 k targets, 128 k at 800 k targets
    uint32_t candCap = 32768u;
    while (candCap < 262144u && (uint64_t)candCap * 8u < idx->nT) candCap *= 2;
    if (getenv("SD_PF_CANDCAP")) candCap = (uint32_t)std::max(32, atoi(getenv("SD_PF_CANDCAP")));

    ScratchBuf<uint32_t> dTable(ctx, SL_PF_TABLE), dResT(ctx, SL_PF_REST);
    DevBuf<uint32_t> dOutCount, dNext, dFail;
    ScratchBuf<uint64_t> dCand(ctx, SL_PF_CAND), dKeep(ctx, SL_PF_KEEP);
    ScratchBuf<int32_t> dResScore(ctx, SL_PF_RESS);
    ScratchBuf<uint16_t> dResDiag(ctx, SL_PF_RESD);
    ScratchBuf<sd_prefilter_hit> dOut(ctx, SL_PF_OUT);
    DevBuf<unsigned long long> dStats;
    if ((rc = dTable.alloc((size_t)nWarps * idx->nT)) || (rc = dCand.alloc((size_t)nWarps * candCap)) || (rc = dKeep.alloc((size_t)nWarps * candCap)) ||
        (rc = dResT.alloc((size_t)nWarps * maxHits)) || (rc = dResScore.alloc((size_t)nWarps * maxHits)) || (rc = dResDiag.alloc((size_t)nWarps * maxHits)) ||
        (rc = dOut.alloc((size_t)nq * maxHits)) || (rc = dOutCount.alloc(nq)) || (rc = dNext.alloc(4)) || (rc = dStats.alloc(ST_N)) ||
        (rc = dFail.alloc(nq)))
        return rc;
    SD_CUDA(cudaMemsetAsync(dTable.p, 0, (size_t)nWarps * idx->nT * sizeof(uint32_t), st));
    SD_CUDA(cudaMemsetAsync(dOutCount.p, 0, nq * sizeof(uint32_t), st));
    SD_CUDA(cudaMemsetAsync(dStats.p, 0, ST_N * sizeof(unsigned long long), st));
    const uint32_t initNext[4] = {0u, 0xffffffffu, 0u, 0u};     // next query, first failing query, number of overflowed queries
    SD_CUDA(cudaMemcpyAsync(dNext.p, initNext, sizeof initNext, cudaMemcpyHostToDevice, st));

    PrefilterArgs a{};
    a.tRes = idx->res.p; a.tOff = idx->off.p; a.nT = idx->nT;
    a.bmBits = idx->bmBits.p; a.bmRank = idx->bmRank; a.rows = idx->rows.p; a.ent = idx->ent.p;
    a.e3score = ctx->e3score.p; a.e3index = ctx->e3index.p; a.e3cnt = ctx->e3cnt.p; a.diagSub = dDiagSub.p;
    a.q.code = dCode.p; a.q.off = dOff.p; a.q.corr = dCorr.p; a.q.thr = dThr.p;
    a.nQ = nq; a.order = dOrder.p; a.queryFirst = query_first; a.sameDb = par->same_db;
    a.maxHits = maxHits; a.minDiagScore = (uint32_t)std::max(par->min_ungapped_score, 0);
    a.covThr = par->cov_thr; a.covMode = par->cov_mode; a.binMask = (uint32_t)binSize - 1u;
    a.table = dTable.p; a.candCap = candCap; a.cand = dCand.p; a.keep = dKeep.p;
    a.resT = dResT.p; a.resScore = dResScore.p; a.resDiag = dResDiag.p;
    a.out = dOut.p; a.outCount = dOutCount.p; a.stats = dStats.p; a.nextQuery = dNext.p; a.errQuery = dNext.p + 1;
    a.failList = dFail.p; a.nFail = dNext.p + 2;

    SD_CUDA(cudaEventRecord(evk0, st));
    static const int envEvict = getenv("SD_PF_EVICT") ? atoi(getenv("SD_PF_EVICT")) : 0;
    static const int envPersist = getenv("SD_PF_PERSIST") ? atoi(getenv("SD_PF_PERSIST")) : 0;
    static const int envGran = getenv("SD_PF_L2GRAN") ? atoi(getenv("SD_PF_L2GRAN")) : 0;
    a.evictFirst = envEvict;
    static const int envPrefetch = getenv("SD_PF_PREFETCH") ? atoi(getenv("SD_PF_PREFETCH")) : 0;
    a.prefetch = envPrefetch;
    a.debugSkip = getenv("SD_PF_SKIP") ? atoi(getenv("SD_PF_SKIP")) : 0;
    if (envGran) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)envGran);
    if (envPersist) {
        // the occupancy bitmap is probed once per similar k-mer: keep it in the persisting part of the L2
        int maxPersist = 0, maxWin = 0;
        cudaDeviceGetAttribute(&maxPersist, cudaDevAttrMaxPersistingL2CacheSize, ctx->device);
        cudaDeviceGetAttribute(&maxWin, cudaDevAttrMaxAccessPolicyWindowSize, ctx->device);
        const size_t bmBytes = idx->bmBits.n * sizeof(uint32_t);
        if (maxPersist > 0 && maxWin > 0) {
            const cudaError_t e1 = cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min<size_t>((size_t)maxPersist, bmBytes + (8u << 20)));
            cudaStreamAttrValue av{};
            av.accessPolicyWindow.base_ptr = (void *)idx->bmBits.p;
            av.accessPolicyWindow.num_bytes = std::min<size_t>(bmBytes, (size_t)maxWin);
            av.accessPolicyWindow.hitRatio = 1.0f;
            av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            av.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
            const cudaError_t e2 = cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av);
            if (getenv("SD_PF_DEBUG")) fprintf(stderr, "[sd] persisting L2: max %d B, window max %d B, bitmap %zu B, setLimit=%d setAttr=%d\n", maxPersist, maxWin, bmBytes, (int)e1, (int)e2);
        } else if (getenv("SD_PF_DEBUG")) fprintf(stderr, "[sd] persisting L2 not available (max %d, window %d)\n", maxPersist, maxWin);
    }
    SD_CUDA(cudaFuncSetAttribute(prefilterKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PF_DYN_SMEM));
    prefilterKernel<<<nCta, PF_WARPS * 32, PF_DYN_SMEM, st>>>(a);
    ctx->launches++;
    SD_CUDA(cudaGetLastError());
    {   // queries with more than candCap double-hit candidates (long queries against very large target sets): second launch
        // over just those, few warps with a 4 M entry list each
        uint32_t hN[4];
        SD_CUDA(cudaMemcpyAsync(hN, dNext.p, sizeof hN, cudaMemcpyDeviceToHost, st));
        SD_CUDA(cudaStreamSynchronize(st));
        const uint32_t nFail = hN[2];
        if (nFail) {
            const uint32_t bigCap = 1u << 22;
            const uint32_t nCta2 = std::max(1u, std::min<uint32_t>((nFail + PF_WARPS - 1) / PF_WARPS, 16u));
            const uint32_t nWarps2 = nCta2 * PF_WARPS;
            DevBuf<uint64_t> dCand2, dKeep2;
            DevBuf<uint32_t> dOrder2;
            if ((rc = dCand2.alloc((size_t)nWarps2 * bigCap)) || (rc = dKeep2.alloc((size_t)nWarps2 * bigCap)) || (rc = dOrder2.alloc(nFail))) return rc;
            SD_CUDA(cudaMemcpyAsync(dOrder2.p, dFail.p, nFail * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
            SD_CUDA(cudaMemsetAsync(dTable.p, 0, (size_t)nWarps2 * idx->nT * sizeof(uint32_t), st));      // epochs restart at 0
            SD_CUDA(cudaMemsetAsync(dStats.p + ST_ERR_CAND, 0, sizeof(unsigned long long), st));
            const uint32_t init2[4] = {0u, 0xffffffffu, 0u, 0u};
            SD_CUDA(cudaMemcpyAsync(dNext.p, init2, sizeof init2, cudaMemcpyHostToDevice, st));
            PrefilterArgs b = a;
            b.order = dOrder2.p; b.nQ = nFail; b.candCap = bigCap; b.cand = dCand2.p; b.keep = dKeep2.p;
            prefilterKernel<<<nCta2, PF_WARPS * 32, PF_DYN_SMEM, st>>>(b);
            ctx->launches++;
            SD_CUDA(cudaGetLastError());
            SD_CUDA(cudaStreamSynchronize(st));     // init2 and the second-pass buffers go out of scope
        }
    }
    SD_CUDA(cudaEventRecord(evk1, st));
    if (envPersist) {
        cudaStreamAttrValue av{};
        av.accessPolicyWindow.num_bytes = 0;
        cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av);
    }

    std::vector<uint32_t> hCount(nq);
    unsigned long long hStats[ST_N];
    uint32_t hNext[4];
    SD_CUDA(cudaMemcpyAsync(hCount.data(), dOutCount.p, nq * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    SD_CUDA(cudaMemcpyAsync(hStats, dStats.p, sizeof hStats, cudaMemcpyDeviceToHost, st));
    SD_CUDA(cudaMemcpyAsync(hNext, dNext.p, sizeof hNext, cudaMemcpyDeviceToHost, st));
    SD_CUDA(cudaStreamSynchronize(st));
    if (hStats[ST_ERR_CAND]) { setError("prefilter: %llu queries produced more than %u diagonal candidates (first: query %u)", hStats[ST_ERR_CAND], 1u << 22, hNext[1]); return SD_ELIMIT; }
    uint64_t totalHits = 0;
    for (uint32_t i = 0; i < nq; i++) totalHits += hCount[i];
    if (hitOffsets)
        for (uint32_t i = 0; i < nq; i++) (*hitOffsets)[i + 1] = (*hitOffsets)[i] + hCount[i];
    *n_hits = totalHits;
    if (!keepOnDevice && totalHits > cap) { setError("sd_prefilter_run: %llu hits, buffer holds %llu", (unsigned long long)totalHits, (unsigned long long)cap); return SD_ESMALL; }
    // gather the per-query slices (strided device layout -> dense) and copy once
    if (totalHits) {
        DevBuf<uint32_t> dOffs;
        DevBuf<sd_prefilter_hit> dDenseLocal;
        DevBuf<sd_prefilter_hit> &dDense = keepOnDevice ? *keepOnDevice : dDenseLocal;
        DevBuf<uint8_t> tmp;
        if ((rc = dOffs.alloc(nq)) || (rc = dDense.alloc(totalHits))) return rc;
        size_t tmpBytes = 0;
        SD_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, dOutCount.p, dOffs.p, (int64_t)nq, st));
        if ((rc = tmp.alloc(tmpBytes))) return rc;
        SD_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmpBytes, dOutCount.p, dOffs.p, (int64_t)nq, st));
        compactHits<<<gridFor((uint64_t)nq * 32, 256), 256, 0, st>>>(dOut.p, dOutCount.p, dOffs.p, nq, maxHits, dDense.p);
        ctx->launches++;
        SD_CUDA(cudaGetLastError());
        if (!keepOnDevice) SD_CUDA(cudaMemcpyAsync(hits, dDense.p, totalHits * sizeof(sd_prefilter_hit), cudaMemcpyDeviceToHost, st));
        SD_CUDA(cudaStreamSynchronize(st));
    }
    SD_CUDA(cudaEventRecord(ev1, st));
    SD_CUDA(cudaStreamSynchronize(st));
    if (stats) {
        float msK = 0, msT = 0;
        cudaEventElapsedTime(&msK, evk0, evk1);
        cudaEventElapsedTime(&msT, ev0, ev1);
        stats->probes = hStats[ST_PROBES]; stats->matches = hStats[ST_MATCHES]; stats->diag_residues = hStats[ST_DIAGRES];
        stats->query_residues = total; stats->hits = hStats[ST_HITS]; stats->candidates = hStats[ST_CANDS];
        stats->truncated = hStats[ST_TRUNC]; stats->kernel_ms = msK; stats->total_ms = msT;
    }
    cudaEventDestroy(ev0); cudaEventDestroy(ev1); cudaEventDestroy(evk0); cudaEventDestroy(evk1);
    return SD_OK;
}

extern "C" int sd_prefilter_run_batch(sd_context *ctx, const sd_target_index *idx, const sd_query_batch *queries, uint32_t query_first,
                                      const sd_prefilter_params *par, sd_prefilter_hit *hits, uint64_t cap, uint64_t *n_hits,
                                      sd_prefilter_stats *stats) {
    return sd_prefilter_core(ctx, idx, queries, query_first, par, hits, cap, n_hits, stats, nullptr);
}
