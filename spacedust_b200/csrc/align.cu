// `align` stage on B200: host orchestration of the Smith-Waterman passes (sw.cu), the banded
// traceback kernel, E-values, acceptance criteria and ordering.
//
// Reference behaviour restated (paths relative to /root/reference/lib/mmseqs/src):
//   alignment/Alignment.cpp:312-511            per-query loop, canBeCovered pre-filter, checkCriteria, sort
//   alignment/Matcher.cpp:60-142               getSWResult (coverage, seqId, bit score)
//   alignment/StripedSmithWaterman.cpp:310-545 ssw_align_private (byte pass, word pass on overflow, early outs,
//                                              reverse pass for the start, banded traceback)
//   alignment/StripedSmithWaterman.cpp:1348-1600 banded_sw (band doubling, direction bytes, tie rules)
//   alignment/StripedSmithWaterman.cpp:1675-1710 scoreIdentical
#include <atomic>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstring>
#include <memory>
#include <thread>

#include <cub/cub.cuh>

#include "sd_internal.h"
#include "prep.cuh"
#include "sw.cuh"

namespace sd {

// ------------------------------------------------------------------------------------------------
// banded traceback (banded_sw + computerBacktrace, StripedSmithWaterman.cpp:1348-1600,548-581)
//
// A warp fills the direction matrices of up to 32 alignments one after the other into its scratch slot -- lanes across the
// band of one query row, the horizontal gap chain F_j = max(H_{j-1} - go, F_{j-1} - ge) as a decayed max-scan, row buffers
// in shared memory while the band fits, ONE direction byte per cell (bit 0: E came from H, bit 1: F came from H, bits 2-4:
// the H source) -- and then walks the 32 paths in parallel, one lane per alignment, so the dependent loads of the walk
// overlap.  The walk run-length encodes on the fly and writes the CIGAR text right-to-left at the end of the alignment's
// pool slot (the path is found end to start), so only the compressed string ever leaves the device.
// ------------------------------------------------------------------------------------------------
struct TbTask {
    uint32_t q, t;
    int32_t qStart, qEnd, tStart, tEnd, score;
    uint32_t pad;
    uint64_t btOff;       // pool slot of 2 * (rows + columns) + 2 bytes (worst case: every run has length 1)
};

struct TbResult {
    int32_t btLen;        // alignment columns (M/I/D operations)
    int32_t ids;
    int32_t status;       // 0 ok, 1 scratch too small, 2 trace walked off the band
    int32_t band;
    uint32_t cigLen;      // CIGAR text bytes
    uint32_t cigStart;    // offset of the text inside the slot
};

struct TbArgs {
    const uint8_t *qCode;
    const int8_t *qCb;
    const uint32_t *qOff;
    const uint8_t *tRes;
    const uint32_t *tOff;
    const int8_t *mat;       // 21x21
    int gapOpen, gapExtend;
    const TbTask *tasks;
    const uint32_t *order;   // optional subset
    TbResult *results;
    uint32_t *cigLen;        // per task (scan input of the pool compaction)
    uint32_t nTasks;
    uint32_t *next;
    // per-warp scratch
    int8_t *dir;
    uint64_t dirCap;
    int *rows;               // 3 x rowCap ints (bands too wide for shared memory)
    uint32_t rowCap;
    char *btPool;
    unsigned long long *cells;
};

constexpr int TB_WARPS = 4;
constexpr int TB_BATCH = 32;
constexpr int TB_SROW = 264;     // shared-memory row buffers: bands up to 129
constexpr int TB_NEG = -(1 << 28);

__device__ __forceinline__ uint32_t tbSlotBytes(const TbTask &tk) { return 2u * (uint32_t)((tk.qEnd - tk.qStart + 1) + (tk.tEnd - tk.tStart + 1)) + 2u; }

__global__ void __launch_bounds__(TB_WARPS * 32, 8) tracebackKernel(TbArgs A) {
    __shared__ int8_t sMat[AA * AA];
    __shared__ int sRows[TB_WARPS][3 * TB_SROW];
    __shared__ uint32_t sId[TB_WARPS][TB_BATCH];
    __shared__ uint32_t sBase[TB_WARPS][TB_BATCH];
    __shared__ int sBand[TB_WARPS][TB_BATCH];
    for (int i = threadIdx.x; i < AA * AA; i += blockDim.x) sMat[i] = A.mat[i];
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t wIn = threadIdx.x >> 5;
    const uint32_t warp = blockIdx.x * TB_WARPS + wIn;
    int8_t *dir = A.dir + (size_t)warp * A.dirCap;
    int *gRows = A.rows + (size_t)warp * 3 * A.rowCap;
    const int go = A.gapOpen, ge = A.gapExtend;
    unsigned long long cells = 0;
    uint32_t carry = 0xffffffffu;
    bool drained = false;

    for (;;) {
        // ---------------- fill: direction bytes of up to TB_BATCH alignments behind each other in the scratch slot
        int nb = 0;
        uint64_t used = 0;
        while (nb < TB_BATCH) {
            uint32_t ti = carry;
            carry = 0xffffffffu;
            if (ti == 0xffffffffu) {
                if (drained) break;
                if (lane == 0) ti = atomicAdd(A.next, 1u);
                ti = __shfl_sync(0xffffffffu, ti, 0);
                if (ti >= A.nTasks) { drained = true; break; }
            }
            const uint32_t id = A.order ? A.order[ti] : ti;
            const TbTask tk = A.tasks[id];
            const int qLen = tk.qEnd - tk.qStart + 1, dbLen = tk.tEnd - tk.tStart + 1;
            const uint8_t *qc = A.qCode + A.qOff[tk.q] + tk.qStart;
            const int8_t *qb = A.qCb + A.qOff[tk.q] + tk.qStart;
            const uint8_t *tc = A.tRes + A.tOff[tk.t] + tk.tStart;
            int band = abs(dbLen - qLen) + 1;
            int maxH = 0;
            int status = 0;
            long widthD = 0;
            uint64_t need = 0;
            for (;;) {
                const long width = (long)band * 2 + 3;
                widthD = (long)band * 2 + 1;
                need = (uint64_t)widthD * (uint64_t)qLen;
                if (need > A.dirCap) { status = 1; break; }                  // needs the large-slot pass
                if (need > A.dirCap - used) { status = 3; break; }           // no room behind this batch: opens the next one
                int *hb, *eb, *hc;
                if (width + 2 <= TB_SROW) { hb = sRows[wIn]; eb = hb + TB_SROW; hc = eb + TB_SROW; }
                else if ((unsigned long long)width + 2 <= A.rowCap) { hb = gRows; eb = hb + A.rowCap; hc = eb + A.rowCap; }
                else { status = 1; break; }
                int8_t *dBase = dir + used;
                for (long u = lane; u < width + 1; u += 32) { hb[u] = 0; eb[u] = 0; hc[u] = 0; }
                __syncwarp();
                for (int i = 0; i < qLen; i++) {
                    const int beg = max(0, i - band), end = min(dbLen - 1, i + band);
                    const int edge = (int)min((long)end + 1, width - 1);
                    const int xi = max(i - band, 0), xim = max(i - 1 - band, 0);
                    if (lane == 0) { hb[0] = 0; eb[0] = 0; hb[edge] = 0; eb[edge] = 0; hc[0] = 0; }
                    __syncwarp();
                    const int8_t *mrow = sMat + (int)qc[i] * AA;
                    const int cbv = qb[i];
                    int8_t *line = dBase + widthD * (long)i;
                    int carryH = 0, carryF = 0;
                    for (int j0 = beg; j0 <= end; j0 += 32) {
                        const int j = j0 + (int)lane;
                        const bool valid = j <= end;
                        const int u = j - xi + 1, up = j - xim + 1, dg = j - xim;
                        int hbu = 0, ebu = 0, hbd = 0, tj = 0;
                        if (valid) { hbu = hb[up]; ebu = eb[up]; hbd = hb[dg]; tj = tc[j]; }
                        __syncwarp();
                        int t1 = (i == 0) ? -go : hbu - go;
                        int t2 = (i == 0) ? -ge : ebu - ge;
                        const int e = t1 > t2 ? t1 : t2;
                        const int de = t1 > t2 ? 3 : 2;
                        const int e1 = e > 0 ? e : 0;
                        const int diagv = hbd + (int)mrow[tj] + cbv;
                        const int h0 = e1 > diagv ? e1 : diagv;
                        // horizontal gap chain F_j = max(H_{j-1} - go, F_{j-1} - ge) as a decayed max-scan
                        int y = valid ? h0 - go : TB_NEG;
#pragma unroll
                        for (int d = 1; d < 32; d <<= 1) {
                            const int o = __shfl_up_sync(0xffffffffu, y, d);
                            if (lane >= (uint32_t)d) y = max(y, o - d * ge);
                        }
                        const int yPrev = __shfl_up_sync(0xffffffffu, y, 1);
                        const int f0 = max(carryH - go, carryF - ge);
                        const int f = (lane == 0) ? f0 : max(yPrev, f0 - (int)lane * ge);
                        const int f1 = f > 0 ? f : 0;
                        const int g1 = e1 > f1 ? e1 : f1;
                        const int h = g1 > diagv ? g1 : diagv;
                        int hPrev = __shfl_up_sync(0xffffffffu, h, 1);
                        int fPrev = __shfl_up_sync(0xffffffffu, f, 1);
                        if (lane == 0) { hPrev = carryH; fPrev = carryF; }
                        const int df = (hPrev - go > fPrev - ge) ? 5 : 4;
                        const int dh = (g1 <= diagv) ? 1 : (e1 > f1 ? de : df);
                        if (valid) {
                            eb[u] = e;
                            hc[u] = h;
                            line[j - xi] = (int8_t)((de == 3 ? 1 : 0) | (df == 5 ? 2 : 0) | (dh << 2));
                            maxH = max(maxH, h);
                        }
                        const int lastLane = min(31, end - j0);
                        carryH = __shfl_sync(0xffffffffu, h, lastLane);
                        carryF = __shfl_sync(0xffffffffu, f, lastLane);
                        __syncwarp();
                    }
                    const int uMax = end - xi + 1;
                    for (int u = 1 + (int)lane; u <= uMax; u += 32) hb[u] = hc[u];
                    __syncwarp();
                    cells += (lane == 0) ? (unsigned long long)(end - beg + 1) : 0ull;
                }
                int m = maxH;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
                maxH = m;
                if (maxH >= tk.score) break;
                band *= 2;
                if (band > 4 * (qLen + dbLen) + 8) { status = 2; break; }
            }
            if (status == 3) { carry = ti; break; }
            if (status != 0) {
                if (lane == 0) {
                    TbResult r;
                    r.btLen = 0; r.ids = 0; r.status = status; r.band = band; r.cigLen = 0; r.cigStart = 0;
                    A.results[id] = r;
                    A.cigLen[id] = 0;
                }
                continue;
            }
            if (lane == 0) { sId[wIn][nb] = id; sBase[wIn][nb] = (uint32_t)used; sBand[wIn][nb] = band; }
            used += (need + 15) & ~(uint64_t)15;
            nb++;
        }
        __syncwarp();
        __threadfence_block();
        if (nb == 0) {
            if (drained && carry == 0xffffffffu) break;
            continue;
        }
        // ---------------- walk: lane l follows the path of the batch's l-th alignment, CIGAR written right to left
        if ((int)lane < nb) {
            const uint32_t id = sId[wIn][lane];
            const TbTask tk = A.tasks[id];
            const int band = sBand[wIn][lane];
            const long widthD = (long)band * 2 + 1;
            const int8_t *d = dir + sBase[wIn][lane];
            const int qLen = tk.qEnd - tk.qStart + 1, dbLen = tk.tEnd - tk.tStart + 1;
            const uint8_t *qc = A.qCode + A.qOff[tk.q] + tk.qStart;
            const uint8_t *tc = A.tRes + A.tOff[tk.t] + tk.tStart;
            const uint32_t slot = tbSlotBytes(tk);
            char *slotEnd = A.btPool + tk.btOff + slot;
            int i = qLen - 1, j = dbLen - 1, state = 2, n = 0, ids = 0, status = 0;
            uint32_t pos = 0, run = 0;
            char cur = 0;
            auto flush = [&]() {
                if (!run) return;
                slotEnd[-1 - (long)pos++] = cur;
                uint32_t r = run;
                do { slotEnd[-1 - (long)pos++] = (char)('0' + r % 10u); r /= 10u; } while (r);
            };
            auto emit = [&](char op) {
                if (op == cur) run++;
                else { flush(); cur = op; run = 1; }
                n++;
            };
            while (i > 0 || j > 0) {
                if (i < 0 || j < 0) { status = 2; break; }
                const int xi = max(i - band, 0);
                const long x = j - xi;
                if (x < 0 || x >= widthD) { status = 2; break; }
                const int c = (int)(uint8_t)d[widthD * (long)i + x];
                const int dd = state == 0 ? ((c & 1) ? 3 : 2) : (state == 1 ? ((c & 2) ? 5 : 4) : (c >> 2));
                if (dd == 1) { ids += (qc[i] == tc[j]); --i; --j; state = 2; emit('M'); }
                else if (dd == 2) { --i; state = 0; emit('I'); }
                else if (dd == 3) { --i; state = 2; emit('I'); }
                else if (dd == 4) { --j; state = 1; emit('D'); }
                else if (dd == 5) { --j; state = 2; emit('D'); }
                else { status = 2; break; }
                // (text <= 2 bytes per operation and every operation lowers i or j: the slot of 2*(rows+columns)+2 cannot overflow)
            }
            if (status == 0) { ids += (qc[0] == tc[0]); emit('M'); flush(); }
            TbResult r;
            r.btLen = n; r.ids = ids; r.status = status; r.band = band; r.cigLen = status == 0 ? pos : 0u; r.cigStart = slot - pos;
            A.results[id] = r;
            A.cigLen[id] = r.cigLen;
        }
        __syncwarp();
    }
    if (lane == 0 && cells) atomicAdd(A.cells, cells);
}

// dense CIGAR pool: one warp copies one alignment's text from its slot to its scanned offset
__global__ void gatherCigars(const TbTask *tasks, const TbResult *res, const unsigned long long *denseOff, uint32_t n, const char *pool, char *dense) {
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n) return;
    const TbResult r = res[w];
    const char *src = pool + tasks[w].btOff + r.cigStart;
    char *dst = dense + denseOff[w];
    for (uint32_t k = lane; k < r.cigLen; k += 32) dst[k] = src[k];
}

struct U32ToU64 {
    __host__ __device__ unsigned long long operator()(uint32_t v) const { return (unsigned long long)v; }
};

// ------------------------------------------------------------------------------------------------
// host helpers
// ------------------------------------------------------------------------------------------------
static inline float covOf(unsigned start, unsigned end, unsigned len) {   // SmithWaterman::computeCov
    return (std::min(len, std::max(start, end)) - std::min(start, end) + 1) / (float)len;
}

static inline bool hasCoverage(float covThr, int covMode, float q, float t) {   // Util::hasCoverage
    switch (covMode) {
        case 0: return (q >= covThr) && (t >= covThr);
        case 2: return q >= covThr;
        case 1: return t >= covThr;
        default: return true;
    }
}

static inline bool canBeCovered(float covThr, int covMode, float q, float t) {   // Util::canBeCovered
    switch (covMode) {
        case 0: return (q / t >= covThr) && (t / q >= covThr);
        case 1: return (q / t) >= covThr;
        case 2: return (t / q) >= covThr;
        case 3: return ((t / q) >= covThr) && (t / q) <= 1.0;
        case 4: return ((q / t) >= covThr) && (q / t) <= 1.0;
        case 5: return (std::min(t, q) / std::max(t, q)) >= covThr;
        default: return true;
    }
}

}  // namespace sd

using namespace sd;

extern "C" void sd_align_params_default(sd_align_params *p) {
    p->evalue_thr = 10.0;
    p->min_seq_id = 0.0;
    p->cov_thr = 0.8f;
    p->cov_mode = 2;
    p->min_aln_len = 30;
    p->comp_bias_corr = 1;
    p->gap_open = 11;
    p->gap_extend = 1;
    p->score_bias = 0.0f;
    p->same_db = 0;
}


extern "C" int sd_align_run(sd_context *ctx, const sd_target_index *idx, const char *qseqs, const uint64_t *qoff, uint32_t nq,
                            uint32_t query_first, const uint32_t *pair_q, const uint32_t *pair_t, uint64_t n_pairs,
                            uint64_t db_residues, const sd_align_params *par, sd_alignment *out, uint64_t cap, uint64_t *n_out,
                            char *bt_pool, uint64_t bt_cap, uint64_t *bt_used, sd_align_stats *stats) {
    if (!ctx || !idx || !qseqs || !qoff || !par || !n_out || !bt_used) { setError("sd_align_run: bad argument"); return SD_EINVAL; }
    *n_out = 0; *bt_used = 0;
    if (stats) memset(stats, 0, sizeof *stats);
    if (nq == 0 || n_pairs == 0) return SD_OK;
    sd_query_batch *b = nullptr;
    int rc = sd_query_batch_create(ctx, qseqs, qoff, nq, &b);
    if (rc) return rc;
    rc = sd_align_run_batch(ctx, idx, b, query_first, pair_q, pair_t, n_pairs, db_residues, par, out, cap, n_out, bt_pool, bt_cap, bt_used, stats);
    sd_query_batch_destroy(b);
    return rc;
}

// ================================================================================================
// device-side filter chain of align: forward 8-bit -> (overflow) forward 16-bit -> E-value score threshold +
// end-point coverage -> reverse pass -> start coverage -> traceback.  Only survivors come back to the host.
// ================================================================================================
namespace {

struct PairHit {            // one record per pair that reaches the reverse pass
    uint32_t q, t;
    int32_t score, qEnd, tEnd, qStart, tStart;
    int32_t word;           // 16-bit semantics were needed
    int32_t stage;          // 1 waits for reverse, 2 waits for traceback, 0 dropped
    uint32_t tb;            // index of its traceback task
};

struct ChainArgs {
    const uint32_t *pairQ, *pairT;
    uint64_t nPairs;
    const uint32_t *qLen;         // per query
    const uint32_t *tOff;         // per target
    const int32_t *limit8;        // per query: 255 - bias
    const int32_t *minScore;      // per query: smallest score with E-value <= threshold
    uint32_t queryFirst;
    int sameDb;
    float covThr;
    int covMode;
    // lists
    SwTask *fwdTasks; uint32_t *nFwd;
    SwTask *wordTasks; uint32_t *nWord;
    SwTask *bigTasks; uint32_t *nBig;
    SwTask *revTasks; uint32_t *nRev;
    PairHit *hits; uint32_t *nHits;
    uint32_t *identity; uint32_t *nIdentity;     // queries whose identity pair takes the scoreIdentical shortcut
    TbTask *tbTasks; uint32_t *nTb;
    unsigned long long *poolCursor;
    uint32_t *classCount;                        // 18 counters per list, used by orderByClass
    unsigned long long *counters;                // [0] aligned pairs, [1] score mismatches
};

__device__ __forceinline__ float devCov(unsigned start, unsigned end, unsigned len) {
    return __fdiv_rn((float)(min(len, max(start, end)) - min(start, end) + 1), (float)len);
}
__device__ __forceinline__ bool devHasCoverage(float thr, int mode, float q, float t) {
    switch (mode) {
        case 0: return (q >= thr) && (t >= thr);
        case 2: return q >= thr;
        case 1: return t >= thr;
        default: return true;
    }
}
__device__ __forceinline__ bool devCanBeCovered(float thr, int mode, float q, float t) {
    switch (mode) {
        case 0: return (__fdiv_rn(q, t) >= thr) && (__fdiv_rn(t, q) >= thr);
        case 1: return __fdiv_rn(q, t) >= thr;
        case 2: return __fdiv_rn(t, q) >= thr;
        case 3: return (__fdiv_rn(t, q) >= thr) && ((double)__fdiv_rn(t, q) <= 1.0);
        case 4: return (__fdiv_rn(q, t) >= thr) && ((double)__fdiv_rn(q, t) <= 1.0);
        case 5: return __fdiv_rn(fminf(t, q), fmaxf(t, q)) >= thr;
        default: return true;
    }
}
__device__ __forceinline__ uint32_t swClass(int nRows) { const int rp = (nRows + 63) / 64; return (uint32_t)(rp > 16 ? 17 : rp); }

// Alignment.cpp:370 length pre-filter, identity shortcut, forward task per remaining pair
__global__ void chainBuildForward(ChainArgs A) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= A.nPairs) return;
    const uint32_t q = A.pairQ[k], t = A.pairT[k];
    const int Lq = (int)A.qLen[q], Lt = (int)(A.tOff[t + 1] - A.tOff[t]);
    if (!devCanBeCovered(A.covThr, A.covMode, (float)Lq, (float)Lt)) return;
    atomicAdd(&A.counters[0], 1ull);
    if (A.sameDb && A.queryFirst + q == t) { A.identity[atomicAdd(A.nIdentity, 1u)] = q; return; }
    SwTask tk;
    tk.q = q; tk.t = t; tk.nRows = Lq; tk.nCols = Lt; tk.limit = A.limit8[q]; tk.width = 32; tk.reverse = 0; tk.pad = 0;
    A.fwdTasks[atomicAdd(A.nFwd, 1u)] = tk;
}

// largest class first; order inside a class is arbitrary (results are keyed by task)
__global__ void chainOrderByClass(const SwTask *tasks, uint32_t n, uint32_t *classCursor /*18, pre-scanned*/, uint32_t *order) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    order[atomicAdd(&classCursor[17 - swClass(tasks[i].nRows)], 1u)] = i;
}

__global__ void chainCountClasses(const SwTask *tasks, uint32_t n, uint32_t *classCount) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    atomicAdd(&classCount[17 - swClass(tasks[i].nRows)], 1u);
}

__global__ void chainScanClasses(uint32_t *classCount /*18*/) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        uint32_t acc = 0;
        for (int c = 0; c < 18; c++) { const uint32_t v = classCount[c]; classCount[c] = acc; acc += v; }
    }
}

// ssw_align_private :383-398 after a forward pass: overflow -> next width; E-value / end-point coverage early outs
__global__ void chainAfterForward(ChainArgs A, const SwTask *tasks, const SwResult *res, uint32_t n, int wordPass) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const SwTask tk = tasks[i];
    const SwResult r = res[i];
    if (r.overflow) {
        SwTask nt = tk;
        nt.limit = INT_MAX; nt.width = 16;
        if (!wordPass && r.overflow == 1) A.wordTasks[atomicAdd(A.nWord, 1u)] = nt;
        else A.bigTasks[atomicAdd(A.nBig, 1u)] = nt;          // 16-bit lanes exhausted: 32-bit pass
        return;
    }
    if (r.col < 0) return;                                     // no residue could be aligned
    const int Lq = tk.nRows, Lt = tk.nCols;
    if (r.score < A.minScore[tk.q]) return;                    // evalue > evalThr
    const float qcov = devCov(0u, (unsigned)r.row, (unsigned)Lq), tcov = devCov(0u, (unsigned)r.col, (unsigned)Lt);
    if (!devHasCoverage(A.covThr, A.covMode, qcov, tcov)) return;
    const uint32_t h = atomicAdd(A.nHits, 1u);
    PairHit ph;
    ph.q = tk.q; ph.t = tk.t; ph.score = r.score; ph.qEnd = r.row; ph.tEnd = r.col; ph.qStart = -1; ph.tStart = -1;
    ph.word = tk.width == 16; ph.stage = 1; ph.tb = 0;
    A.hits[h] = ph;
    SwTask rt;
    rt.q = tk.q; rt.t = tk.t; rt.nRows = r.row + 1; rt.nCols = r.col + 1;
    rt.limit = ph.word ? INT_MAX : A.limit8[tk.q]; rt.width = ph.word ? 16 : 32; rt.reverse = 1; rt.pad = (int32_t)h;
    A.revTasks[atomicAdd(A.nRev, 1u)] = rt;
}

// ssw_align_private :456-480 after the reverse pass: start coordinates, coverage, traceback task
__global__ void chainAfterReverse(ChainArgs A, const SwTask *tasks, const SwResult *res, uint32_t n, int bigPass) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const SwTask tk = tasks[i];
    const SwResult r = res[i];
    PairHit &ph = A.hits[tk.pad];
    if (r.overflow == 2 && !bigPass) { A.bigTasks[atomicAdd(A.nBig, 1u)] = tk; return; }
    if (r.overflow || r.score != ph.score) { atomicAdd(&A.counters[1], 1ull); ph.stage = 0; return; }
    const int Lq = (int)A.qLen[ph.q], Lt = (int)(A.tOff[ph.t + 1] - A.tOff[ph.t]);
    ph.tStart = ph.tEnd - r.col;
    ph.qStart = ph.qEnd - r.row;
    const float qcov = devCov((unsigned)ph.qStart, (unsigned)ph.qEnd, (unsigned)Lq), tcov = devCov((unsigned)ph.tStart, (unsigned)ph.tEnd, (unsigned)Lt);
    if (!devHasCoverage(A.covThr, A.covMode, qcov, tcov)) { ph.stage = 0; return; }
    ph.stage = 2;
    TbTask tb;
    tb.q = ph.q; tb.t = ph.t; tb.qStart = ph.qStart; tb.qEnd = ph.qEnd; tb.tStart = ph.tStart; tb.tEnd = ph.tEnd; tb.score = ph.score; tb.pad = (uint32_t)tk.pad;
    tb.btOff = atomicAdd(A.poolCursor, (unsigned long long)tbSlotBytes(tb));
    const uint32_t ti = atomicAdd(A.nTb, 1u);
    ph.tb = ti;
    A.tbTasks[ti] = tb;
}

__global__ void splitPrefilterHits(const sd_prefilter_hit *hits, uint64_t n, uint32_t *q, uint32_t *t) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { q[i] = hits[i].query; t[i] = hits[i].target; }
}

inline unsigned gridOf(uint64_t n, unsigned b) { return (unsigned)std::max<uint64_t>(1, (n + b - 1) / b); }

}  // namespace

// Device pair arrays in, accepted alignments out.  dPairQ/dPairT live on the device (grouped or not: the order of
// the output is defined by (query, Matcher::compareHits)).
// sum of target lengths over the (query, target) pairs of a call (reporting: algorithmic bytes of the score pass)
static __global__ void sumPairTargetResidues(const uint32_t *__restrict__ pairT, uint64_t n, const uint32_t *__restrict__ tOff,
                                             unsigned long long *__restrict__ out) {
    unsigned long long s = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t t = pairT[i];
        s += tOff[t + 1] - tOff[t];
    }
    for (int d = 16; d > 0; d >>= 1) s += __shfl_down_sync(0xffffffffu, s, d);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, s);
}

int sd_align_device_pairs(sd_context *ctx, const sd_target_index *idx, const sd_query_batch *queries, uint32_t query_first,
                          const uint32_t *dPairQ, const uint32_t *dPairT, uint64_t n_pairs, uint64_t db_residues,
                          const sd_align_params *par, std::vector<sd_alignment> &acc, sd::CharPool &pool,
                          sd_align_stats *stats) {
    acc.clear(); pool.clear();
    if (stats) memset(stats, 0, sizeof *stats);
    const uint32_t nq = queries->nq;
    if (nq == 0 || n_pairs == 0) return SD_OK;
    if (par->gap_open < par->gap_extend) { setError("sd_align_run: gap open < gap extend is not supported"); return SD_ELIMIT; }
    if (par->gap_open != 11 || par->gap_extend != 1) {
        // EvalueComputation::init (alignment/EvalueComputation.h:56-110): only BLOSUM62 (11, 1) has tabulated Gumbel parameters; for
        // any other gap costs the reference ESTIMATES them with ALP's simulation (lib/alp, seed 42, up to 60 s) -- E-values and bit
        // scores would silently differ here, so the call is refused
        setError("sd_align_run: --gap-open %d --gap-extend %d: the reference estimates the E-value parameters for non-default gap "
                 "costs by simulation (ALP); only 11/1 is implemented", par->gap_open, par->gap_extend);
        return SD_ELIMIT;
    }
    if (n_pairs >= 0xffffffffull) { setError("sd_align_run: too many pairs in one call (%llu)", (unsigned long long)n_pairs); return SD_ELIMIT; }
    const uint64_t total = queries->total;
    SD_CUDA(cudaSetDevice(ctx->device));
    sd::StreamScope allocScope(ctx->stream);
    cudaStream_t st = ctx->stream;
    // chunked clustersearch: hold the GPU token from here to the end of the traceback (released for the host tail)
    std::unique_lock<std::mutex> gpuTurn;
    if (ctx->gpuToken) gpuTurn = std::unique_lock<std::mutex>(*ctx->gpuToken);
    int rc;
    sd::EventQuad events;
    { const int erc = events.create(); if (erc) return erc; }
    cudaEvent_t ev0 = events.e[0], ev1 = events.e[1], evA = events.e[2], evB = events.e[3];
    SD_CUDA(cudaEventRecord(ev0, st));
    const bool timing = getenv("SD_TIMING") != nullptr;
    auto tPrev = std::chrono::steady_clock::now();
    std::string tLog;
    auto tick = [&](const char *name) {
        if (!timing) return;
        cudaStreamSynchronize(st);
        auto now = std::chrono::steady_clock::now();
        char b[96];
        snprintf(b, sizeof b, " %s=%.1f", name, std::chrono::duration<double, std::milli>(now - tPrev).count());
        tLog += b;
        tPrev = now;
    };
    double swMs = 0, tbMs = 0;
    auto lap = [&](double &accMs) -> int {
        float ms = 0;
        SD_CUDA(cudaEventSynchronize(evB));
        cudaEventElapsedTime(&ms, evA, evB);
        accMs += ms;
        return SD_OK;
    };

    // the alignment matrix follows --score-bias (Alignment.cpp:152)
    SubMat alnMat = ctx->alnMat;
    if (par->score_bias != 0.0f) buildSubMat(BLOSUM62, 2.0f, par->score_bias, alnMat);
    int16_t sub16[AA * AA];
    int8_t mat8[AA * AA];
    int minMat = 0;
    for (int i = 0; i < AA * AA; i++) { sub16[i] = alnMat.sub[i]; mat8[i] = (int8_t)alnMat.sub[i]; minMat = std::min(minMat, (int)mat8[i]); }

    // ---- query side on the device: composition bias, profiles
    const uint32_t nPos = (uint32_t)total;
    const std::vector<uint32_t> &hOff = queries->hostOff;
    std::vector<uint32_t> hStride(nq), hLen(nq);
    std::vector<uint64_t> hProfOff(nq + 1), hStgOff(nq + 1);
    std::vector<uint32_t> hStgStride(nq);
    hProfOff[0] = 0;
    hStgOff[0] = 0;
    uint32_t maxQ = 1;
    for (uint32_t i = 0; i < nq; i++) {
        hLen[i] = hOff[i + 1] - hOff[i];
        maxQ = std::max(maxQ, hLen[i]);
        hStride[i] = ((hLen[i] + 3u) & ~3u) + 72u;          // >= 68 bytes of -128 after every row
        hProfOff[i + 1] = hProfOff[i] + (uint64_t)(AA + 1) * hStride[i];   // + dummy-letter row
        int rp, rpPad, nStages;
        swStageLayout((int)hLen[i], rp, rpPad, nStages);
        hStgStride[i] = (uint32_t)(nStages * rpPad);                       // 0: no staged profile (32-lane classes)
        hStgOff[i + 1] = hStgOff[i] + (uint64_t)(AA + 1) * hStgStride[i];
    }
    const uint64_t stgBytes = hStgOff[nq] + 8192;
    const uint64_t profBytes = hProfOff[nq] + 8192;
    DevBuf<uint32_t> dStride, dLen, dNext, dStgStride;
    DevBuf<uint64_t> dProfOff, dStgOff;
    ScratchBuf<int8_t> dProfS(ctx, SL_AL_PROFS);
    DevBuf<float> dBias;
    DevBuf<int8_t> dCb, dMat8;
    ScratchBuf<int8_t> dProfF(ctx, SL_AL_PROFF), dProfR(ctx, SL_AL_PROFR);
    DevBuf<int16_t> dSub;
    DevBuf<double> dPBack;
    DevBuf<unsigned long long> dCells;
    if ((rc = dStride.alloc(nq)) || (rc = dLen.alloc(nq)) || (rc = dProfOff.alloc(nq + 1)) || (rc = dBias.alloc(total + 16)) ||
        (rc = dCb.alloc(total + 16)) || (rc = dMat8.alloc(AA * AA)) || (rc = dSub.alloc(AA * AA)) || (rc = dPBack.alloc(AA)) ||
        (rc = dProfF.alloc(profBytes)) || (rc = dProfR.alloc(profBytes)) || (rc = dNext.alloc(4)) || (rc = dCells.alloc(4)) ||
        (rc = dStgStride.alloc(nq)) || (rc = dStgOff.alloc(nq + 1)) || (rc = dProfS.alloc(stgBytes)))
        return rc;
    SD_CUDA(cudaMemcpyAsync(dStgStride.p, hStgStride.data(), nq * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    SD_CUDA(cudaMemcpyAsync(dStgOff.p, hStgOff.data(), (nq + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    SD_CUDA(cudaMemsetAsync(dProfS.p, 0x80, stgBytes, st));
    SD_CUDA(cudaMemcpyAsync(dStride.p, hStride.data(), nq * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    SD_CUDA(cudaMemcpyAsync(dLen.p, hLen.data(), nq * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    SD_CUDA(cudaMemcpyAsync(dProfOff.p, hProfOff.data(), (nq + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    SD_CUDA(cudaMemcpyAsync(dMat8.p, mat8, sizeof mat8, cudaMemcpyHostToDevice, st));
    SD_CUDA(cudaMemcpyAsync(dSub.p, sub16, sizeof sub16, cudaMemcpyHostToDevice, st));
    SD_CUDA(cudaMemcpyAsync(dPBack.p, alnMat.pBack, AA * sizeof(double), cudaMemcpyHostToDevice, st));
    SD_CUDA(cudaMemsetAsync(dCells.p, 0, 4 * sizeof(unsigned long long), st));
    if (n_pairs) {      // algorithmic bytes of the score pass: sum of target lengths over the pairs (L_t + 64 per pair)
        sumPairTargetResidues<<<(unsigned)std::min<uint64_t>((n_pairs + 255) / 256, 4096), 256, 0, st>>>(dPairT, n_pairs, idx->off.p, dCells.p + 3);
        ctx->launches++;
    }
    SD_CUDA(cudaMemsetAsync(dProfF.p, 0x80, profBytes, st));
    SD_CUDA(cudaMemsetAsync(dProfR.p, 0x80, profBytes, st));
    compositionBias<<<(nPos + 127) / 128, 128, 0, st>>>(queries->code.p, queries->off.p, nq, nPos, dSub.p, dPBack.p, par->comp_bias_corr, dBias.p);
    ctx->launches += 1;
    SD_CUDA(cudaGetLastError());
    if ((rc = launchAlignCompositionBias(ctx, dBias.p, nPos, dCb.p))) return rc;
    if ((rc = launchSwProfiles(ctx, queries->code.p, dCb.p, queries->off.p, nq, nPos, dMat8.p, dProfOff.p, dStride.p, dProfF.p, dProfR.p, dStgOff.p, dStgStride.p, dProfS.p))) return rc;
    std::vector<int8_t> hCb(total);
    const std::vector<uint8_t> &hCode = queries->hostCode;
    SD_CUDA(cudaMemcpyAsync(hCb.data(), dCb.p, total, cudaMemcpyDeviceToHost, st));
    SD_CUDA(cudaStreamSynchronize(st));

    tick("profiles");
    // per query: 8-bit limit (ssw_init :1283-1291) and the smallest raw score whose E-value passes -e
    const std::vector<uint32_t> &tOff = idx->hostOff;
    const Evaluer evaluer((double)(db_residues ? db_residues : idx->totalRes));
    std::vector<int32_t> limit8(nq), minScore(nq);
    {
        // pass marks depend only on (database size, -e, query length): kept with the context between calls
        if (ctx->minScoreDb != (double)(db_residues ? db_residues : idx->totalRes) || ctx->minScoreThr != par->evalue_thr) {
            ctx->minScoreByLen.clear();
            ctx->minScoreDb = (double)(db_residues ? db_residues : idx->totalRes);
            ctx->minScoreThr = par->evalue_thr;
        }
        std::vector<int32_t> &byLen = ctx->minScoreByLen;      // 0 = not computed yet, -2 = not monotone
        if (byLen.size() < (size_t)maxQ + 1) byLen.resize(maxQ + 1, 0);
        std::vector<uint8_t> need(maxQ + 1, 0);
        for (uint32_t q = 0; q < nq; q++) if (byLen[hLen[q]] == 0) need[hLen[q]] = 1;
        parallelFor(maxQ + 1, [&](uint64_t L) {
            if (!need[L] || L == 0) return;
            // E(s) falls monotonically in s (checked below); the first s with E <= thr is the pass mark
            int s = 1;
            while (s < 100000 && evaluer.evalue((double)s, (double)L) > par->evalue_thr) s++;
            bool mono = true;
            double prev = evaluer.evalue((double)s, (double)L);
            for (int k = 1; k <= 64; k++) {
                const double e = evaluer.evalue((double)(s + k), (double)L);
                if (e > prev || e > par->evalue_thr) mono = false;
                prev = e;
            }
            byLen[L] = mono ? s : -2;
        });
        std::atomic<int> bad{-1};
        parallelFor(nq, [&](uint64_t q) {
            int mn = 0;
            for (uint32_t i = hOff[q]; i < hOff[q + 1]; i++) mn = std::min(mn, (int)hCb[i]);
            limit8[q] = 255 - (std::abs(minMat) + std::abs(mn));
            minScore[q] = hLen[q] ? byLen[hLen[q]] : 1;
            if (minScore[q] == -2) bad = (int)hLen[q];
        });
        if (bad >= 0) { setError("sd_align_run: E-value is not monotone in the score for query length %d", bad.load()); return SD_ELIMIT; }
    }

    tick("minScore");
    // ---- device lists
    const uint32_t nP = (uint32_t)n_pairs;
    ScratchBuf<SwTask> dFwd(ctx, SL_AL_FWD), dWord(ctx, SL_AL_WORD), dBig(ctx, SL_AL_BIG), dRev(ctx, SL_AL_REV);
    ScratchBuf<SwResult> dRes(ctx, SL_AL_RES);
    ScratchBuf<PairHit> dHits(ctx, SL_AL_HITS);
    ScratchBuf<TbTask> dTb(ctx, SL_AL_TB);
    ScratchBuf<TbResult> dTbRes(ctx, SL_AL_TBRES);
    ScratchBuf<uint32_t> dIdent(ctx, SL_AL_IDENT), dOrder(ctx, SL_AL_ORDER);
    DevBuf<uint32_t> dCnt, dClass;
    DevBuf<int32_t> dLimit8, dMinScore;
    DevBuf<unsigned long long> dCounters;
    if ((rc = dFwd.alloc(nP)) || (rc = dWord.alloc(nP)) || (rc = dBig.alloc(nP)) || (rc = dRev.alloc(nP)) || (rc = dRes.alloc(nP)) ||
        (rc = dHits.alloc(nP)) || (rc = dTb.alloc(nP)) || (rc = dIdent.alloc(nP)) || (rc = dOrder.alloc(nP)) ||
        (rc = dCnt.alloc(16)) || (rc = dClass.alloc(18)) || (rc = dLimit8.alloc(nq)) || (rc = dMinScore.alloc(nq)) || (rc = dCounters.alloc(4)))
        return rc;
    SD_CUDA(cudaMemcpyAsync(dLimit8.p, limit8.data(), nq * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    SD_CUDA(cudaMemcpyAsync(dMinScore.p, minScore.data(), nq * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    SD_CUDA(cudaMemsetAsync(dCnt.p, 0, 16 * sizeof(uint32_t), st));
    SD_CUDA(cudaMemsetAsync(dClass.p, 0, 18 * sizeof(uint32_t), st));
    SD_CUDA(cudaMemsetAsync(dCounters.p, 0, 4 * sizeof(unsigned long long), st));
    enum { C_FWD = 0, C_WORD, C_BIG, C_REV, C_HITS, C_IDENT, C_TB };
    ChainArgs C;
    C.pairQ = dPairQ; C.pairT = dPairT; C.nPairs = n_pairs; C.qLen = dLen.p; C.tOff = idx->off.p; C.limit8 = dLimit8.p; C.minScore = dMinScore.p;
    C.queryFirst = query_first; C.sameDb = par->same_db; C.covThr = par->cov_thr; C.covMode = par->cov_mode;
    C.fwdTasks = dFwd.p; C.nFwd = dCnt.p + C_FWD; C.wordTasks = dWord.p; C.nWord = dCnt.p + C_WORD; C.bigTasks = dBig.p; C.nBig = dCnt.p + C_BIG;
    C.revTasks = dRev.p; C.nRev = dCnt.p + C_REV; C.hits = dHits.p; C.nHits = dCnt.p + C_HITS; C.identity = dIdent.p; C.nIdentity = dCnt.p + C_IDENT;
    C.tbTasks = dTb.p; C.nTb = dCnt.p + C_TB; C.poolCursor = dCounters.p + 2; C.classCount = dClass.p; C.counters = dCounters.p;

    SwBatch batch;
    batch.profFwd = dProfF.p; batch.profRev = dProfR.p; batch.profOff = dProfOff.p; batch.profStride = dStride.p; batch.qLen = dLen.p;
    batch.profStg = dProfS.p; batch.stgOff = dStgOff.p; batch.stgStride = dStgStride.p;
    batch.tRes = idx->raw.p; batch.tOff = idx->off.p; batch.gapOpen = par->gap_open; batch.gapExtend = par->gap_extend;
    static const uint32_t swCtasPerSm = getenv("SD_SW_CTAS") ? (uint32_t)atoi(getenv("SD_SW_CTAS")) : 5u;     // = __launch_bounds__ of swKernel16
    const uint32_t swWarps = (uint32_t)ctx->smCount * swCtasPerSm * swWarpsPerCta();
    const uint32_t bndStride = std::max<uint32_t>(idx->maxLen, 32u) + 32u;
    ScratchBuf<int> dBoundary(ctx, SL_AL_BOUND);
    if ((rc = dBoundary.alloc((size_t)swWarps * 6 * bndStride))) return rc;

    uint32_t hCnt[16];
    auto readCounts = [&]() -> int {
        SD_CUDA(cudaMemcpyAsync(hCnt, dCnt.p, sizeof hCnt, cudaMemcpyDeviceToHost, st));
        SD_CUDA(cudaStreamSynchronize(st));
        return SD_OK;
    };
    // one Smith-Waterman launch over a device task list (scheduling by class happens inside launchSw)
    auto runList = [&](SwTask *tasks, uint32_t n, int packed16, int keyed, unsigned long long *cellSlot) -> int {
        if (n == 0) return SD_OK;
        SD_CUDA(cudaEventRecord(evA, st));
        int r2 = launchSw(ctx, batch, tasks, dRes.p, n, dBoundary.p, bndStride, swWarps, dNext.p, cellSlot, packed16, keyed);
        if (r2) return r2;
        SD_CUDA(cudaEventRecord(evB, st));
        return lap(swMs);
    };

    chainBuildForward<<<gridOf(n_pairs, 256), 256, 0, st>>>(C);
    ctx->launches++;
    SD_CUDA(cudaGetLastError());
    if ((rc = readCounts())) return rc;
    tick("build");
    const uint32_t nFwd = hCnt[C_FWD];
    if ((rc = runList(dFwd.p, nFwd, 1, 1, dCells.p))) return rc;
    if (nFwd) { chainAfterForward<<<gridOf(nFwd, 256), 256, 0, st>>>(C, dFwd.p, dRes.p, nFwd, 0); ctx->launches++; }
    if ((rc = readCounts())) return rc;
    tick("fwd8");
    const uint32_t nWord = hCnt[C_WORD];
    if ((rc = runList(dWord.p, nWord, 1, 0, dCells.p))) return rc;
    if (nWord) { chainAfterForward<<<gridOf(nWord, 256), 256, 0, st>>>(C, dWord.p, dRes.p, nWord, 1); ctx->launches++; }
    if ((rc = readCounts())) return rc;
    uint32_t nBig = hCnt[C_BIG];
    if (nBig) {     // scores near 2^15: 32-bit pass
        if ((rc = runList(dBig.p, nBig, 0, 0, dCells.p))) return rc;
        chainAfterForward<<<gridOf(nBig, 256), 256, 0, st>>>(C, dBig.p, dRes.p, nBig, 1);
        ctx->launches++;
        SD_CUDA(cudaMemsetAsync(dCnt.p + C_BIG, 0, sizeof(uint32_t), st));
        if ((rc = readCounts())) return rc;
    }
    tick("fwd16");
    const uint32_t nRev = hCnt[C_REV];
    if ((rc = runList(dRev.p, nRev, 1, 0, dCells.p + 1))) return rc;
    if (nRev) { chainAfterReverse<<<gridOf(nRev, 256), 256, 0, st>>>(C, dRev.p, dRes.p, nRev, 0); ctx->launches++; }
    if ((rc = readCounts())) return rc;
    nBig = hCnt[C_BIG];
    if (nBig) {
        if ((rc = runList(dBig.p, nBig, 0, 0, dCells.p + 1))) return rc;
        chainAfterReverse<<<gridOf(nBig, 256), 256, 0, st>>>(C, dBig.p, dRes.p, nBig, 1);
        ctx->launches++;
        if ((rc = readCounts())) return rc;
    }
    tick("rev");
    const uint32_t nHit = hCnt[C_HITS], nIdent = hCnt[C_IDENT], nTb = hCnt[C_TB];
    unsigned long long hCounters[4];
    SD_CUDA(cudaMemcpyAsync(hCounters, dCounters.p, sizeof hCounters, cudaMemcpyDeviceToHost, st));
    SD_CUDA(cudaStreamSynchronize(st));
    if (hCounters[1]) {
        // the reference aborts in this situation ("Score of forward/backward SW differ")
        setError("sd_align_run: forward and reverse Smith-Waterman scores differ for %llu pairs", hCounters[1]);
        return SD_ELIMIT;
    }
    const uint64_t poolNeed = hCounters[2];

    // ---- banded traceback on the survivors: CIGAR strings, compacted on the device, one copy to the host
    std::vector<TbResult> tres(nTb);
    std::vector<unsigned long long> cigOff(nTb + 1, 0);
    char *devPool = nullptr;
    if (nTb) {
        ScratchBuf<char> dPool(ctx, SL_TB_POOL), dDense(ctx, SL_TB_REV);
        ScratchBuf<int8_t> dDir(ctx, SL_TB_DIR);
        ScratchBuf<int> dRows(ctx, SL_TB_ROWS);
        DevBuf<uint32_t> dPend, dCigLen;
        DevBuf<unsigned long long> dCigOff;
        if ((rc = dTbRes.alloc(nTb)) || (rc = dPool.alloc(poolNeed + 1)) || (rc = dCigLen.alloc(nTb + 1)) || (rc = dCigOff.alloc(nTb + 1))) return rc;
        SD_CUDA(cudaMemsetAsync(dCigLen.p, 0, ((size_t)nTb + 1) * sizeof(uint32_t), st));
        const uint32_t maxL = std::max(maxQ, idx->maxLen);
        TbArgs A;
        A.qCode = queries->code.p; A.qCb = dCb.p; A.qOff = queries->off.p; A.tRes = idx->raw.p; A.tOff = idx->off.p; A.mat = dMat8.p;
        A.gapOpen = par->gap_open; A.gapExtend = par->gap_extend; A.tasks = dTb.p; A.results = dTbRes.p; A.cigLen = dCigLen.p; A.next = dNext.p;
        A.btPool = dPool.p; A.cells = dCells.p + 2;
        std::vector<uint32_t> pending;
        for (int pass = 0; pass < 2; pass++) {
            const uint32_t nTasks = pass == 0 ? nTb : (uint32_t)pending.size();
            if (nTasks == 0) break;
            uint32_t nWarps;
            uint64_t dirCap;
            uint32_t rowCap;
            if (pass == 0) {
                nWarps = std::min<uint32_t>((uint32_t)ctx->smCount * 8u * TB_WARPS, (nTasks + 7) / 8);     // 8 CTAs per SM: the row loop is latency bound
                nWarps = std::max<uint32_t>(nWarps, 1);
                dirCap = 1ull << 20;
                rowCap = 4096;
            } else {
                nWarps = std::min<uint32_t>(64u, nTasks);
                const uint64_t worstBand = 4ull * maxL + 8ull;
                dirCap = std::min<uint64_t>((2 * worstBand + 1) * (uint64_t)maxL, 1ull << 30);
                rowCap = (uint32_t)std::min<uint64_t>(2 * worstBand + 8, 1u << 20);
            }
            nWarps = ((nWarps + TB_WARPS - 1) / TB_WARPS) * TB_WARPS;
            if ((rc = dDir.alloc((size_t)nWarps * dirCap)) || (rc = dRows.alloc((size_t)nWarps * 3 * rowCap))) return rc;
            A.dir = dDir.p; A.dirCap = dirCap; A.rows = dRows.p; A.rowCap = rowCap;
            A.nTasks = nTasks;
            if (pass == 0) A.order = nullptr;
            else {
                if ((rc = dPend.alloc(pending.size()))) return rc;
                SD_CUDA(cudaMemcpyAsync(dPend.p, pending.data(), pending.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
                A.order = dPend.p;
            }
            SD_CUDA(cudaMemsetAsync(dNext.p, 0, sizeof(uint32_t), st));
            SD_CUDA(cudaEventRecord(evA, st));
            tracebackKernel<<<nWarps / TB_WARPS, TB_WARPS * 32, 0, st>>>(A);
            ctx->launches++;
            SD_CUDA(cudaGetLastError());
            SD_CUDA(cudaEventRecord(evB, st));
            SD_CUDA(cudaMemcpyAsync(tres.data(), dTbRes.p, (size_t)nTb * sizeof(TbResult), cudaMemcpyDeviceToHost, st));
            SD_CUDA(cudaStreamSynchronize(st));
            if ((rc = lap(tbMs))) return rc;
            pending.clear();
            for (uint32_t i = 0; i < nTb; i++)
                if (tres[i].status == 1) pending.push_back(i);
            if (pass == 1 && !pending.empty()) { setError("sd_align_run: %zu tracebacks exceed the scratch limit", pending.size()); return SD_ELIMIT; }
        }
        for (uint32_t i = 0; i < nTb; i++)
            if (tres[i].status != 0) { setError("sd_align_run: a traceback left the band (inconsistent start/end)"); return SD_ELIMIT; }
        // dense pool: offsets = exclusive scan of the text lengths, one gather, one device->host copy
        {
            cub::TransformInputIterator<unsigned long long, U32ToU64, const uint32_t *> lenIt(dCigLen.p, U32ToU64());
            size_t tmpBytes = 0;
            SD_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, lenIt, dCigOff.p, (int64_t)nTb + 1, st));
            DevBuf<uint8_t> tmp;
            if ((rc = tmp.alloc(tmpBytes))) return rc;
            SD_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmpBytes, lenIt, dCigOff.p, (int64_t)nTb + 1, st));
            SD_CUDA(cudaMemcpyAsync(cigOff.data(), dCigOff.p, ((size_t)nTb + 1) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
            SD_CUDA(cudaStreamSynchronize(st));
            const unsigned long long denseBytes = cigOff[nTb];
            if ((rc = dDense.alloc(denseBytes + 1))) return rc;
            devPool = static_cast<char *>(ctxPinnedScratch(ctx, 0, denseBytes + 1));
            if (!devPool) return SD_ENOMEM;
            gatherCigars<<<gridOf((uint64_t)nTb * 32, 256), 256, 0, st>>>(dTb.p, dTbRes.p, dCigOff.p, nTb, dPool.p, dDense.p);
            ctx->launches++;
            SD_CUDA(cudaGetLastError());
            SD_CUDA(cudaMemcpyAsync(devPool, dDense.p, denseBytes, cudaMemcpyDeviceToHost, st));
            SD_CUDA(cudaStreamSynchronize(st));
        }
    }
    tick("traceback");
    if (gpuTurn.owns_lock()) gpuTurn.unlock();
    // ---- survivors and identity pairs to the host: E-values, criteria, order
    std::vector<PairHit> hHits(nHit);
    std::vector<uint32_t> hIdent(nIdent);
    if (nHit) SD_CUDA(cudaMemcpyAsync(hHits.data(), dHits.p, (size_t)nHit * sizeof(PairHit), cudaMemcpyDeviceToHost, st));
    if (nIdent) SD_CUDA(cudaMemcpyAsync(hIdent.data(), dIdent.p, (size_t)nIdent * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    SD_CUDA(cudaStreamSynchronize(st));
    const std::vector<uint32_t> &identQ = hIdent;

    // every stage-2 hit owns one traceback task: slot = task index (order is fixed by the sort below)
    std::vector<sd_alignment> traced(nTb);
    std::vector<uint8_t> tracedOk(nTb, 0);
    parallelFor(nHit, [&](uint64_t i) {
        const PairHit &h = hHits[i];
        if (h.stage != 2 || h.tb >= nTb) return;
        const TbResult &tr = tres[h.tb];
        sd_alignment a;
        memset(&a, 0, sizeof a);
        a.query = h.q; a.target = h.t; a.score = h.score;
        a.q_start = h.qStart; a.q_end = h.qEnd; a.q_len = (int)hLen[h.q];
        a.t_start = h.tStart; a.t_end = h.tEnd; a.t_len = (int)(tOff[h.t + 1] - tOff[h.t]);
        a.qcov = covOf((unsigned)h.qStart, (unsigned)h.qEnd, (unsigned)a.q_len);
        a.tcov = covOf((unsigned)h.tStart, (unsigned)h.tEnd, (unsigned)a.t_len);
        a.aln_len = tr.btLen;
        a.seq_id = (float)tr.ids / (float)tr.btLen;
        a.bt_len = tr.cigLen;           // the pool holds the run-length CIGAR text
        a.bt_off = 0;
        a.bt_cigar = h.tb;              // remember the task: its text is fetched below
        a.evalue = evaluer.evalue(a.score, a.q_len);
        a.bits = evaluer.bitScore(a.score);
        // checkCriteria (Alignment.cpp:548-567)
        tracedOk[h.tb] = (a.evalue <= par->evalue_thr) && (a.seq_id >= par->min_seq_id) && hasCoverage(par->cov_thr, par->cov_mode, a.qcov, a.tcov) &&
                         (a.aln_len >= par->min_aln_len);
        traced[h.tb] = a;
    });
    tick("gather");
    // identity shortcut (scoreIdentical StripedSmithWaterman.cpp:1675-1710; note the 16-bit accumulator)
    std::vector<sd_alignment> ident(nIdent);
    parallelFor(nIdent, [&](uint64_t i) {
        const uint32_t q = identQ[i], t = query_first + q;
        const int Lq = (int)hLen[q], Lt = (int)(tOff[t + 1] - tOff[t]);
        short sc = 0;
        const uint8_t *c = hCode.data() + hOff[q];
        const int8_t *b = hCb.data() + hOff[q];
        for (int k = 0; k < Lq && k < Lt; k++) sc = (short)(sc + mat8[c[k] * AA + c[k]] + b[k]);
        sd_alignment a;
        memset(&a, 0, sizeof a);
        a.query = q; a.target = t; a.score = sc; a.evalue = evaluer.evalue(sc, Lq); a.bits = evaluer.bitScore(sc);
        a.q_start = 0; a.t_start = 0; a.q_end = Lt - 1; a.t_end = Lt - 1; a.q_len = Lq; a.t_len = Lt;
        a.qcov = 1.0f; a.tcov = 1.0f; a.seq_id = 1.0f; a.aln_len = Lt; a.bt_cigar = 0xffffffffu;
        char tmpCig[16];
        a.bt_len = (uint32_t)snprintf(tmpCig, sizeof tmpCig, "%dM", Lt);
        ident[i] = a;
    });
    // bucket by query (counting sort), then order every query's list with Matcher::compareHits in parallel
    {
        std::vector<uint32_t> start(nq + 1, 0);
        for (uint32_t i = 0; i < nTb; i++) if (tracedOk[i]) start[traced[i].query + 1]++;
        for (uint32_t i = 0; i < nIdent; i++) start[ident[i].query + 1]++;
        for (uint32_t q = 0; q < nq; q++) start[q + 1] += start[q];
        acc.resize(start[nq]);
        std::vector<uint32_t> cur(start.begin(), start.end() - 1);
        for (uint32_t i = 0; i < nTb; i++) if (tracedOk[i]) acc[cur[traced[i].query]++] = traced[i];
        for (uint32_t i = 0; i < nIdent; i++) acc[cur[ident[i].query]++] = ident[i];
        parallelFor(nq, [&](uint64_t q) {
            std::sort(acc.begin() + start[q], acc.begin() + start[q + 1], [](const sd_alignment &x, const sd_alignment &y) {
                if (x.evalue != y.evalue) return x.evalue < y.evalue;          // Matcher::compareHits
                if (x.bits != y.bits) return x.bits > y.bits;
                if (x.t_len != y.t_len) return x.t_len < y.t_len;
                return x.target < y.target;
            });
        });
    }
    tick("evalue+sort");
    uint64_t poolBytes = 0;
    for (const sd_alignment &a : acc) poolBytes += a.bt_len;
    pool.resize(poolBytes);
    {
        uint64_t w = 0;
        for (sd_alignment &a : acc) { a.bt_off = w; w += a.bt_len; }
        parallelFor(acc.size(), [&](uint64_t i) {
            sd_alignment &a = acc[i];
            if (a.bt_cigar == 0xffffffffu) {          // identity shortcut: one run of matches
                char tmpCig[16];
                const int n = snprintf(tmpCig, sizeof tmpCig, "%dM", a.aln_len);
                memcpy(pool.data() + a.bt_off, tmpCig, (size_t)n);
            } else memcpy(pool.data() + a.bt_off, devPool + cigOff[a.bt_cigar], a.bt_len);
            a.bt_cigar = 1;
        });
    }
    tick("pool");
    if (timing) fprintf(stderr, "[sd_align]%s\n", tLog.c_str());
    SD_CUDA(cudaEventRecord(ev1, st));
    SD_CUDA(cudaStreamSynchronize(st));
    if (stats) {
        unsigned long long hc[4];
        SD_CUDA(cudaMemcpyAsync(hc, dCells.p, sizeof hc, cudaMemcpyDeviceToHost, st));
        SD_CUDA(cudaStreamSynchronize(st));
        float ms = 0;
        cudaEventElapsedTime(&ms, ev0, ev1);
        stats->alignments = hCounters[0]; stats->passed = acc.size(); stats->word_mode = nWord;
        stats->reverse_passes = nRev; stats->tracebacks = nTb;
        stats->cells_forward = hc[0]; stats->cells_reverse = hc[1]; stats->cells_traceback = hc[2]; stats->target_residues = hc[3];
        stats->sw_ms = swMs; stats->traceback_ms = tbMs; stats->total_ms = ms;
    }
    return SD_OK;
}

extern "C" int sd_align_run_batch(sd_context *ctx, const sd_target_index *idx, const sd_query_batch *queries, uint32_t query_first,
                                  const uint32_t *pair_q, const uint32_t *pair_t, uint64_t n_pairs, uint64_t db_residues,
                                  const sd_align_params *par, sd_alignment *out, uint64_t cap, uint64_t *n_out, char *bt_pool,
                                  uint64_t bt_cap, uint64_t *bt_used, sd_align_stats *stats) {
    if (!ctx || !idx || !queries || !par || !n_out || !bt_used || (n_pairs && (!pair_q || !pair_t))) { setError("sd_align_run_batch: bad argument"); return SD_EINVAL; }
    *n_out = 0; *bt_used = 0;
    if (stats) memset(stats, 0, sizeof *stats);
    if (queries->nq == 0 || n_pairs == 0) return SD_OK;
    for (uint64_t k = 0; k < n_pairs; k++)
        if (pair_q[k] >= queries->nq || pair_t[k] >= idx->nT) { setError("sd_align_run: pair %llu out of range", (unsigned long long)k); return SD_EINVAL; }
    SD_CUDA(cudaSetDevice(ctx->device));
    sd::StreamScope allocScope(ctx->stream);
    // identity of a call (FNV-1a over the arguments, the sizes and a sample of the pairs and query residues), computed only when an
    // SD_ESMALL result is pending or has to be parked: the retry with larger buffers finds its result in the context
    auto callKey = [&]() -> uint64_t {
        uint64_t k = 1469598103934665603ull;
        auto mix = [&](const void *p, size_t n) { const unsigned char *b = static_cast<const unsigned char *>(p); for (size_t i = 0; i < n; i++) { k ^= b[i]; k *= 1099511628211ull; } };
        const void *ptrs[1] = {idx};          // (the query batch is identified by its content: sd_align_run builds a new one per call)
        mix(ptrs, sizeof ptrs); mix(&query_first, sizeof query_first); mix(&n_pairs, sizeof n_pairs); mix(&db_residues, sizeof db_residues);
        mix(par, sizeof *par); mix(&queries->nq, sizeof queries->nq); mix(&queries->total, sizeof queries->total);
        const uint64_t sample = std::min<uint64_t>(n_pairs, 4096);
        mix(pair_q, sample * sizeof(uint32_t)); mix(pair_t, sample * sizeof(uint32_t));
        mix(pair_q + (n_pairs - sample), sample * sizeof(uint32_t)); mix(pair_t + (n_pairs - sample), sample * sizeof(uint32_t));
        const size_t rs = std::min<size_t>(queries->hostCode.size(), 4096);
        mix(queries->hostCode.data(), rs); mix(queries->hostCode.data() + (queries->hostCode.size() - rs), rs);
        return k ? k : 1;
    };
    const uint64_t key = ctx->esmallKey ? callKey() : 0;
    std::vector<sd_alignment> acc;
    sd::CharPool pool;
    int rc;
    if (ctx->esmallKey && ctx->esmallKey == key) {
        acc.swap(ctx->esmallAlns);
        pool.swap(ctx->esmallPool);
        if (stats) *stats = ctx->esmallStats;
        ctx->esmallKey = 0;
    } else {
        ctx->esmallKey = 0;
        ctx->esmallAlns.clear(); ctx->esmallAlns.shrink_to_fit();
        ctx->esmallPool.clear(); ctx->esmallPool.shrink_to_fit();
        ScratchBuf<uint32_t> dQ(ctx, SL_AL_PAIRQ), dT(ctx, SL_AL_PAIRT);
        if ((rc = dQ.alloc(n_pairs)) || (rc = dT.alloc(n_pairs))) return rc;
        SD_CUDA(cudaMemcpyAsync(dQ.p, pair_q, n_pairs * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
        SD_CUDA(cudaMemcpyAsync(dT.p, pair_t, n_pairs * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
        sd_align_stats localStats;
        if ((rc = sd_align_device_pairs(ctx, idx, queries, query_first, dQ.p, dT.p, n_pairs, db_residues, par, acc, pool, &localStats))) return rc;
        if (stats) *stats = localStats;
        ctx->esmallStats = localStats;
    }
    uint64_t used = 0;
    for (const sd_alignment &a : acc) used += a.bt_len;
    *n_out = acc.size();
    *bt_used = used;
    if (acc.size() > cap || used > bt_cap || (acc.size() && (!out || !bt_pool))) {
        setError("sd_align_run: needs %zu alignments / %llu backtrace bytes", acc.size(), (unsigned long long)used);
        ctx->esmallAlns.swap(acc);       // kept for the retry with the same arguments
        ctx->esmallPool.swap(pool);
        ctx->esmallKey = callKey();
        return SD_ESMALL;
    }
    if (!acc.empty()) memcpy(out, acc.data(), acc.size() * sizeof(sd_alignment));
    if (used) memcpy(bt_pool, pool.data(), used);
    return SD_OK;
}

int sd_align_device_hits(sd_context *ctx, const sd_target_index *idx, const sd_query_batch *queries, uint32_t query_first,
                         const sd_prefilter_hit *dHits, uint64_t n_hits, uint64_t db_residues, const sd_align_params *par,
                         std::vector<sd_alignment> &acc, sd::CharPool &pool, sd_align_stats *stats) {
    acc.clear(); pool.clear();
    if (stats) memset(stats, 0, sizeof *stats);
    if (n_hits == 0) return SD_OK;
    SD_CUDA(cudaSetDevice(ctx->device));
    sd::StreamScope allocScope(ctx->stream);
    ScratchBuf<uint32_t> dQ(ctx, SL_AL_PAIRQ), dT(ctx, SL_AL_PAIRT);
    int rc;
    if ((rc = dQ.alloc(n_hits)) || (rc = dT.alloc(n_hits))) return rc;
    splitPrefilterHits<<<gridOf(n_hits, 256), 256, 0, ctx->stream>>>(dHits, n_hits, dQ.p, dT.p);
    ctx->launches++;
    SD_CUDA(cudaGetLastError());
    return sd_align_device_pairs(ctx, idx, queries, query_first, dQ.p, dT.p, n_hits, db_residues, par, acc, pool, stats);
}

extern "C" int sd_format_alignment(const sd_alignment *a, uint32_t target_key, const char *bt_pool, char *buf, int cap) {
    if (!a || !buf || cap < 160) return 0;
    int n = snprintf(buf, (size_t)cap, "%u\t%d\t", target_key, a->bits);
    // Util::fastSeqIdToBuffer (Util.cpp:222-251)
    if (a->seq_id == 1.0f) n += snprintf(buf + n, (size_t)(cap - n), "1.00");
    else {
        buf[n++] = '0'; buf[n++] = '.';
        if (a->seq_id < 0.10) buf[n++] = '0';
        if (a->seq_id < 0.01) buf[n++] = '0';
        n += snprintf(buf + n, (size_t)(cap - n), "%d", (int)(a->seq_id * 1000));
    }
    n += snprintf(buf + n, (size_t)(cap - n), "\t%.3E\t%d\t%d\t%d\t%d\t%d\t%d\t", a->evalue, a->q_start, a->q_end, a->q_len,
                  a->t_start, a->t_end, a->t_len);
    const char *bt = bt_pool + a->bt_off;
    if (a->bt_cigar) {            // already run-length encoded on the device
        if (n + (int)a->bt_len + 2 > cap) return 0;
        memcpy(buf + n, bt, a->bt_len);
        n += (int)a->bt_len;
        buf[n++] = '\n';
        buf[n] = '\0';
        return n;
    }
    // Matcher::compressAlignment (Matcher.cpp:166-185)
    char state = 'M';
    unsigned count = 0;
    for (uint32_t i = 0; i < a->bt_len; i++) {
        if (n + 24 > cap) return 0;
        if (bt[i] != state) {
            n += snprintf(buf + n, (size_t)(cap - n), "%u%c", count, state);
            state = bt[i];
            count = 1;
        } else count++;
    }
    if (n + 24 > cap) return 0;
    n += snprintf(buf + n, (size_t)(cap - n), "%u%c\n", count, state);
    return n;
}
