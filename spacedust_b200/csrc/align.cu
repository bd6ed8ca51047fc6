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
#include <climits>
#include <cmath>
#include <cstring>
#include <memory>
#include <thread>

#include "sd_internal.h"
#include "prep.cuh"
#include "sw.cuh"

namespace sd {

// ------------------------------------------------------------------------------------------------
// banded traceback: one warp per alignment, lanes across the band of one query row
// ------------------------------------------------------------------------------------------------
struct TbTask {
    uint32_t q, t;
    int32_t qStart, qEnd, tStart, tEnd, score;
    uint32_t pad;
    uint64_t btOff;
};

struct TbResult {
    int32_t btLen;
    int32_t ids;
    int32_t status;    // 0 ok, 1 scratch too small, 2 trace walked off the band
    int32_t band;
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
    uint32_t nTasks;
    uint32_t *next;
    // per-warp scratch
    int8_t *dir;
    uint64_t dirCap;
    int *rows;               // 3 x rowCap ints
    uint32_t rowCap;
    char *rev;
    uint32_t revCap;
    char *btPool;
    unsigned long long *cells;
};

constexpr int TB_WARPS = 4;
constexpr int TB_NEG = -(1 << 28);

__global__ void __launch_bounds__(TB_WARPS * 32) tracebackKernel(TbArgs A) {
    __shared__ int8_t sMat[AA * AA];
    for (int i = threadIdx.x; i < AA * AA; i += blockDim.x) sMat[i] = A.mat[i];
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = blockIdx.x * TB_WARPS + (threadIdx.x >> 5);
    int8_t *dir = A.dir + (size_t)warp * A.dirCap;
    int *hb = A.rows + (size_t)warp * 3 * A.rowCap;
    int *eb = hb + A.rowCap;
    int *hc = eb + A.rowCap;
    char *rev = A.rev + (size_t)warp * A.revCap;
    const int go = A.gapOpen, ge = A.gapExtend;
    unsigned long long cells = 0;

    for (;;) {
        uint32_t ti = 0;
        if (lane == 0) ti = atomicAdd(A.next, 1u);
        ti = __shfl_sync(0xffffffffu, ti, 0);
        if (ti >= A.nTasks) break;
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
        for (;;) {
            const long width = (long)band * 2 + 3;
            widthD = (long)band * 2 + 1;
            if ((unsigned long long)widthD * (unsigned long long)qLen * 3ull > A.dirCap || (unsigned long long)width + 2 > A.rowCap) { status = 1; break; }
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
                int8_t *line = dir + widthD * (long)i * 3;
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
                    const int8_t de = t1 > t2 ? 3 : 2;
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
                    const int8_t df = (hPrev - go > fPrev - ge) ? 5 : 4;
                    const int8_t dh = (g1 <= diagv) ? 1 : (e1 > f1 ? de : df);
                    if (valid) {
                        eb[u] = e;
                        hc[u] = h;
                        int8_t *cell = line + (long)(j - xi) * 3;
                        cell[0] = de; cell[1] = df; cell[2] = dh;
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
        int n = 0, ids = 0;
        if (status == 0) {
            if (lane == 0) {
                int i = qLen - 1, j = dbLen - 1, state = 2;
                while (i > 0 || j > 0) {
                    if (i < 0 || j < 0) { status = 2; break; }
                    const int xi = max(i - band, 0);
                    const long x = j - xi;
                    if (x < 0 || x >= widthD) { status = 2; break; }
                    const int8_t d = dir[widthD * (long)i * 3 + x * 3 + state];
                    char op;
                    if (d == 1) { ids += (qc[i] == tc[j]); --i; --j; state = 2; op = 'M'; }
                    else if (d == 2) { --i; state = 0; op = 'I'; }
                    else if (d == 3) { --i; state = 2; op = 'I'; }
                    else if (d == 4) { --j; state = 1; op = 'D'; }
                    else if (d == 5) { --j; state = 2; op = 'D'; }
                    else { status = 2; break; }
                    if ((uint32_t)n + 2 > A.revCap) { status = 2; break; }
                    rev[n++] = op;
                }
                if (status == 0) { ids += (qc[0] == tc[0]); rev[n++] = 'M'; }
            }
            n = __shfl_sync(0xffffffffu, n, 0);
            status = __shfl_sync(0xffffffffu, status, 0);
            __syncwarp();
            if (status == 0) {
                char *out = A.btPool + tk.btOff;
                for (int k = (int)lane; k < n; k += 32) out[k] = rev[n - 1 - k];
            }
        }
        if (lane == 0) {
            TbResult r;
            r.btLen = n; r.ids = ids; r.status = status; r.band = band;
            A.results[id] = r;
        }
        __syncwarp();
    }
    if (lane == 0 && cells) atomicAdd(A.cells, cells);
}

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

template <typename F>
static void parallelFor(uint64_t n, F fn) {
    unsigned nThreads = std::max(1u, std::min(std::thread::hardware_concurrency(), 64u));
    if (n < 4096) nThreads = 1;
    std::atomic<uint64_t> next{0};
    const uint64_t chunk = 2048;
    auto work = [&]() {
        for (;;) {
            const uint64_t lo = next.fetch_add(chunk);
            if (lo >= n) break;
            const uint64_t hi = std::min(n, lo + chunk);
            for (uint64_t i = lo; i < hi; i++) fn(i);
        }
    };
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < nThreads; t++) pool.emplace_back(work);
    work();
    for (auto &t : pool) t.join();
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

namespace {
struct PairState {
    uint32_t q, t;
    int score = 0, qEnd = -1, tEnd = -1, qStart = -1, tStart = -1;
    int word = 0;
    double evalue = 0;
    float qcov = 0, tcov = 0;
    int stage = 0;      // 0 dropped, 1 needs reverse, 2 needs traceback, 3 accepted
    int btLen = 0, ids = 0;
    uint64_t btOff = 0;
    bool identity = false;
};
}  // namespace

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

extern "C" int sd_align_run_batch(sd_context *ctx, const sd_target_index *idx, const sd_query_batch *queries, uint32_t query_first,
                                  const uint32_t *pair_q, const uint32_t *pair_t, uint64_t n_pairs, uint64_t db_residues,
                                  const sd_align_params *par, sd_alignment *out, uint64_t cap, uint64_t *n_out, char *bt_pool,
                                  uint64_t bt_cap, uint64_t *bt_used, sd_align_stats *stats) {
    if (!ctx || !idx || !queries || !par || !n_out || !bt_used || (n_pairs && (!pair_q || !pair_t))) { setError("sd_align_run_batch: bad argument"); return SD_EINVAL; }
    *n_out = 0; *bt_used = 0;
    if (stats) memset(stats, 0, sizeof *stats);
    const uint32_t nq = queries->nq;
    if (nq == 0 || n_pairs == 0) return SD_OK;
    if (par->gap_open < par->gap_extend) { setError("sd_align_run: gap open < gap extend is not supported"); return SD_ELIMIT; }
    if (n_pairs >= 0xffffffffull) { setError("sd_align_run: too many pairs in one call (%llu)", (unsigned long long)n_pairs); return SD_ELIMIT; }
    const uint64_t total = queries->total;
    for (uint64_t k = 0; k < n_pairs; k++) {
        if (pair_q[k] >= nq || pair_t[k] >= idx->nT) { setError("sd_align_run: pair %llu out of range", (unsigned long long)k); return SD_EINVAL; }
        if (k && pair_q[k] < pair_q[k - 1]) { setError("sd_align_run: pairs must be grouped by ascending query"); return SD_EINVAL; }
    }
    SD_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    int rc;
    cudaEvent_t ev0, ev1, evA, evB;
    SD_CUDA(cudaEventCreate(&ev0)); SD_CUDA(cudaEventCreate(&ev1)); SD_CUDA(cudaEventCreate(&evA)); SD_CUDA(cudaEventCreate(&evB));
    SD_CUDA(cudaEventRecord(ev0, st));
    double swMs = 0, tbMs = 0;
    auto lap = [&](double &acc) -> int {
        float ms = 0;
        SD_CUDA(cudaEventSynchronize(evB));
        cudaEventElapsedTime(&ms, evA, evB);
        acc += ms;
        return SD_OK;
    };

    // the alignment matrix follows --score-bias (Alignment.cpp:152)
    SubMat alnMat = ctx->alnMat;
    if (par->score_bias != 0.0f) buildSubMat(BLOSUM62, 2.0f, par->score_bias, alnMat);
    int16_t sub16[AA * AA];
    int8_t mat8[AA * AA];
    int minMat = 0;
    for (int i = 0; i < AA * AA; i++) { sub16[i] = alnMat.sub[i]; mat8[i] = (int8_t)alnMat.sub[i]; minMat = std::min(minMat, (int)mat8[i]); }

    // ---- query side on the device
    const uint32_t nPos = (uint32_t)total;
    const std::vector<uint32_t> &hOff = queries->hostOff;
    std::vector<uint32_t> hStride(nq), hLen(nq);
    std::vector<uint64_t> hProfOff(nq + 1);
    hProfOff[0] = 0;
    for (uint32_t i = 0; i < nq; i++) {
        hLen[i] = hOff[i + 1] - hOff[i];
        hStride[i] = ((hLen[i] + 3u) & ~3u) + 72u;          // >= 68 bytes of -128 after every row
        hProfOff[i + 1] = hProfOff[i] + (uint64_t)(AA + 1) * hStride[i];   // + dummy-letter row
    }
    const uint64_t profBytes = hProfOff[nq] + 8192;
    struct { const uint8_t *p; } dCode{queries->code.p};
    struct { const uint32_t *p; } dOff{queries->off.p};
    DevBuf<uint32_t> dStride, dLen, dNext;
    DevBuf<uint64_t> dProfOff;
    DevBuf<float> dBias;
    DevBuf<int8_t> dCb, dMat8, dProfF, dProfR;
    DevBuf<int16_t> dSub;
    DevBuf<double> dPBack;
    DevBuf<unsigned long long> dCells;
    if ((rc = dStride.alloc(nq)) || (rc = dLen.alloc(nq)) || (rc = dProfOff.alloc(nq + 1)) || (rc = dBias.alloc(total + 16)) ||
        (rc = dCb.alloc(total + 16)) || (rc = dMat8.alloc(AA * AA)) || (rc = dSub.alloc(AA * AA)) || (rc = dPBack.alloc(AA)) ||
        (rc = dProfF.alloc(profBytes)) || (rc = dProfR.alloc(profBytes)) || (rc = dNext.alloc(4)) || (rc = dCells.alloc(4)))
        return rc;
    SD_CUDA(cudaMemcpyAsync(dStride.p, hStride.data(), nq * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    SD_CUDA(cudaMemcpyAsync(dLen.p, hLen.data(), nq * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    SD_CUDA(cudaMemcpyAsync(dProfOff.p, hProfOff.data(), (nq + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    SD_CUDA(cudaMemcpyAsync(dMat8.p, mat8, sizeof mat8, cudaMemcpyHostToDevice, st));
    SD_CUDA(cudaMemcpyAsync(dSub.p, sub16, sizeof sub16, cudaMemcpyHostToDevice, st));
    SD_CUDA(cudaMemcpyAsync(dPBack.p, alnMat.pBack, AA * sizeof(double), cudaMemcpyHostToDevice, st));
    SD_CUDA(cudaMemsetAsync(dCells.p, 0, 4 * sizeof(unsigned long long), st));
    SD_CUDA(cudaMemsetAsync(dProfF.p, 0x80, profBytes, st));
    SD_CUDA(cudaMemsetAsync(dProfR.p, 0x80, profBytes, st));
    compositionBias<<<(nPos + 127) / 128, 128, 0, st>>>(dCode.p, dOff.p, nq, nPos, dSub.p, dPBack.p, par->comp_bias_corr, dBias.p);
    ctx->launches += 1;
    SD_CUDA(cudaGetLastError());
    if ((rc = launchAlignCompositionBias(ctx, dBias.p, nPos, dCb.p))) return rc;
    if ((rc = launchSwProfiles(ctx, dCode.p, dCb.p, dOff.p, nq, nPos, dMat8.p, dProfOff.p, dStride.p, dProfF.p, dProfR.p))) return rc;
    std::vector<int8_t> hCb(total);
    const std::vector<uint8_t> &hCode = queries->hostCode;
    SD_CUDA(cudaMemcpyAsync(hCb.data(), dCb.p, total, cudaMemcpyDeviceToHost, st));
    SD_CUDA(cudaStreamSynchronize(st));
    // 8-bit bias per query (ssw_init :1283-1291): |min matrix entry| + |min composition bias|
    std::vector<int> limit8(nq);
    for (uint32_t q = 0; q < nq; q++) {
        int mn = 0;
        for (uint32_t i = hOff[q]; i < hOff[q + 1]; i++) mn = std::min(mn, (int)hCb[i]);
        limit8[q] = 255 - (std::abs(minMat) + std::abs(mn));
    }

    // ---- pair states, length pre-filter (Alignment.cpp:370), identity shortcut
    const std::vector<uint32_t> &tOff = idx->hostOff;
    const Evaluer evaluer((double)(db_residues ? db_residues : idx->totalRes));
    std::vector<PairState> ps;
    ps.reserve(n_pairs);
    uint64_t nAligned = 0;
    for (uint64_t k = 0; k < n_pairs; k++) {
        const uint32_t q = pair_q[k], t = pair_t[k];
        const float ql = (float)hLen[q], tl = (float)(tOff[t + 1] - tOff[t]);
        if (!canBeCovered(par->cov_thr, par->cov_mode, ql, tl)) continue;
        PairState s;
        s.q = q; s.t = t;
        s.identity = par->same_db && (query_first + q == t);
        ps.push_back(s);
        nAligned++;
    }
    const uint32_t nP = (uint32_t)ps.size();

    SwBatch batch;
    batch.profFwd = dProfF.p; batch.profRev = dProfR.p; batch.profOff = dProfOff.p; batch.profStride = dStride.p; batch.qLen = dLen.p;
    batch.tRes = idx->raw.p; batch.tOff = idx->off.p; batch.gapOpen = par->gap_open; batch.gapExtend = par->gap_extend;

    const uint32_t swWarps = (uint32_t)ctx->smCount * 4u * swWarpsPerCta();
    const uint32_t bndStride = std::max<uint32_t>(idx->maxLen, 32u) + 32u;
    DevBuf<int> dBoundary;
    if ((rc = dBoundary.alloc((size_t)swWarps * 6 * bndStride))) return rc;

    auto runSw = [&](std::vector<SwTask> &tasks, std::vector<SwResult> &results, unsigned long long *cellSlot, int packed16) -> int {
        results.resize(tasks.size());
        if (tasks.empty()) return SD_OK;
        DevBuf<SwTask> dTasks;
        DevBuf<SwResult> dRes;
        DevBuf<uint32_t> dOrd;
        int r2;
        if ((r2 = dTasks.alloc(tasks.size())) || (r2 = dRes.alloc(tasks.size())) || (r2 = dOrd.alloc(tasks.size()))) return r2;
        // group tasks by rows-per-stage class, largest class first: warps that share an SM then run the same
        // template instance (instruction-cache locality) and the long tasks start early
        std::vector<uint32_t> ord(tasks.size());
        {
            uint32_t cnt[18] = {0};
            auto cls = [](const SwTask &t) { const int rp = (t.nRows + 63) / 64; return (uint32_t)(rp > 16 ? 17 : rp); };
            for (const SwTask &t : tasks) cnt[17 - cls(t)]++;
            uint32_t start[18], acc = 0;
            for (int c = 0; c < 18; c++) { start[c] = acc; acc += cnt[c]; }
            for (uint32_t i = 0; i < tasks.size(); i++) ord[start[17 - cls(tasks[i])]++] = i;
        }
        SD_CUDA(cudaMemcpyAsync(dOrd.p, ord.data(), ord.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        SD_CUDA(cudaMemcpyAsync(dTasks.p, tasks.data(), tasks.size() * sizeof(SwTask), cudaMemcpyHostToDevice, st));
        SD_CUDA(cudaEventRecord(evA, st));
        if ((r2 = launchSw(ctx, batch, dTasks.p, dOrd.p, dRes.p, (uint32_t)tasks.size(), dBoundary.p, bndStride, swWarps, dNext.p, cellSlot, packed16))) return r2;
        SD_CUDA(cudaEventRecord(evB, st));
        SD_CUDA(cudaMemcpyAsync(results.data(), dRes.p, tasks.size() * sizeof(SwResult), cudaMemcpyDeviceToHost, st));
        SD_CUDA(cudaStreamSynchronize(st));
        return lap(swMs);
    };

    // ---- forward pass, 8-bit semantics
    std::vector<uint32_t> taskPair;
    std::vector<SwTask> tasks;
    std::vector<SwResult> results;
    tasks.reserve(nP); taskPair.reserve(nP);
    for (uint32_t k = 0; k < nP; k++) {
        if (ps[k].identity) continue;
        SwTask tk;
        tk.q = ps[k].q; tk.t = ps[k].t; tk.nRows = (int)hLen[ps[k].q]; tk.nCols = (int)(tOff[ps[k].t + 1] - tOff[ps[k].t]);
        tk.limit = limit8[ps[k].q]; tk.width = 32; tk.reverse = 0; tk.pad = 0;
        tasks.push_back(tk); taskPair.push_back(k);
    }
    if ((rc = runSw(tasks, results, dCells.p, 1))) return rc;
    // ---- forward pass again in 16-bit semantics where the 8-bit pass saturated
    std::vector<SwTask> wtasks;
    std::vector<uint32_t> wPair;
    for (size_t i = 0; i < tasks.size(); i++) {
        PairState &s = ps[taskPair[i]];
        if (results[i].overflow) {
            SwTask tk = tasks[i];
            tk.limit = INT_MAX; tk.width = 16;
            wtasks.push_back(tk); wPair.push_back(taskPair[i]);
            s.word = 1;
        } else { s.score = results[i].score; s.tEnd = results[i].col; s.qEnd = results[i].row; }
    }
    std::vector<SwResult> wres;
    if ((rc = runSw(wtasks, wres, dCells.p, 1))) return rc;
    {   // scores close to 2^15 leave the packed 16-bit lanes: redo those pairs with the 32-bit pass
        std::vector<SwTask> big;
        std::vector<size_t> bigIdx;
        for (size_t i = 0; i < wtasks.size(); i++)
            if (wres[i].overflow) { big.push_back(wtasks[i]); bigIdx.push_back(i); }
        std::vector<SwResult> bres;
        if ((rc = runSw(big, bres, dCells.p, 0))) return rc;
        for (size_t i = 0; i < big.size(); i++) wres[bigIdx[i]] = bres[i];
    }
    for (size_t i = 0; i < wtasks.size(); i++) {
        PairState &s = ps[wPair[i]];
        s.score = wres[i].score; s.tEnd = wres[i].col; s.qEnd = wres[i].row;
    }
    const uint64_t nWord = wtasks.size();

    // ---- E-value and end-point coverage (ssw_align_private :383-398), identity shortcut (scoreIdentical)
    parallelFor(nP, [&](uint64_t k) {
        PairState &s = ps[k];
        const int Lq = (int)hLen[s.q], Lt = (int)(tOff[s.t + 1] - tOff[s.t]);
        if (s.identity) {
            short sc = 0;
            const uint8_t *c = hCode.data() + hOff[s.q];
            const int8_t *b = hCb.data() + hOff[s.q];
            for (int i = 0; i < Lq && i < Lt; i++) sc = (short)(sc + mat8[c[i] * AA + c[i]] + b[i]);
            s.score = sc;
            s.evalue = evaluer.evalue(s.score, Lq);
            s.qStart = 0; s.tStart = 0; s.qEnd = Lt - 1; s.tEnd = Lt - 1;
            s.qcov = 1.0f; s.tcov = 1.0f;
            s.btLen = Lt; s.ids = Lt;
            s.stage = 3;
            return;
        }
        if (s.tEnd < 0) { s.stage = 0; return; }
        s.evalue = evaluer.evalue(s.score, Lq);
        s.qcov = covOf(0, (unsigned)s.qEnd, (unsigned)Lq);
        s.tcov = covOf(0, (unsigned)s.tEnd, (unsigned)Lt);
        if (s.evalue > par->evalue_thr || !hasCoverage(par->cov_thr, par->cov_mode, s.qcov, s.tcov)) { s.stage = 0; return; }
        s.stage = 1;
    });

    // ---- reverse pass on the prefixes: alignment start (ssw_align_private :400-480)
    std::vector<SwTask> rtasks;
    std::vector<uint32_t> rPair;
    for (uint32_t k = 0; k < nP; k++) {
        if (ps[k].stage != 1) continue;
        SwTask tk;
        tk.q = ps[k].q; tk.t = ps[k].t; tk.nRows = ps[k].qEnd + 1; tk.nCols = ps[k].tEnd + 1;
        tk.limit = ps[k].word ? INT_MAX : limit8[ps[k].q]; tk.width = ps[k].word ? 16 : 32; tk.reverse = 1; tk.pad = 0;
        rtasks.push_back(tk); rPair.push_back(k);
    }
    std::vector<SwResult> rres;
    if ((rc = runSw(rtasks, rres, dCells.p + 1, 1))) return rc;
    {
        std::vector<SwTask> big;
        std::vector<size_t> bigIdx;
        for (size_t i = 0; i < rtasks.size(); i++)
            if (rres[i].overflow == 2) { big.push_back(rtasks[i]); bigIdx.push_back(i); }
        std::vector<SwResult> bres;
        if ((rc = runSw(big, bres, dCells.p + 1, 0))) return rc;
        for (size_t i = 0; i < big.size(); i++) rres[bigIdx[i]] = bres[i];
    }
    uint64_t scoreMismatch = 0;
    for (size_t i = 0; i < rtasks.size(); i++) {
        PairState &s = ps[rPair[i]];
        if (rres[i].overflow || rres[i].score != s.score) { scoreMismatch++; s.stage = 0; continue; }
        const int Lq = (int)hLen[s.q], Lt = (int)(tOff[s.t + 1] - tOff[s.t]);
        s.tStart = s.tEnd - rres[i].col;
        s.qStart = s.qEnd - rres[i].row;
        s.qcov = covOf((unsigned)s.qStart, (unsigned)s.qEnd, (unsigned)Lq);
        s.tcov = covOf((unsigned)s.tStart, (unsigned)s.tEnd, (unsigned)Lt);
        s.stage = hasCoverage(par->cov_thr, par->cov_mode, s.qcov, s.tcov) ? 2 : 0;
    }
    if (scoreMismatch) {
        // the reference aborts in this situation ("Score of forward/backward SW differ")
        setError("sd_align_run: forward and reverse Smith-Waterman scores differ for %llu pairs", (unsigned long long)scoreMismatch);
        return SD_ELIMIT;
    }

    // ---- banded traceback
    std::vector<TbTask> ttasks;
    std::vector<uint32_t> tPair;
    uint64_t poolNeed = 0;
    for (uint32_t k = 0; k < nP; k++) {
        PairState &s = ps[k];
        if (s.stage == 3) { s.btOff = poolNeed; poolNeed += (uint64_t)s.btLen; continue; }
        if (s.stage != 2) continue;
        TbTask tk;
        tk.q = s.q; tk.t = s.t; tk.qStart = s.qStart; tk.qEnd = s.qEnd; tk.tStart = s.tStart; tk.tEnd = s.tEnd; tk.score = s.score; tk.pad = 0;
        tk.btOff = poolNeed;
        s.btOff = poolNeed;
        poolNeed += (uint64_t)(s.qEnd - s.qStart + 1) + (uint64_t)(s.tEnd - s.tStart + 1) + 2;
        ttasks.push_back(tk); tPair.push_back(k);
    }
    std::vector<char> hostPool(poolNeed + 1);
    std::vector<TbResult> tres(ttasks.size());
    if (!ttasks.empty()) {
        DevBuf<TbTask> dT;
        DevBuf<TbResult> dR;
        DevBuf<char> dPool, dRev;
        DevBuf<int8_t> dDir;
        DevBuf<int> dRows;
        DevBuf<uint32_t> dOrder;
        if ((rc = dT.alloc(ttasks.size())) || (rc = dR.alloc(ttasks.size())) || (rc = dPool.alloc(poolNeed + 1))) return rc;
        SD_CUDA(cudaMemcpyAsync(dT.p, ttasks.data(), ttasks.size() * sizeof(TbTask), cudaMemcpyHostToDevice, st));
        uint32_t maxQ = 1;
        for (uint32_t q = 0; q < nq; q++) maxQ = std::max(maxQ, hLen[q]);
        const uint32_t maxL = std::max(maxQ, idx->maxLen);
        TbArgs A;
        A.qCode = dCode.p; A.qCb = dCb.p; A.qOff = dOff.p; A.tRes = idx->raw.p; A.tOff = idx->off.p; A.mat = dMat8.p;
        A.gapOpen = par->gap_open; A.gapExtend = par->gap_extend; A.tasks = dT.p; A.results = dR.p; A.next = dNext.p;
        A.btPool = dPool.p; A.cells = dCells.p + 2;
        // pass 1: every task with a modest scratch slot; pass 2: the few that need a wide band
        std::vector<uint32_t> pending;
        for (int pass = 0; pass < 2; pass++) {
            uint32_t nTasks = pass == 0 ? (uint32_t)ttasks.size() : (uint32_t)pending.size();
            if (nTasks == 0) break;
            uint32_t nWarps;
            uint64_t dirCap;
            uint32_t rowCap;
            if (pass == 0) {
                nWarps = std::min<uint32_t>((uint32_t)ctx->smCount * 4u * TB_WARPS, (nTasks + 0u));
                dirCap = 1ull << 20;
                rowCap = 4096;
            } else {
                nWarps = std::min<uint32_t>(64u, nTasks);
                const uint64_t worstBand = 4ull * maxL + 8ull;
                dirCap = std::min<uint64_t>((2 * worstBand + 1) * (uint64_t)maxL * 3ull, 1ull << 30);
                rowCap = (uint32_t)std::min<uint64_t>(2 * worstBand + 8, 1u << 20);
            }
            nWarps = ((nWarps + TB_WARPS - 1) / TB_WARPS) * TB_WARPS;
            if ((rc = dDir.alloc((size_t)nWarps * dirCap)) || (rc = dRows.alloc((size_t)nWarps * 3 * rowCap)) || (rc = dRev.alloc((size_t)nWarps * (2 * maxL + 8)))) return rc;
            A.dir = dDir.p; A.dirCap = dirCap; A.rows = dRows.p; A.rowCap = rowCap; A.rev = dRev.p; A.revCap = 2 * maxL + 8;
            A.nTasks = nTasks;
            if (pass == 0) A.order = nullptr;
            else {
                if ((rc = dOrder.alloc(pending.size()))) return rc;
                SD_CUDA(cudaMemcpyAsync(dOrder.p, pending.data(), pending.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
                A.order = dOrder.p;
            }
            SD_CUDA(cudaMemsetAsync(dNext.p, 0, sizeof(uint32_t), st));
            SD_CUDA(cudaEventRecord(evA, st));
            tracebackKernel<<<nWarps / TB_WARPS, TB_WARPS * 32, 0, st>>>(A);
            ctx->launches++;
            SD_CUDA(cudaGetLastError());
            SD_CUDA(cudaEventRecord(evB, st));
            SD_CUDA(cudaMemcpyAsync(tres.data(), dR.p, ttasks.size() * sizeof(TbResult), cudaMemcpyDeviceToHost, st));
            SD_CUDA(cudaStreamSynchronize(st));
            if ((rc = lap(tbMs))) return rc;
            pending.clear();
            for (uint32_t i = 0; i < ttasks.size(); i++)
                if (tres[i].status == 1) pending.push_back(i);
            if (pass == 1 && !pending.empty()) { setError("sd_align_run: %zu tracebacks exceed the scratch limit", pending.size()); return SD_ELIMIT; }
        }
        SD_CUDA(cudaMemcpyAsync(hostPool.data(), dPool.p, poolNeed, cudaMemcpyDeviceToHost, st));
        SD_CUDA(cudaStreamSynchronize(st));
        uint64_t tbErr = 0;
        for (uint32_t i = 0; i < ttasks.size(); i++) {
            PairState &s = ps[tPair[i]];
            if (tres[i].status != 0) { tbErr++; s.stage = 0; continue; }
            s.btLen = tres[i].btLen; s.ids = tres[i].ids; s.stage = 3;
        }
        if (tbErr) { setError("sd_align_run: %llu tracebacks left the band (inconsistent start/end)", (unsigned long long)tbErr); return SD_ELIMIT; }
    }
    for (uint32_t k = 0; k < nP; k++)
        if (ps[k].identity && ps[k].stage == 3) memset(hostPool.data() + ps[k].btOff, 'M', (size_t)ps[k].btLen);

    // ---- acceptance (Alignment::checkCriteria), order (Matcher::compareHits), output
    std::vector<sd_alignment> acc;
    acc.reserve(nP / 4 + 16);
    for (uint32_t k = 0; k < nP; k++) {
        const PairState &s = ps[k];
        if (s.stage != 3) continue;
        sd_alignment a;
        memset(&a, 0, sizeof a);
        a.query = s.q; a.target = s.t; a.score = s.score; a.evalue = s.evalue;
        a.q_start = s.qStart; a.q_end = s.qEnd; a.q_len = (int)hLen[s.q];
        a.t_start = s.tStart; a.t_end = s.tEnd; a.t_len = (int)(tOff[s.t + 1] - tOff[s.t]);
        a.qcov = s.qcov; a.tcov = s.tcov;
        a.aln_len = s.btLen;
        a.seq_id = s.identity ? 1.0f : (float)s.ids / (float)s.btLen;
        a.bt_off = s.btOff; a.bt_len = (uint32_t)s.btLen;
        const bool ok = s.identity || ((a.evalue <= par->evalue_thr) && (a.seq_id >= par->min_seq_id) &&
                                       hasCoverage(par->cov_thr, par->cov_mode, a.qcov, a.tcov) && (a.aln_len >= par->min_aln_len));
        if (!ok) continue;
        a.bits = evaluer.bitScore(a.score);
        acc.push_back(a);
    }
    // pairs are grouped by query already; sort each group
    {
        size_t i = 0;
        while (i < acc.size()) {
            size_t j = i;
            while (j < acc.size() && acc[j].query == acc[i].query) j++;
            std::sort(acc.begin() + i, acc.begin() + j, [](const sd_alignment &x, const sd_alignment &y) {
                if (x.evalue != y.evalue) return x.evalue < y.evalue;
                if (x.bits != y.bits) return x.bits > y.bits;
                if (x.t_len != y.t_len) return x.t_len < y.t_len;
                return x.target < y.target;      // keys ascend with the target index
            });
            i = j;
        }
    }
    // compact the backtraces of accepted alignments into the caller's pool
    uint64_t used = 0;
    for (const sd_alignment &a : acc) used += a.bt_len;
    *n_out = acc.size();
    *bt_used = used;
    if (acc.size() > cap || used > bt_cap || (acc.size() && (!out || !bt_pool))) {
        setError("sd_align_run: needs %zu alignments / %llu backtrace bytes", acc.size(), (unsigned long long)used);
        return SD_ESMALL;
    }
    uint64_t w = 0;
    for (size_t i = 0; i < acc.size(); i++) {
        memcpy(bt_pool + w, hostPool.data() + acc[i].bt_off, acc[i].bt_len);
        acc[i].bt_off = w;
        w += acc[i].bt_len;
        out[i] = acc[i];
    }
    SD_CUDA(cudaEventRecord(ev1, st));
    SD_CUDA(cudaStreamSynchronize(st));
    if (stats) {
        unsigned long long hc[4];
        SD_CUDA(cudaMemcpy(hc, dCells.p, sizeof hc, cudaMemcpyDeviceToHost));
        float ms = 0;
        cudaEventElapsedTime(&ms, ev0, ev1);
        stats->alignments = nAligned; stats->passed = acc.size(); stats->word_mode = nWord;
        stats->reverse_passes = rtasks.size(); stats->tracebacks = ttasks.size();
        stats->cells_forward = hc[0]; stats->cells_reverse = hc[1]; stats->cells_traceback = hc[2];
        stats->sw_ms = swMs; stats->traceback_ms = tbMs; stats->total_ms = ms;
    }
    cudaEventDestroy(ev0); cudaEventDestroy(ev1); cudaEventDestroy(evA); cudaEventDestroy(evB);
    return SD_OK;
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
    // Matcher::compressAlignment (Matcher.cpp:166-185)
    const char *bt = bt_pool + a->bt_off;
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
