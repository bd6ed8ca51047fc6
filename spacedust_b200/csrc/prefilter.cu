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
        uint32_t half[2] = {0, 0}, mul = 1;
        int self = 0;
        bool ok = true;
#pragma unroll
        for (int k = 0; k < KMER; k++) {
            uint8_t c = res[gp + c_seedPos[k]];
            ok = ok && (c < 20);
            self += selfScore[c];
            half[k / 3] += (uint32_t)c * mul;
            mul = (k == 2) ? 1 : mul * 20;
        }
        const uint32_t idx = kmerSlot(half[0], half[1]);
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

// sorted neighbour rows as one word per element: score (low half) | 3-mer index (high half)
__global__ void pairExt3(const int16_t *__restrict__ score, const uint16_t *__restrict__ index, uint64_t n, uint32_t *__restrict__ pair) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) pair[i] = (uint32_t)(uint16_t)score[i] | ((uint32_t)index[i] << 16);
}

// {bits, rank} side by side: one 8-byte load answers a probe (empty or not) and gives the row of a non-empty k-mer
__global__ void interleaveBitmap(const uint32_t *__restrict__ bits, const uint32_t *__restrict__ rank, uint32_t nWords,
                                 uint2 *__restrict__ bm) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nWords) bm[i] = make_uint2(bits[i], rank[i]);
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

#include "prefilter_kernel.cuh"
#include "prefilter_stream.cuh"

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
    DevBuf<int16_t> dScore;
    DevBuf<uint16_t> dIndex;
    if ((rc = dScore.alloc(sc.size())) != SD_OK) return rc;
    if ((rc = dIndex.alloc(ix.size())) != SD_OK) return rc;
    if ((rc = ctx->e3pair.alloc(sc.size())) != SD_OK) return rc;
    if ((rc = ctx->e3cnt.alloc(cn.size())) != SD_OK) return rc;
    SD_CUDA(cudaMemcpyAsync(ctx->e3cnt.p, cn.data(), cn.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, ctx->stream));
    SD_CUDA(cudaMemcpyAsync(dScore.p, sc.data(), sc.size() * sizeof(int16_t), cudaMemcpyHostToDevice, ctx->stream));
    SD_CUDA(cudaMemcpyAsync(dIndex.p, ix.data(), ix.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, ctx->stream));
    pairExt3<<<gridFor(sc.size(), 256), 256, 0, ctx->stream>>>(dScore.p, dIndex.p, sc.size(), ctx->e3pair.p);
    ctx->launches++;
    SD_CUDA(cudaGetLastError());
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
    std::vector<uint8_t> codes(total + 32);
    mapSequence(seqs, total, codes.data());
    std::fill(codes.begin() + total, codes.end(), (uint8_t)XCODE);
    int rc;
    cudaStream_t st = ctx->stream;
    if ((rc = idx->raw.alloc(total + 32)) != SD_OK) return rc;
    if ((rc = idx->res.alloc(total + 32)) != SD_OK) return rc;
    if ((rc = idx->off.alloc(n + 1)) != SD_OK) return rc;
    SD_CUDA(cudaMemcpyAsync(idx->raw.p, codes.data(), total + 32, cudaMemcpyHostToDevice, st));
    SD_CUDA(cudaMemcpyAsync(idx->res.p, idx->raw.p, total + 32, cudaMemcpyDeviceToDevice, st));
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
        DevBuf<uint32_t> cnt, bits, rank;
        if ((rc = bits.alloc((size_t)nWords + 1)) || (rc = rank.alloc((size_t)nWords + 1)) || (rc = cnt.alloc(nWords + 1)) || (rc = idx->bm.alloc((size_t)nWords + 1))) return rc;
        SD_CUDA(cudaMemsetAsync(cnt.p + nWords, 0, sizeof(uint32_t), st));
        occupancyBits<<<gridFor((uint64_t)((nWords + 31) / 32) * 32, 256), 256, 0, st>>>(kOff.p, nWords, bits.p, cnt.p);
        size_t tmpBytes = 0;
        SD_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, cnt.p, rank.p, (int64_t)nWords + 1, st));
        DevBuf<uint8_t> tmp;
        if ((rc = tmp.alloc(tmpBytes)) != SD_OK) return rc;
        SD_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmpBytes, cnt.p, rank.p, (int64_t)nWords + 1, st));
        uint32_t nDistinct = 0;
        SD_CUDA(cudaMemcpyAsync(&nDistinct, rank.p + nWords, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        SD_CUDA(cudaStreamSynchronize(st));
        idx->nDistinct = nDistinct;
        if ((rc = idx->rows.alloc((size_t)nDistinct + 1)) != SD_OK) return rc;
        compactRows<<<gridFor((uint64_t)nWords * 32, 256), 256, 0, st>>>(kOff.p, bits.p, rank.p, nWords, idx->ent.p, idx->rows.p);
        interleaveBitmap<<<gridFor(nWords, 256), 256, 0, st>>>(bits.p, rank.p, nWords, idx->bm.p);
        ctx->launches += 3;
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
    sd::EventQuad events;
    { const int erc = events.create(); if (erc) return erc; }
    cudaEvent_t ev0 = events.e[0], ev1 = events.e[1], evk0 = events.e[2], evk1 = events.e[3];
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
    uint32_t maxLq = 1;
    for (uint32_t i = 0; i < nq; i++) maxLq = std::max(maxLq, hOff[i + 1] - hOff[i]);

    // QueryMatcher.cpp:43 (SD_PF_MAXDBMATCHES: test hook, the oracle has the same knob)
    unsigned long long maxDbMatches = 2ull * std::max<unsigned long long>(1000000ull, idx->nT);
    if (getenv("SD_PF_MAXDBMATCHES")) maxDbMatches = strtoull(getenv("SD_PF_MAXDBMATCHES"), nullptr, 10);

    ScratchBuf<sd_prefilter_hit> dOut(ctx, SL_PF_OUT);
    DevBuf<uint32_t> dOutCount, dNext, dFail, dOrder2;
    DevBuf<unsigned long long> dStats;
    if ((rc = dOut.alloc((size_t)nq * maxHits)) || (rc = dOutCount.alloc(nq)) || (rc = dNext.alloc(4)) || (rc = dStats.alloc(ST_N)) || (rc = dFail.alloc(nq)) ||
        (rc = dOrder2.alloc(nq)))
        return rc;
    SD_CUDA(cudaMemsetAsync(dOutCount.p, 0, nq * sizeof(uint32_t), st));
    SD_CUDA(cudaMemsetAsync(dStats.p, 0, ST_N * sizeof(unsigned long long), st));
    const uint32_t initNext[4] = {0u, 0xffffffffu, 0u, 0u};     // next query, first failing query, number of listed queries
    SD_CUDA(cudaMemcpyAsync(dNext.p, initNext, sizeof initNext, cudaMemcpyHostToDevice, st));
    SD_CUDA(cudaEventRecord(evk0, st));

    // ---- kernel choice by the number of targets: the streaming kernel (one warp per query, per-warp last-diagonal table in
    // global memory) has the lower cost per index entry while its tables stay small; the binned kernel (one CTA per query,
    // shared-memory tables) does not depend on the table size and restates the reference's hit-buffer overflow, so it also
    // takes the queries the streaming kernel hands over.  SD_PF_STREAM_MAX: largest target count for the streaming kernel.
    const long streamMax = getenv("SD_PF_STREAM_MAX") ? atol(getenv("SD_PF_STREAM_MAX")) : 262144;
    const bool useStream = (long)idx->nT <= streamMax;
    const uint32_t *binnedOrder = dOrder.p;      // queries for the binned kernel (device list) ...
    uint32_t binnedCount = nq;                   // ... and how many
    if (useStream) {
        static const int ctasPerSmS = getenv("SD_PS_CTAS") ? std::max(1, std::min(4, atoi(getenv("SD_PS_CTAS")))) : 4;      // tuning hook
        const uint32_t nCtaS = std::max(1u, std::min<uint32_t>((uint32_t)(ctx->smCount * ctasPerSmS), (nq + PS_WARPS - 1) / PS_WARPS));
        const uint32_t nWarpsS = nCtaS * PS_WARPS;
        // A warp works through its query alone: queries much longer than a warp's share of the batch set the kernel's tail (strong
        // scaling of small batches).  They go to the binned kernel instead, where a whole CTA shares one query.
        uint32_t nLong = 0;
        {
            const double share = (double)total / (double)nWarpsS;
            const double longThr = std::max(512.0, share);
            while (nLong < nq && (double)(hOff[order[nLong] + 1] - hOff[order[nLong]]) > longThr) nLong++;
            if (getenv("SD_PF_NOLONG")) nLong = 0;
        }
        // first-pass capacity of the per-warp candidate lists (SD_PF_CANDCAP: test hook that forces the second pass)
        uint32_t candCapS = 32768u;
        while (candCapS < 262144u && (uint64_t)candCapS * 8u < idx->nT) candCapS *= 2;
        if (getenv("SD_PF_CANDCAP")) candCapS = (uint32_t)std::max(32, atoi(getenv("SD_PF_CANDCAP")));
        ScratchBuf<uint32_t> sTable(ctx, SL_PF_TABLE), sResT(ctx, SL_PF_REST);
        ScratchBuf<uint64_t> sCand(ctx, SL_PF_CAND), sKeep(ctx, SL_PF_KEEP);
        ScratchBuf<int32_t> sResScore(ctx, SL_PF_RESS);
        ScratchBuf<uint16_t> sResDiag(ctx, SL_PF_RESD);
        if ((rc = sTable.alloc((size_t)nWarpsS * idx->nT)) || (rc = sCand.alloc((size_t)nWarpsS * candCapS)) || (rc = sKeep.alloc((size_t)nWarpsS * candCapS)) ||
            (rc = sResT.alloc((size_t)nWarpsS * maxHits)) || (rc = sResScore.alloc((size_t)nWarpsS * maxHits)) || (rc = sResDiag.alloc((size_t)nWarpsS * maxHits)))
            return rc;
        SD_CUDA(cudaMemsetAsync(sTable.p, 0, (size_t)nWarpsS * idx->nT * sizeof(uint32_t), st));
        StreamArgs sa{};
        sa.tRes = idx->res.p; sa.tOff = idx->off.p; sa.nT = idx->nT;
        sa.bm = idx->bm.p; sa.rows = idx->rows.p; sa.ent = idx->ent.p;
        sa.e3pair = ctx->e3pair.p; sa.e3cnt = ctx->e3cnt.p; sa.diagSub = dDiagSub.p;
        sa.q.code = dCode.p; sa.q.off = dOff.p; sa.q.corr = dCorr.p; sa.q.thr = dThr.p;
        sa.nQ = nq - nLong; sa.order = dOrder.p + nLong; sa.queryFirst = query_first; sa.sameDb = par->same_db;
        sa.maxHits = maxHits; sa.minDiagScore = (uint32_t)std::max(par->min_ungapped_score, 0);
        sa.covThr = par->cov_thr; sa.covMode = par->cov_mode; sa.binMask = (uint32_t)binSize - 1u;
        sa.maxDbMatches = maxDbMatches;
        sa.table = sTable.p; sa.candCap = candCapS; sa.cand = sCand.p; sa.keep = sKeep.p;
        sa.resT = sResT.p; sa.resScore = sResScore.p; sa.resDiag = sResDiag.p;
        sa.out = dOut.p; sa.outCount = dOutCount.p; sa.stats = dStats.p; sa.nextQuery = dNext.p; sa.errQuery = dNext.p + 1;
        sa.failList = dFail.p; sa.nFail = dNext.p + 2;
        // profiling switches (results with SD_PF_SKIP are wrong): L2 policy of the streamed data, prefetches, dropped stages
        sa.evictFirst = getenv("SD_PS_EVICT") ? atoi(getenv("SD_PS_EVICT")) : 0;
        sa.prefetch = getenv("SD_PS_PREFETCH") ? atoi(getenv("SD_PS_PREFETCH")) : 0;
        sa.debugSkip = getenv("SD_PF_SKIP") ? atoi(getenv("SD_PF_SKIP")) : 0;
        const int persistS = getenv("SD_PS_PERSIST") ? atoi(getenv("SD_PS_PERSIST")) : 0;
        if (persistS) {
            int maxPersist = 0, maxWin = 0;
            cudaDeviceGetAttribute(&maxPersist, cudaDevAttrMaxPersistingL2CacheSize, ctx->device);
            cudaDeviceGetAttribute(&maxWin, cudaDevAttrMaxAccessPolicyWindowSize, ctx->device);
            const size_t bmBytes = idx->bm.n * sizeof(uint2);
            if (getenv("SD_PF_DEBUG")) fprintf(stderr, "[sd] persisting L2: max %d bytes, window max %d bytes, bitmap %zu bytes\n", maxPersist, maxWin, bmBytes);
            if (maxPersist > 0 && maxWin > 0) {
                cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min<size_t>((size_t)maxPersist, (size_t)persistS << 20));
                cudaStreamAttrValue av{};
                av.accessPolicyWindow.base_ptr = (void *)idx->bm.p;
                av.accessPolicyWindow.num_bytes = std::min<size_t>(bmBytes, (size_t)maxWin);
                av.accessPolicyWindow.hitRatio = 1.0f;
                av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                av.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
                cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av);
            }
        }
        SD_CUDA(cudaFuncSetAttribute(prefilterStreamKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PS_DYN_SMEM));
        if (getenv("SD_PF_DEBUG")) fprintf(stderr, "[sd] prefilter: streaming kernel, %u CTAs x %d warps, candidate lists %u\n", nCtaS, PS_WARPS, candCapS);
        if (sa.nQ) {
            prefilterStreamKernel<<<nCtaS, PS_WARPS * 32, PS_DYN_SMEM, st>>>(sa);
            ctx->launches++;
            SD_CUDA(cudaGetLastError());
        }
        // listed queries: bit 31 = reached the reference's hit-buffer limit (-> binned kernel); otherwise the candidate list
        // did not fit (-> second streaming launch over just those, few warps with 4 M entry lists)
        uint32_t hN[4];
        SD_CUDA(cudaMemcpyAsync(hN, dNext.p, sizeof hN, cudaMemcpyDeviceToHost, st));
        SD_CUDA(cudaStreamSynchronize(st));
        std::vector<uint32_t> listed(hN[2]), big, over;
        if (hN[2]) {
            SD_CUDA(cudaMemcpyAsync(listed.data(), dFail.p, hN[2] * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            SD_CUDA(cudaStreamSynchronize(st));
        }
        for (uint32_t i = 0; i < nLong; i++) over.push_back(order[i]);
        for (uint32_t v : listed) (v >> 31 ? over : big).push_back(v & 0x7fffffffu);
        if (getenv("SD_PF_DEBUG"))
            fprintf(stderr, "[sd] prefilter: %u long queries to the binned kernel, %zu queries reach the hit-buffer limit, %zu overflow the candidate lists\n",
                    nLong, over.size() - nLong, big.size());
        if (!big.empty()) {
            const uint32_t bigCap = 1u << 22;
            const uint32_t nCta2 = std::max(1u, std::min<uint32_t>(((uint32_t)big.size() + PS_WARPS - 1) / PS_WARPS, 16u));
            const uint32_t nWarps2 = nCta2 * PS_WARPS;
            DevBuf<uint64_t> dCand2, dKeep2;
            if ((rc = dCand2.alloc((size_t)nWarps2 * bigCap)) || (rc = dKeep2.alloc((size_t)nWarps2 * bigCap))) return rc;
            SD_CUDA(cudaMemcpyAsync(dOrder2.p, big.data(), big.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
            SD_CUDA(cudaMemsetAsync(sTable.p, 0, (size_t)nWarps2 * idx->nT * sizeof(uint32_t), st));      // epochs restart at 0
            SD_CUDA(cudaMemsetAsync(dStats.p + ST_ERR_CAP, 0, sizeof(unsigned long long), st));
            SD_CUDA(cudaMemcpyAsync(dNext.p, initNext, sizeof initNext, cudaMemcpyHostToDevice, st));
            StreamArgs sb = sa;
            sb.order = dOrder2.p; sb.nQ = (uint32_t)big.size(); sb.candCap = bigCap; sb.cand = dCand2.p; sb.keep = dKeep2.p;
            prefilterStreamKernel<<<nCta2, PS_WARPS * 32, PS_DYN_SMEM, st>>>(sb);
            ctx->launches++;
            SD_CUDA(cudaGetLastError());
            SD_CUDA(cudaMemcpyAsync(hN, dNext.p, sizeof hN, cudaMemcpyDeviceToHost, st));
            SD_CUDA(cudaStreamSynchronize(st));     // the second-pass buffers go out of scope
            if (hN[2]) {     // still listed: hit-buffer limit reached (cannot happen: same stream) or 4 M candidates exceeded
                std::vector<uint32_t> l2(hN[2]);
                SD_CUDA(cudaMemcpyAsync(l2.data(), dFail.p, hN[2] * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
                SD_CUDA(cudaStreamSynchronize(st));
                for (uint32_t v : l2) over.push_back(v & 0x7fffffffu);     // the binned kernel's scratch is sized per query length
            }
        }
        binnedCount = (uint32_t)over.size();
        if (binnedCount) {
            SD_CUDA(cudaMemcpyAsync(dOrder2.p, over.data(), over.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
            SD_CUDA(cudaMemsetAsync(dStats.p + ST_ERR_CAP, 0, sizeof(unsigned long long), st));
            SD_CUDA(cudaStreamSynchronize(st));     // `over` goes out of scope
            binnedOrder = dOrder2.p;
        }
    }

  if (binnedCount) {
    SD_CUDA(cudaMemcpyAsync(dNext.p, initNext, sizeof initNext, cudaMemcpyHostToDevice, st));
    // ---- launch geometry: two CTAs per SM when the shared-memory split allows it
    cudaFuncAttributes fa{};
    SD_CUDA(cudaFuncGetAttributes(&fa, prefilterKernel));
    int smemPerSm = 0, smemOptin = 0;
    SD_CUDA(cudaDeviceGetAttribute(&smemPerSm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, ctx->device));
    SD_CUDA(cudaDeviceGetAttribute(&smemOptin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
    static const int ctasPerSm = getenv("SD_PF_CTAS") ? std::max(1, atoi(getenv("SD_PF_CTAS"))) : 3;
    const size_t dynBudget = std::min<size_t>((size_t)smemOptin, (size_t)smemPerSm / ctasPerSm - 1024) - fa.sharedSizeBytes;
    // ---- target bins: id ranges whose table (2 bytes + 1 bit per target) fits a warp's share of the dynamic shared memory;
    // nBins is a multiple of the warps so that every round of P2 keeps all warps busy
    uint32_t nBins = 0, tBinSize = 0, tabBytes = 0;
    for (uint32_t rounds = 1; rounds <= 4096; rounds++) {
        const uint32_t nb = rounds * PF_WARPS;
        const size_t cntBytes = ((size_t)PF_WARPS * nb * 4 + 15) & ~(size_t)15;
        if (cntBytes + PF_WARPS * 256 > dynBudget) break;
        const size_t perWarp = ((dynBudget - cntBytes) / PF_WARPS) & ~(size_t)15;
        uint32_t bs = (idx->nT + nb - 1) / nb;
        bs = std::max(bs, 512u);
        const size_t need = (((size_t)bs * 2 + 15) & ~(size_t)15) + ((((size_t)bs + 31) / 32 * 4 + 15) & ~(size_t)15);
        if (bs <= PF_MAXBINSIZE && need <= perWarp) { nBins = nb; tBinSize = bs; tabBytes = (uint32_t)need; break; }
    }
    if (!nBins) { setError("prefilter: %u targets do not fit the shared-memory bin tables", idx->nT); return SD_ELIMIT; }
    const size_t dynSmem = (((size_t)PF_WARPS * nBins * 4 + 15) & ~(size_t)15) + (size_t)PF_WARPS * tabBytes;
    SD_CUDA(cudaFuncSetAttribute(prefilterKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dynSmem));
    int occ = 0;
    SD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, prefilterKernel, PF_THREADS, dynSmem));
    if (occ < 1) { setError("prefilter kernel does not fit an SM (%zu bytes of shared memory)", dynSmem + fa.sharedSizeBytes); return SD_ELIMIT; }
    const uint32_t nCta = std::max(1u, std::min<uint32_t>((uint32_t)(ctx->smCount * std::min(occ, ctasPerSm)), binnedCount));

    // ---- per-CTA scratch.  Expected index entries per query position: similar k-mers per position x entries per k-mer
    // (measured on the synthetic proteomes: 425 x nEntries / 20^6); the regions of (slice, bin) get 4x the share of the
    // longest query, queries that still overflow are re-run by a second launch with 16x the space
    const double hitsPerPos = 425.0 * (double)std::max<uint64_t>(idx->nEntries, 1) / (double)NKMER;
    const double mMax = std::max(4096.0, hitsPerPos * (double)maxLq);
    auto roundUp = [](double x, uint32_t m) { return (uint32_t)(((uint64_t)x + m - 1) / m * m); };
    uint32_t regionCap = std::max(256u, roundUp(4.0 * mMax / ((double)PF_WARPS * nBins), 32));
    uint32_t candCap = std::max(2048u, roundUp(mMax / 16.0, 32));
    if (getenv("SD_PF_REGIONCAP")) regionCap = (uint32_t)std::max(32, atoi(getenv("SD_PF_REGIONCAP")));     // test hooks: force the second pass
    if (getenv("SD_PF_CANDCAP")) candCap = (uint32_t)std::max(32, atoi(getenv("SD_PF_CANDCAP")));
    candCap = std::min(candCap, 1u << 23);
    ScratchBuf<uint2> dPool(ctx, SL_PF_TABLE);
    ScratchBuf<uint64_t> dCand(ctx, SL_PF_CAND);
    ScratchBuf<uint32_t> dCandSeq(ctx, SL_PF_CSEQ), dSortIdx(ctx, SL_PF_SORTIDX), dResT(ctx, SL_PF_REST), dResAux(ctx, SL_PF_RESD);
    ScratchBuf<uint4> dKeep(ctx, SL_PF_KEEP);
    ScratchBuf<int32_t> dResScore(ctx, SL_PF_RESS);
    ScratchBuf<uint16_t> dQpack(ctx, SL_PF_QPACK);
    ScratchBuf<uint4> dPosInfo(ctx, SL_PF_POSINFO);
    ScratchBuf<uint32_t> dPosPrefix(ctx, SL_PF_POSPREFIX);
    auto allocScratch = [&](uint32_t ctas, uint32_t rCap, uint32_t cCap, uint32_t &kCap) -> int {
        int r;
        kCap = (uint32_t)std::min<uint64_t>(idx->nT, (uint64_t)PF_WARPS * cCap);
        if ((r = dPool.alloc((size_t)ctas * PF_WARPS * nBins * rCap)) || (r = dCand.alloc((size_t)ctas * PF_WARPS * cCap)) ||
            (r = dCandSeq.alloc((size_t)ctas * PF_WARPS * cCap)) || (r = dSortIdx.alloc((size_t)ctas * PF_WARPS * cCap)) || (r = dKeep.alloc((size_t)ctas * kCap)) ||
            (r = dResT.alloc((size_t)ctas * maxHits)) || (r = dResScore.alloc((size_t)ctas * maxHits)) || (r = dResAux.alloc((size_t)ctas * maxHits)) ||
            (r = dQpack.alloc((size_t)ctas * (maxLq + 16))) || (r = dPosInfo.alloc((size_t)ctas * (maxLq + 16))) ||
            (r = dPosPrefix.alloc((size_t)ctas * (maxLq + 16))))
            return r;
        return SD_OK;
    };
    uint32_t keepCap = 0;
    if ((rc = allocScratch(nCta, regionCap, candCap, keepCap))) return rc;

    PrefilterArgs a{};
    a.tRes = idx->res.p; a.tOff = idx->off.p; a.nT = idx->nT;
    a.bm = idx->bm.p; a.rows = idx->rows.p; a.ent = idx->ent.p;
    a.e3pair = ctx->e3pair.p; a.e3cnt = ctx->e3cnt.p; a.diagSub = dDiagSub.p;
    a.q.code = dCode.p; a.q.off = dOff.p; a.q.corr = dCorr.p; a.q.thr = dThr.p;
    a.nQ = binnedCount; a.order = binnedOrder; a.queryFirst = query_first; a.sameDb = par->same_db;
    a.maxHits = maxHits; a.minDiagScore = (uint32_t)std::max(par->min_ungapped_score, 0);
    a.covThr = par->cov_thr; a.covMode = par->cov_mode; a.binMask = (uint32_t)binSize - 1u;
    a.maxDbMatches = maxDbMatches;
    a.nBins = nBins; a.binSize = tBinSize; a.binMagic = ((1ull << 40) + tBinSize - 1) / tBinSize; a.tabBytes = tabBytes;
    a.regionCap = regionCap; a.pool = dPool.p; a.candCap = candCap; a.cand = dCand.p; a.candSeq = dCandSeq.p; a.sortIdx = dSortIdx.p;
    a.keepCap = keepCap; a.keep = dKeep.p; a.resT = dResT.p; a.resScore = dResScore.p; a.resAux = dResAux.p;
    a.qpackG = dQpack.p; a.posInfo = dPosInfo.p; a.posPrefix = dPosPrefix.p; a.maxLq = maxLq + 16;
    a.out = dOut.p; a.outCount = dOutCount.p; a.stats = dStats.p; a.nextQuery = dNext.p; a.errQuery = dNext.p + 1;
    a.failList = dFail.p; a.nFail = dNext.p + 2;
    a.debugSkip = getenv("SD_PF_SKIP") ? atoi(getenv("SD_PF_SKIP")) : 0;
    if (getenv("SD_PF_DEBUG"))
        fprintf(stderr, "[sd] prefilter: %u CTAs x %d threads, %zu B dynamic + %zu B static smem, %u bins of %u targets, region %u, cand %u, keep %u, maxDbMatches %llu\n",
                nCta, PF_THREADS, dynSmem, (size_t)fa.sharedSizeBytes, nBins, tBinSize, regionCap, candCap, keepCap, maxDbMatches);

    static const int envPersist = getenv("SD_PF_PERSIST") ? atoi(getenv("SD_PF_PERSIST")) : 0;
    if (envPersist) {
        // the occupancy bitmap is probed once per similar k-mer: keep it in the persisting part of the L2
        int maxPersist = 0, maxWin = 0;
        cudaDeviceGetAttribute(&maxPersist, cudaDevAttrMaxPersistingL2CacheSize, ctx->device);
        cudaDeviceGetAttribute(&maxWin, cudaDevAttrMaxAccessPolicyWindowSize, ctx->device);
        const size_t bmBytes = idx->bm.n * sizeof(uint2);
        if (maxPersist > 0 && maxWin > 0) {
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min<size_t>((size_t)maxPersist, bmBytes + (8u << 20)));
            cudaStreamAttrValue av{};
            av.accessPolicyWindow.base_ptr = (void *)idx->bm.p;
            av.accessPolicyWindow.num_bytes = std::min<size_t>(bmBytes, (size_t)maxWin);
            av.accessPolicyWindow.hitRatio = 1.0f;
            av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            av.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
            cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av);
        }
    }
    prefilterKernel<<<nCta, PF_THREADS, dynSmem, st>>>(a);
    ctx->launches++;
    SD_CUDA(cudaGetLastError());
    {   // queries whose hit / candidate lists did not fit the per-CTA scratch: second launch over just those, few CTAs with
        // 16x the space each
        uint32_t hN[4];
        SD_CUDA(cudaMemcpyAsync(hN, dNext.p, sizeof hN, cudaMemcpyDeviceToHost, st));
        SD_CUDA(cudaStreamSynchronize(st));
        const uint32_t nFail = hN[2];
        if (getenv("SD_PF_DEBUG")) fprintf(stderr, "[sd] prefilter: %u queries overflowed the first-pass scratch\n", nFail);
        if (nFail) {
            const uint32_t nCta2 = std::max(1u, std::min<uint32_t>(nFail, std::max(1u, nCta / 16)));
            DevBuf<uint32_t> dOrder3;
            if ((rc = dOrder3.alloc(nFail))) return rc;
            SD_CUDA(cudaMemcpyAsync(dOrder3.p, dFail.p, nFail * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
            SD_CUDA(cudaMemsetAsync(dStats.p + ST_ERR_CAP, 0, sizeof(unsigned long long), st));
            const uint32_t init2[4] = {0u, hN[1], 0u, 0u};
            SD_CUDA(cudaMemcpyAsync(dNext.p, init2, sizeof init2, cudaMemcpyHostToDevice, st));
            PrefilterArgs b = a;
            b.regionCap = regionCap * 16; b.candCap = std::min(candCap * 16, 1u << 23);
            if ((rc = allocScratch(nCta2, b.regionCap, b.candCap, b.keepCap))) return rc;
            b.pool = dPool.p; b.cand = dCand.p; b.candSeq = dCandSeq.p; b.sortIdx = dSortIdx.p; b.keep = dKeep.p; b.resT = dResT.p; b.resScore = dResScore.p;
            b.resAux = dResAux.p; b.qpackG = dQpack.p; b.posInfo = dPosInfo.p; b.posPrefix = dPosPrefix.p;
            b.order = dOrder3.p; b.nQ = nFail;
            prefilterKernel<<<nCta2, PF_THREADS, dynSmem, st>>>(b);
            ctx->launches++;
            SD_CUDA(cudaGetLastError());
            SD_CUDA(cudaStreamSynchronize(st));     // init2 and the second-pass buffers go out of scope
        }
    }
    if (envPersist) {
        cudaStreamAttrValue av{};
        av.accessPolicyWindow.num_bytes = 0;
        cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av);
    }
  }
    SD_CUDA(cudaEventRecord(evk1, st));

    std::vector<uint32_t> hCount(nq);
    unsigned long long hStats[ST_N];
    uint32_t hNext[4];
    SD_CUDA(cudaMemcpyAsync(hCount.data(), dOutCount.p, nq * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    SD_CUDA(cudaMemcpyAsync(hStats, dStats.p, sizeof hStats, cudaMemcpyDeviceToHost, st));
    SD_CUDA(cudaMemcpyAsync(hNext, dNext.p, sizeof hNext, cudaMemcpyDeviceToHost, st));
    SD_CUDA(cudaStreamSynchronize(st));
    if (hStats[ST_ERR_OVERFLOW]) {
        setError("prefilter: %llu queries overflow the reference's hit buffer (maxDbMatches = %llu index entries) more than %u times (first: query %u): "
                 "not followed", hStats[ST_ERR_OVERFLOW], maxDbMatches, PF_MAXBOUND, hNext[1]);
        return SD_ELIMIT;
    }
    if (hStats[ST_ERR_CAP]) { setError("prefilter: %llu queries overflow the enlarged hit / candidate scratch of the second pass", hStats[ST_ERR_CAP]); return SD_ELIMIT; }
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
    if (getenv("SD_PF_DEBUG"))
        fprintf(stderr, "[sd] prefilter phase cycles (sum over queries, M): P1 %.1f  P2-4 %.1f  P5 %.1f (select %.1f)  chunked queries %llu\n",
                hStats[ST_CYC_P1] / 1e6, hStats[ST_CYC_P24] / 1e6, hStats[ST_CYC_P5] / 1e6, hStats[ST_CYC_SEL] / 1e6, hStats[ST_CHUNKED]);
    if (getenv("SD_PF_DEBUG"))
        fprintf(stderr, "[sd] prefilter P1, warp 0 (M cycles): set-up %.1f  stage B %.1f (%llu calls, %llu hit rows)  queued rows incl. stage B %.1f  big rows %.1f  tail %.1f\n",
                hStats[ST_CYC_P0B] / 1e6, hStats[ST_CYC_STAGEB] / 1e6, hStats[ST_N_STAGEB], hStats[ST_N_HITROWS], hStats[ST_CYC_DRAIN] / 1e6,
                hStats[ST_CYC_BIG] / 1e6, hStats[ST_CYC_TAIL] / 1e6);
    if (stats) {
        float msK = 0, msT = 0;
        cudaEventElapsedTime(&msK, evk0, evk1);
        cudaEventElapsedTime(&msT, ev0, ev1);
        stats->probes = hStats[ST_PROBES]; stats->matches = hStats[ST_MATCHES]; stats->diag_residues = hStats[ST_DIAGRES];
        stats->query_residues = total; stats->hits = hStats[ST_HITS]; stats->candidates = hStats[ST_CANDS];
        stats->truncated = hStats[ST_TRUNC]; stats->kernel_ms = msK; stats->total_ms = msT;
    }
    return SD_OK;
}

extern "C" int sd_prefilter_run_batch(sd_context *ctx, const sd_target_index *idx, const sd_query_batch *queries, uint32_t query_first,
                                      const sd_prefilter_params *par, sd_prefilter_hit *hits, uint64_t cap, uint64_t *n_hits,
                                      sd_prefilter_stats *stats) {
    return sd_prefilter_core(ctx, idx, queries, query_first, par, hits, cap, n_hits, stats, nullptr);
}

// ------------------------------------------------------------------------------------------------
// persisted GPU-ready index (the reference's counterpart: createindex / PrefilteringIndexReader, which stores the CPU index table
// next to the DB so that repeated searches skip IndexBuilder::fillDatabase).  One flat file: header, target offsets, then the
// device buffers exactly as the kernels read them (masked + unmasked residue codes, occupancy bitmap {bits, rank}, bucket rows,
// entries).  Loading is a read + upload; nothing is recomputed.
// ------------------------------------------------------------------------------------------------
namespace {
struct IndexFileHeader {
    char magic[8];             // "SDGPUIX3"
    uint32_t version, nT;
    uint64_t totalRes, nEntries, masked;
    int32_t kmerThr;
    uint32_t maxLen, nDistinct, bmWords;
    float sensitivity, maskProb;
    int32_t mask, pad;
};
}  // namespace

extern "C" int sd_target_index_save(const sd_target_index *idx, const sd_prefilter_params *par, const char *path) {
    if (!idx || !par || !path) { setError("sd_target_index_save: bad argument"); return SD_EINVAL; }
    sd_context *ctx = idx->ctx;
    SD_CUDA(cudaSetDevice(ctx->device));
    FILE *f = fopen(path, "wb");
    if (!f) { setError("cannot write %s", path); return SD_EINVAL; }
    IndexFileHeader h{};
    memcpy(h.magic, "SDGPUIX3", 8);
    h.version = 3 + SD_KMER_LAYOUT; h.nT = idx->nT; h.totalRes = idx->totalRes; h.nEntries = idx->nEntries; h.masked = idx->masked;
    h.kmerThr = idx->kmerThr; h.maxLen = idx->maxLen; h.nDistinct = idx->nDistinct; h.bmWords = (uint32_t)idx->bm.n;
    h.sensitivity = par->sensitivity; h.maskProb = par->mask_prob; h.mask = par->mask;
    bool ok = fwrite(&h, sizeof h, 1, f) == 1 && fwrite(idx->hostOff.data(), sizeof(uint32_t), (size_t)idx->nT + 1, f) == (size_t)idx->nT + 1;
    auto dump = [&](const void *dev, size_t bytes) {
        std::vector<char> host(bytes);
        if (cudaMemcpy(host.data(), dev, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) return false;
        return bytes == 0 || fwrite(host.data(), 1, bytes, f) == bytes;
    };
    ok = ok && dump(idx->res.p, idx->totalRes + 32) && dump(idx->raw.p, idx->totalRes + 32) && dump(idx->bm.p, idx->bm.n * sizeof(uint2)) &&
         dump(idx->rows.p, ((size_t)idx->nDistinct + 1) * sizeof(uint2)) && dump(idx->ent.p, (size_t)idx->nEntries * sizeof(uint2));
    ok = (fclose(f) == 0) && ok;
    if (!ok) { setError("writing %s failed", path); return SD_EINVAL; }
    return SD_OK;
}

extern "C" int sd_target_index_load(sd_context *ctx, const char *path, const sd_prefilter_params *par, sd_target_index **out) {
    if (!ctx || !path || !par || !out) { setError("sd_target_index_load: bad argument"); return SD_EINVAL; }
    *out = nullptr;
    SD_CUDA(cudaSetDevice(ctx->device));
    sd::StreamScope allocScope(ctx->stream);
    FILE *f = fopen(path, "rb");
    if (!f) { setError("cannot read %s", path); return SD_EINVAL; }
    std::unique_ptr<FILE, int (*)(FILE *)> closer(f, fclose);
    IndexFileHeader h;
    if (fread(&h, sizeof h, 1, f) != 1 || memcmp(h.magic, "SDGPUIX3", 8) != 0 || h.version != 3 + SD_KMER_LAYOUT) { setError("%s is not a libsdgpu index file (version 3)", path); return SD_EINVAL; }
    // the index depends on the sensitivity (k-mer threshold) and on the masking parameters: refuse a file built with others
    if (h.kmerThr != kmerThreshold(par->sensitivity, KMER) || h.mask != par->mask || (par->mask && h.maskProb != par->mask_prob)) {
        setError("%s was built with -s %g --mask %d --mask-prob %g; the search asks for -s %g --mask %d --mask-prob %g", path, h.sensitivity, h.mask,
                 h.maskProb, par->sensitivity, par->mask, par->mask_prob);
        return SD_EINVAL;
    }
    std::unique_ptr<sd_target_index> idx(new sd_target_index());
    idx->ctx = ctx; idx->nT = h.nT; idx->totalRes = h.totalRes; idx->nEntries = h.nEntries; idx->masked = h.masked;
    idx->kmerThr = h.kmerThr; idx->maxLen = h.maxLen; idx->nDistinct = h.nDistinct;
    idx->hostOff.resize((size_t)h.nT + 1);
    if (fread(idx->hostOff.data(), sizeof(uint32_t), (size_t)h.nT + 1, f) != (size_t)h.nT + 1) { setError("%s is truncated", path); return SD_EINVAL; }
    int rc;
    if ((rc = idx->res.alloc(h.totalRes + 32)) || (rc = idx->raw.alloc(h.totalRes + 32)) || (rc = idx->off.alloc((size_t)h.nT + 1)) ||
        (rc = idx->bm.alloc(h.bmWords)) || (rc = idx->rows.alloc((size_t)h.nDistinct + 1)) || (rc = idx->ent.alloc(h.nEntries)))
        return rc;
    std::vector<char> host;
    auto load = [&](void *dev, size_t bytes) {
        host.resize(bytes);
        if (bytes && fread(host.data(), 1, bytes, f) != bytes) return false;
        return cudaMemcpyAsync(dev, host.data(), bytes, cudaMemcpyHostToDevice, ctx->stream) == cudaSuccess &&
               cudaStreamSynchronize(ctx->stream) == cudaSuccess;
    };
    if (!load(idx->res.p, h.totalRes + 32) || !load(idx->raw.p, h.totalRes + 32) || !load(idx->bm.p, (size_t)h.bmWords * sizeof(uint2)) ||
        !load(idx->rows.p, ((size_t)h.nDistinct + 1) * sizeof(uint2)) || !load(idx->ent.p, (size_t)h.nEntries * sizeof(uint2))) {
        setError("%s is truncated or the upload failed", path);
        return SD_EINVAL;
    }
    SD_CUDA(cudaMemcpyAsync(idx->off.p, idx->hostOff.data(), ((size_t)h.nT + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    SD_CUDA(cudaStreamSynchronize(ctx->stream));
    if ((rc = ensureExt3(ctx)) != SD_OK) return rc;
    *out = idx.release();
    return SD_OK;
}

extern "C" uint32_t sd_target_index_targets(const sd_target_index *idx) { return idx ? idx->nT : 0; }
