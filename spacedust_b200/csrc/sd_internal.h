// Internal declarations shared by the translation units of libsdgpu.so.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <atomic>
#include <mutex>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "../../include/sdgpu.h"
#include "hostprep.h"

struct sd_context;
struct sd_target_index;
struct sd_query_batch;

namespace sd {

void setError(const char *fmt, ...);

#define SD_CUDA(call)                                                                         \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            sd::setError("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return SD_ENODEV;                                                                 \
        }                                                                                     \
    } while (0)

constexpr int KMER = 6;
constexpr int SPAN = 10;                 // spaced seed 1101010011 (Sequence.h:24)
constexpr uint32_t N3 = 8000;            // 20^3
constexpr uint32_t NKMER = 64000000;     // 20^6

// Place of a k-mer in the occupancy bitmap and the row table.  The reference numbers a k-mer first3 + 8000 * second3
// (Indexer.h); the similar-k-mer stream varies the SECOND 3-mer fastest (KmerGenerator.cpp:201-207), so with the reference
// numbering successive probes of a query position land 1000 bytes apart in the bitmap.  Layout 1 stores k-mer
// (first3, second3) at first3 * 8000 + second3: all partners of one first-row neighbour fall into one 1000-byte window of
// the bitmap (8 cache lines), so a probe step of 32 lanes touches a few lines instead of 32 sectors.  Only the storage
// order of the buckets changes (stream order is the probe order), results are identical.
#ifndef SD_KMER_LAYOUT
#define SD_KMER_LAYOUT 1
#endif
#ifdef __CUDACC__
__host__ __device__ __forceinline__ uint32_t kmerSlot(uint32_t first3, uint32_t second3) {
    return SD_KMER_LAYOUT ? first3 * N3 + second3 : first3 + N3 * second3;
}
#endif

// Stream-ordered allocations: entry points set the calling context's stream (StreamScope); buffers then come from
// the device's default memory pool (cudaMallocAsync), which keeps freed blocks for reuse -- no cudaMalloc/cudaFree
// round trips in steady state.
cudaStream_t &tlsAllocStream();
struct StreamScope {
    cudaStream_t prev;
    explicit StreamScope(cudaStream_t s) : prev(tlsAllocStream()) { tlsAllocStream() = s; }
    ~StreamScope() { tlsAllocStream() = prev; }
};

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    cudaStream_t owner = nullptr;
    bool async = false;
    int alloc(size_t count) {
        release();
        if (count == 0) count = 1;
        cudaStream_t s = tlsAllocStream();
        cudaError_t e;
        if (s) { e = cudaMallocAsync(reinterpret_cast<void **>(&p), count * sizeof(T), s); async = true; owner = s; }
        else { e = cudaMalloc(reinterpret_cast<void **>(&p), count * sizeof(T)); async = false; }
        if (e != cudaSuccess) {
            p = nullptr;
            setError("device allocation of %zu bytes failed: %s", count * sizeof(T), cudaGetErrorString(e));
            return SD_ENOMEM;
        }
        n = count;
        return SD_OK;
    }
    void release() {
        if (p) { if (async) cudaFreeAsync(p, owner); else cudaFree(p); }
        p = nullptr;
        n = 0;
    }
    ~DevBuf() { release(); }
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
};

enum ScratchSlot {
    SL_PF_TABLE = 0, SL_PF_CAND, SL_PF_KEEP, SL_PF_REST, SL_PF_RESS, SL_PF_RESD, SL_PF_OUT, SL_PF_DENSE, SL_PF_CSEQ, SL_PF_QPACK, SL_PF_POSINFO, SL_PF_POSPREFIX, SL_PF_SORTIDX,
    SL_AL_FWD, SL_AL_WORD, SL_AL_BIG, SL_AL_REV, SL_AL_RES, SL_AL_HITS, SL_AL_TB, SL_AL_TBRES, SL_AL_IDENT, SL_AL_ORDER,
    SL_AL_PROFF, SL_AL_PROFR, SL_AL_PROFS, SL_AL_BOUND, SL_AL_PAIRQ, SL_AL_PAIRT,
    SL_TB_DIR, SL_TB_ROWS, SL_TB_REV, SL_TB_POOL, SL_SW_KEYA, SL_SW_KEYB, SL_SW_VALA, SL_SW_VALB, SL_SW_ORDER, SL_SW_TMP, SL_SH_SIZES, SL_SH_SEND, SL_SH_RECV, SL_COUNT
};
static_assert(SL_COUNT <= 48, "sd_context::devSlot is too small");
void *ctxDeviceScratch(sd_context *ctx, int slot, size_t bytes);     // nullptr + error message on failure
void *ctxPinnedScratch(sd_context *ctx, int slot, size_t bytes);

// four timing events that are destroyed on every return path
struct EventQuad {
    cudaEvent_t e[4] = {nullptr, nullptr, nullptr, nullptr};
    int create() {
        for (auto &x : e)
            if (cudaEventCreate(&x) != cudaSuccess) { setError("cudaEventCreate failed"); return SD_ENODEV; }
        return SD_OK;
    }
    ~EventQuad() { for (auto &x : e) if (x) cudaEventDestroy(x); }
    EventQuad() = default;
    EventQuad(const EventQuad &) = delete;
    EventQuad &operator=(const EventQuad &) = delete;
};

// DevBuf look-alike backed by a context slot (not freed when it goes out of scope)
template <typename T>
struct ScratchBuf {
    T *p = nullptr;
    size_t n = 0;
    sd_context *ctx;
    int slot;
    ScratchBuf(sd_context *c, int s) : ctx(c), slot(s) {}
    int alloc(size_t count) {
        if (count == 0) count = 1;
        p = static_cast<T *>(ctxDeviceScratch(ctx, slot, count * sizeof(T)));
        n = p ? count : 0;
        return p ? SD_OK : SD_ENOMEM;
    }
    void release() {}
};

// run fn(i) for i in [0,n) on a pool of host threads (small per-item work, chunked)
template <typename F>
inline void parallelFor(uint64_t n, F fn) {
    unsigned nThreads = hostThreads();
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


// few heavy items (genome pairs, ...): one item per grab, always threaded
template <typename F>
inline void parallelItems(uint64_t n, F fn) {
    unsigned nThreads = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(hostThreads(), n));
    std::atomic<uint64_t> next{0};
    auto work = [&]() {
        for (;;) {
            const uint64_t i = next.fetch_add(1);
            if (i >= n) break;
            fn(i);
        }
    };
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < nThreads; t++) pool.emplace_back(work);
    work();
    for (auto &t : pool) t.join();
}

// byte vector whose resize() does not zero-fill (backtrace pools are ~100 MB and completely overwritten)
template <typename T>
struct NoInitAlloc {
    using value_type = T;
    NoInitAlloc() = default;
    template <typename U> NoInitAlloc(const NoInitAlloc<U> &) {}
    T *allocate(size_t n) { return static_cast<T *>(::operator new(n * sizeof(T))); }
    void deallocate(T *p, size_t) { ::operator delete(p); }
    template <typename U, typename... A> void construct(U *p, A &&...a) { if (sizeof...(A)) ::new ((void *)p) U(std::forward<A>(a)...); }
    template <typename U> bool operator==(const NoInitAlloc<U> &) const { return true; }
    template <typename U> bool operator!=(const NoInitAlloc<U> &) const { return false; }
};
using CharPool = std::vector<char, NoInitAlloc<char>>;

// run fn(chunk, lo, hi) over [0,n) split into nChunks contiguous ranges (deterministic partition, any thread order)
template <typename F>
inline void parallelChunks(uint64_t n, unsigned nChunks, F fn) {
    if (nChunks == 0) nChunks = 1;
    unsigned nThreads = hostThreads();
    if (n < 4096) nThreads = 1;
    std::atomic<unsigned> next{0};
    auto work = [&]() {
        for (;;) {
            const unsigned c = next.fetch_add(1);
            if (c >= nChunks) break;
            const uint64_t lo = n * c / nChunks, hi = n * (c + 1) / nChunks;
            if (hi > lo) fn(c, lo, hi);
        }
    };
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < nThreads; t++) pool.emplace_back(work);
    work();
    for (auto &t : pool) t.join();
}

struct SwBatch;
struct SwTask;
struct SwResult;
int launchSw(sd_context *ctx, const SwBatch &batch, const SwTask *dTasks, SwResult *dResults, uint32_t nTasks, int *dBoundary,
             uint32_t boundaryStride, uint32_t nWarpsAllocated, uint32_t *dNext, unsigned long long *dCells, int packed16, int keyed);
int launchSwProfiles(sd_context *ctx, const uint8_t *code, const int8_t *cb, const uint32_t *qOff, uint32_t nQ,
                     uint32_t totalRes, const int8_t *mat, const uint64_t *profOff, const uint32_t *profStride,
                     int8_t *fwd, int8_t *rev, const uint64_t *stgOff, const uint32_t *stgStride, int8_t *stg);
int launchAlignCompositionBias(sd_context *ctx, const float *bias, uint32_t totalRes, int8_t *cb);
uint32_t swWarpsPerCta();
int launchTantan(sd_context *ctx, uint8_t *dRes, const uint32_t *dOff, const std::vector<uint32_t> &hostOff, uint32_t n,
                 const SubMat &seedMat, double minMaskProb, uint64_t *maskedOut);

}  // namespace sd

struct sd_context {
    // persistent scratch: large per-call work buffers are kept between calls (slot = fixed purpose), so the steady
    // state has no multi-GB allocations; released with the context
    struct Slot { void *p = nullptr; size_t cap = 0; };
    Slot devSlot[48];
    Slot pinSlot[8];
    int device = 0;
    int smCount = 148;
    cudaStream_t stream = nullptr;
    uint64_t launches = 0;
    // fused clustersearch: a second context (own stream + scratch) runs alternate query chunks on a helper thread; the token
    // serialises the GPU-heavy part of the align stage so that one chunk's host post-processing overlaps the other's kernels
    sd_context *helper = nullptr;
    // multi-GPU (csrc/shard.cu): NCCL communicator of this rank (ncclComm_t), one process per GPU
    void *comm = nullptr;
    int rank = 0, world = 1;
    std::mutex *gpuToken = nullptr;
    // seed / diagonal matrices and the sorted 3-mer neighbour table are per context (device resident)
    sd::SubMat seedMat;       // VTML80 x8, bias -0.2   (Prefiltering.cpp:68)
    sd::SubMat diagMat;       // BLOSUM62 x2, bias -0.2 (Prefiltering.cpp:69)
    sd::SubMat alnMat;        // BLOSUM62 x2, bias 0    (Alignment.cpp:152)
    sd::DevBuf<uint32_t> e3pair;     // sorted 3-mer neighbour rows 8000 x 8000: score | index << 16 (256 MB)
    sd::DevBuf<uint16_t> e3cnt;      // 8000 x 512 cumulative score counts of the rows (8 MB, L2 resident)
    bool haveExt3 = false;
    // align: smallest passing raw score per query length for (database residues, -e)
    double minScoreDb = -1, minScoreThr = -1;
    std::vector<int32_t> minScoreByLen;
    // sd_align_run[_batch]: a call that ends with SD_ESMALL keeps its result here, so that the caller's retry with larger buffers
    // (same arguments) copies it out instead of aligning everything again
    std::vector<sd_alignment> esmallAlns;
    sd::CharPool esmallPool;
    sd_align_stats esmallStats;
    uint64_t esmallKey = 0;
};

struct sd_query_batch {
    sd_context *ctx = nullptr;
    uint32_t nq = 0;
    uint64_t total = 0;
    std::vector<uint32_t> hostOff;     // nq+1
    std::vector<uint8_t> hostCode;     // residue codes on the host (identity shortcut of align)
    sd::DevBuf<uint8_t> code;          // residue codes, concatenated (+16 pad)
    sd::DevBuf<uint32_t> off;          // nq+1
};

int sd_prefilter_core(sd_context *ctx, const sd_target_index *idx, const sd_query_batch *queries, uint32_t query_first,
                      const sd_prefilter_params *par, sd_prefilter_hit *hits, uint64_t cap, uint64_t *n_hits,
                      sd_prefilter_stats *stats, sd::DevBuf<sd_prefilter_hit> *keepOnDevice, std::vector<uint64_t> *hitOffsets = nullptr);
// merged result of an exchange: one chunk per rank block ([4 x u64 counts][alns][pool][clusters][members], 16-byte aligned parts);
// the chunks are VIEWS into `blocks` (the context's pinned receive buffer), valid until the next exchange on that context
namespace sd {
int shardAllgather(sd_context *ctx, const sd_search_result *mine, uint32_t query_first, int pinSlot, sd_search_result **all, uint64_t *bytes_gathered);
}
int sd_search_result_from_blocks(const char *blocks, uint64_t stride, const uint64_t *sizes, int n, sd_search_result **out);
int sd_align_device_pairs(sd_context *ctx, const sd_target_index *idx, const sd_query_batch *queries, uint32_t query_first,
                          const uint32_t *dPairQ, const uint32_t *dPairT, uint64_t n_pairs, uint64_t db_residues,
                          const sd_align_params *par, std::vector<sd_alignment> &acc, sd::CharPool &pool,
                          sd_align_stats *stats);
int sd_align_device_hits(sd_context *ctx, const sd_target_index *idx, const sd_query_batch *queries, uint32_t query_first,
                         const sd_prefilter_hit *dHits, uint64_t n_hits, uint64_t db_residues, const sd_align_params *par,
                         std::vector<sd_alignment> &acc, sd::CharPool &pool, sd_align_stats *stats);

struct sd_target_index {
    sd_context *ctx = nullptr;
    uint32_t nT = 0;
    uint64_t totalRes = 0;
    uint64_t nEntries = 0;
    uint64_t masked = 0;
    int kmerThr = 0;
    std::vector<uint32_t> hostOff;     // nT+1
    sd::DevBuf<uint8_t> res;           // masked residue codes, concatenated (+ SPAN pad): prefilter side
    sd::DevBuf<uint8_t> raw;           // unmasked residue codes: alignment side
    uint32_t maxLen = 0;
    sd::DevBuf<uint32_t> off;          // nT+1
    uint32_t nDistinct = 0;            // k-mers with a non-empty bucket
    sd::DevBuf<uint2> bm;              // per 32 k-mers: {occupancy bits, row of the first non-empty k-mer} (16 MB, L2 resident)
    sd::DevBuf<uint2> rows;            // per non-empty k-mer: inline single entry or {start, length} into ent
    sd::DevBuf<uint2> ent;             // bucket entries: {target id, first position of the k-mer in that target}
};
