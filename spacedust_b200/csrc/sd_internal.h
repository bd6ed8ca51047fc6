// Internal declarations shared by the translation units of libsdgpu.so.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
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

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    int alloc(size_t count) {
        release();
        if (count == 0) count = 1;
        cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&p), count * sizeof(T));
        if (e != cudaSuccess) {
            p = nullptr;
            setError("cudaMalloc(%zu bytes) failed: %s", count * sizeof(T), cudaGetErrorString(e));
            return SD_ENOMEM;
        }
        n = count;
        return SD_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    ~DevBuf() { release(); }
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
};

struct SwBatch;
struct SwTask;
struct SwResult;
int launchSw(sd_context *ctx, const SwBatch &batch, const SwTask *dTasks, const uint32_t *dOrder, SwResult *dResults,
             uint32_t nTasks, int *dBoundary, uint32_t boundaryStride, uint32_t nWarpsAllocated, uint32_t *dNext,
             unsigned long long *dCells, int packed16);
int launchSwProfiles(sd_context *ctx, const uint8_t *code, const int8_t *cb, const uint32_t *qOff, uint32_t nQ,
                     uint32_t totalRes, const int8_t *mat, const uint64_t *profOff, const uint32_t *profStride,
                     int8_t *fwd, int8_t *rev);
int launchAlignCompositionBias(sd_context *ctx, const float *bias, uint32_t totalRes, int8_t *cb);
uint32_t swWarpsPerCta();

}  // namespace sd

struct sd_context {
    int device = 0;
    int smCount = 148;
    cudaStream_t stream = nullptr;
    uint64_t launches = 0;
    // seed / diagonal matrices and the sorted 3-mer neighbour table are per context (device resident)
    sd::SubMat seedMat;       // VTML80 x8, bias -0.2   (Prefiltering.cpp:68)
    sd::SubMat diagMat;       // BLOSUM62 x2, bias -0.2 (Prefiltering.cpp:69)
    sd::SubMat alnMat;        // BLOSUM62 x2, bias 0    (Alignment.cpp:152)
    sd::DevBuf<int16_t> e3score;
    sd::DevBuf<uint16_t> e3index;
    bool haveExt3 = false;
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
                      sd_prefilter_stats *stats, sd::DevBuf<sd_prefilter_hit> *keepOnDevice);
int sd_align_device_pairs(sd_context *ctx, const sd_target_index *idx, const sd_query_batch *queries, uint32_t query_first,
                          const uint32_t *dPairQ, const uint32_t *dPairT, uint64_t n_pairs, uint64_t db_residues,
                          const sd_align_params *par, std::vector<sd_alignment> &acc, std::vector<char> &pool,
                          sd_align_stats *stats);
int sd_align_device_hits(sd_context *ctx, const sd_target_index *idx, const sd_query_batch *queries, uint32_t query_first,
                         const sd_prefilter_hit *dHits, uint64_t n_hits, uint64_t db_residues, const sd_align_params *par,
                         std::vector<sd_alignment> &acc, std::vector<char> &pool, sd_align_stats *stats);

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
    sd::DevBuf<uint32_t> kOff;         // NKMER+1 bucket starts
    sd::DevBuf<uint32_t> entSeq;       // bucket entries: target id ...
    sd::DevBuf<uint16_t> entPos;       // ... and first position of the k-mer in that target
};
