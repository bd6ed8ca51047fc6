// C ABI plumbing of libsdgpu.so: error slot, context life cycle.
#include <cstdlib>
#include <memory>
#include "sd_internal.h"

#include <cstring>

namespace sd {
static thread_local char g_err[1024] = "";
void setError(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
void *ctxDeviceScratch(sd_context *ctx, int slot, size_t bytes) {
    sd_context::Slot &s = ctx->devSlot[slot];
    if (s.cap >= bytes && s.p) return s.p;
    if (s.p) { cudaStreamSynchronize(ctx->stream); cudaFree(s.p); s.p = nullptr; s.cap = 0; }
    const size_t want = bytes + bytes / 8 + 4096;
    cudaError_t e = cudaMalloc(&s.p, want);
    if (e != cudaSuccess) { s.p = nullptr; setError("device scratch allocation of %zu bytes failed: %s", want, cudaGetErrorString(e)); return nullptr; }
    s.cap = want;
    return s.p;
}

void *ctxPinnedScratch(sd_context *ctx, int slot, size_t bytes) {
    sd_context::Slot &s = ctx->pinSlot[slot];
    if (s.cap >= bytes && s.p) return s.p;
    if (s.p) { cudaStreamSynchronize(ctx->stream); cudaFreeHost(s.p); s.p = nullptr; s.cap = 0; }
    const size_t want = bytes + bytes / 8 + 4096;
    cudaError_t e = cudaHostAlloc(&s.p, want, cudaHostAllocDefault);
    if (e != cudaSuccess) { s.p = nullptr; setError("pinned host allocation of %zu bytes failed: %s", want, cudaGetErrorString(e)); return nullptr; }
    s.cap = want;
    return s.p;
}

cudaStream_t &tlsAllocStream() {
    static thread_local cudaStream_t s = nullptr;
    return s;
}
}  // namespace sd

extern "C" const char *sd_last_error(void) { return sd::g_err; }
extern "C" int sd_version(void) { return 100; }

extern "C" int sd_context_create(int device, sd_context **out) {
    if (!out) { sd::setError("sd_context_create: out is NULL"); return SD_EINVAL; }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        sd::setError("no CUDA device available (%s); libsdgpu has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "0 devices");
        return SD_ENODEV;
    }
    if (device < 0) SD_CUDA(cudaGetDevice(&device));
    if (device >= count) { sd::setError("device %d out of range (%d devices)", device, count); return SD_EINVAL; }
    SD_CUDA(cudaSetDevice(device));
    std::unique_ptr<sd_context> holder(new sd_context());      // released to the caller at the end: no leak on an early return
    sd_context *ctx = holder.get();
    ctx->device = device;
    cudaDeviceProp prop;
    SD_CUDA(cudaGetDeviceProperties(&prop, device));
    ctx->smCount = prop.multiProcessorCount;
    SD_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    {   // keep freed device blocks in the pool instead of returning them to the driver at every synchronisation
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    sd::buildSubMat(sd::VTML80, 8.0f, -0.2f, ctx->seedMat);
    sd::buildSubMat(sd::BLOSUM62, 2.0f, -0.2f, ctx->diagMat);
    sd::buildSubMat(sd::BLOSUM62, 2.0f, 0.0f, ctx->alnMat);
    *out = holder.release();
    return SD_OK;
}

extern "C" void sd_context_destroy(sd_context *ctx) {
    if (!ctx) return;
    if (ctx->helper) { sd_context_destroy(ctx->helper); ctx->helper = nullptr; }
    sd_shard_finalize(ctx);
    cudaSetDevice(ctx->device);
    for (auto &sl : ctx->devSlot) if (sl.p) cudaFree(sl.p);
    for (auto &sl : ctx->pinSlot) if (sl.p) cudaFreeHost(sl.p);
    ctx->e3pair.release();
    ctx->e3cnt.release();
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" uint64_t sd_context_launches(const sd_context *ctx) { return ctx ? ctx->launches : 0; }
