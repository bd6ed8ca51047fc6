// tantan repeat masking of the target residues on the device (Masker::maskSequence -> tantan::maskSequences,
// lib/mmseqs/src/commons/Masker.cpp:52-92, lib/tantan/tantan.cpp:44-130,520-535; parameters 50, 0.005, 0.05, 0.9, gap
// probabilities 0, --mask-prob).  Arithmetic as in the reference build:
// doubles, no FMA contraction (--fmad=false), the foreground sum associated like the reference's AVX2
// build -- four interleaved partial sums reduced as (l0+l2)+(l1+l3), then a scalar tail -- so the posteriors and
// therefore the masked residues are bit-identical (tests compare them with the oracle restatement in oracle/orc_prefilter.c).
//
// One THREAD per sequence (sequences sorted by length so a warp's threads finish together): the 50 repeat-offset
// states live in registers (fully unrolled), the previous 56 residues in seven 64-bit shift registers, the 21x21
// likelihood-ratio table in shared memory.  Per position ~50 x (2 mul + 1 add + 1 mul): FP64-pipe work, a few ms for
// a 20 M residue DB instead of ~100 ms on 16 host cores.
#include <algorithm>
#include <numeric>
#include <cmath>

#include "sd_internal.h"

namespace sd {

namespace {

constexpr int TW = 50;          // repeat offsets

struct TantanArgs {
    uint8_t *res;               // residue codes, masked in place
    const uint32_t *off;        // n+1
    const uint32_t *order;      // sequence ids, longest first
    uint32_t n;
    const double *lr;           // AA*AA likelihood ratios
    const double *toFg;         // TW transition probabilities into the repeat states
    double bb, fb, ff;          // background->background, foreground->background, foreground->foreground
    double minMaskProb;
    float *post;                // per residue
    double *scale;              // per sequence: L/16+1 entries at (off>>4)+seq
    unsigned long long *masked;
};

__device__ __forceinline__ int histAt(const uint64_t (&h)[7], int i) { return (int)((h[i >> 3] >> ((i & 7) * 8)) & 0xffu); }

__global__ void __launch_bounds__(128) tantanKernel(TantanArgs A) {
    __shared__ double sLr[AA * AA];
    __shared__ double sToFg[TW];
    for (int i = threadIdx.x; i < AA * AA; i += blockDim.x) sLr[i] = A.lr[i];
    for (int i = threadIdx.x; i < TW; i += blockDim.x) sToFg[i] = A.toFg[i];
    __syncthreads();
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= A.n) return;
    const uint32_t s = A.order[tid];
    const uint32_t o = A.off[s];
    const int L = (int)(A.off[s + 1] - o);
    if (L <= 0) return;
    uint8_t *seq = A.res + o;
    float *post = A.post + o;
    double *scale = A.scale + (o >> 4) + s;
    const double bb = A.bb, fb = A.fb, ff = A.ff;

    double fg[TW];
#pragma unroll
    for (int i = 0; i < TW; i++) fg[i] = 0.0;
    uint64_t h[7] = {0, 0, 0, 0, 0, 0, 0};       // h byte i = residue at p-1-i
    double bg = 1.0;
    for (int p = 0; p < L; p++) {
        const int cur = seq[p];
        const double *row = sLr + cur * AA;
        const int reach = min(p, TW);
        const double b = bg;
        double l0 = 0, l1 = 0, l2 = 0, l3 = 0;
#pragma unroll
        for (int i = 0; i + 4 <= TW; i += 4) {
            if (i + 4 <= reach) {
                { const double f = fg[i + 0]; l0 = l0 + f; fg[i + 0] = (b * sToFg[i + 0] + f * ff) * row[histAt(h, i + 0)]; }
                { const double f = fg[i + 1]; l1 = l1 + f; fg[i + 1] = (b * sToFg[i + 1] + f * ff) * row[histAt(h, i + 1)]; }
                { const double f = fg[i + 2]; l2 = l2 + f; fg[i + 2] = (b * sToFg[i + 2] + f * ff) * row[histAt(h, i + 2)]; }
                { const double f = fg[i + 3]; l3 = l3 + f; fg[i + 3] = (b * sToFg[i + 3] + f * ff) * row[histAt(h, i + 3)]; }
            }
        }
        double fromFg = (l0 + l2) + (l1 + l3);
        const int reach4 = reach & ~3;
#pragma unroll
        for (int i = 0; i < TW; i++) {
            if (i >= reach4 && i < reach) {
                const double f = fg[i];
                fromFg += f;
                fg[i] = (b * sToFg[i] + f * ff) * row[histAt(h, i)];
            }
        }
        bg = b * bb + fromFg * fb;
        if ((p & 15) == 15) {
            const double sc = 1 / bg;
            scale[p >> 4] = sc;
            bg *= sc;
#pragma unroll
            for (int k = 0; k < TW; k++) fg[k] *= sc;
        }
        post[p] = (float)bg;
#pragma unroll
        for (int w = 6; w > 0; w--) h[w] = (h[w] << 8) | (h[w - 1] >> 56);
        h[0] = (h[0] << 8) | (uint64_t)cur;
    }
    double fgTotal = 0.0;
#pragma unroll
    for (int i = 0; i < TW; i++) fgTotal += fg[i];
    const double z = bg * bb + fgTotal * fb;
    bg = bb;
#pragma unroll
    for (int i = 0; i < TW; i++) fg[i] = fb;
    // backward pass: h must hold the residues before p; rebuild it for p = L-1 and then slide it the other way
#pragma unroll
    for (int w = 0; w < 7; w++) h[w] = 0;
    for (int i = 55; i >= 0; i--) {       // h byte i = seq[L-2-i]
        const int q = L - 2 - i;
        const uint64_t v = q >= 0 ? (uint64_t)seq[q] : 0ull;
#pragma unroll
        for (int w = 6; w > 0; w--) h[w] = (h[w] << 8) | (h[w - 1] >> 56);
        h[0] = (h[0] << 8) | v;
    }
    unsigned long long nMasked = 0;
    for (int p = L - 1; p >= 0; p--) {
        const double notRepeat = post[p] * bg / z;
        const float pr = 1 - (float)notRepeat;
        if ((p & 15) == 15) {
            const double sc = scale[p >> 4];
            bg *= sc;
#pragma unroll
            for (int k = 0; k < TW; k++) fg[k] *= sc;
        }
        const int cur = seq[p];
        const double *row = sLr + cur * AA;
        const int reach = min(p, TW);
        const double toBg = fb * bg;
        double l0 = 0, l1 = 0, l2 = 0, l3 = 0;
#pragma unroll
        for (int i = 0; i + 4 <= TW; i += 4) {
            if (i + 4 <= reach) {
                { const double f = fg[i + 0] * row[histAt(h, i + 0)]; l0 = l0 + sToFg[i + 0] * f; fg[i + 0] = toBg + ff * f; }
                { const double f = fg[i + 1] * row[histAt(h, i + 1)]; l1 = l1 + sToFg[i + 1] * f; fg[i + 1] = toBg + ff * f; }
                { const double f = fg[i + 2] * row[histAt(h, i + 2)]; l2 = l2 + sToFg[i + 2] * f; fg[i + 2] = toBg + ff * f; }
                { const double f = fg[i + 3] * row[histAt(h, i + 3)]; l3 = l3 + sToFg[i + 3] * f; fg[i + 3] = toBg + ff * f; }
            }
        }
        double toFg = (l0 + l2) + (l1 + l3);
        const int reach4 = reach & ~3;
#pragma unroll
        for (int i = 0; i < TW; i++) {
            if (i >= reach4 && i < reach) {
                const double f = fg[i] * row[histAt(h, i)];
                toFg += sToFg[i] * f;
                fg[i] = toBg + ff * f;
            }
        }
        bg = bb * bg + toFg;
        // later iterations only read residues before p (current letter and history), so masking position p in place is safe
        if ((double)pr >= A.minMaskProb) { seq[p] = (uint8_t)XCODE; nMasked++; }
        // slide the history one position towards the sequence start: drop seq[p-1], append seq[p-57]
        const int q = p - 57;
        const uint64_t in = q >= 0 ? (uint64_t)seq[q] : 0ull;
#pragma unroll
        for (int w = 0; w < 6; w++) h[w] = (h[w] >> 8) | (h[w + 1] << 56);
        h[6] = (h[6] >> 8) | (in << 56);
    }
    if (nMasked) atomicAdd(A.masked, nMasked);
}

}  // namespace

// Mask all targets in place on the device.  dRes: residue codes (concatenated), dOff/hostOff: n+1 offsets.
int launchTantan(sd_context *ctx, uint8_t *dRes, const uint32_t *dOff, const std::vector<uint32_t> &hostOff, uint32_t n,
                 const SubMat &seedMat, double minMaskProb, uint64_t *maskedOut) {
    *maskedOut = 0;
    if (n == 0) return SD_OK;
    cudaStream_t st = ctx->stream;
    double lr[AA * AA], toFg[TW];
    likelihoodRatios(seedMat, lr);
    const double pRepeat = 0.005, pEnd = 0.05, decay = 0.9;
    {
        double p = pRepeat * ((1 - decay) / (1 - std::pow(decay, TW)));
        for (int i = 0; i < TW; i++) { toFg[i] = p; p *= decay; }
    }
    // longest first: counting sort on the length
    uint32_t maxLen = 0;
    for (uint32_t i = 0; i < n; i++) maxLen = std::max(maxLen, hostOff[i + 1] - hostOff[i]);
    std::vector<uint32_t> head(maxLen + 2, 0), order(n);
    for (uint32_t i = 0; i < n; i++) head[maxLen - (hostOff[i + 1] - hostOff[i]) + 1]++;
    for (uint32_t l = 0; l <= maxLen; l++) head[l + 1] += head[l];
    for (uint32_t i = 0; i < n; i++) order[head[maxLen - (hostOff[i + 1] - hostOff[i])]++] = i;
    const uint64_t total = hostOff[n];
    DevBuf<uint32_t> dOrder;
    DevBuf<double> dLr, dToFg, dScale;
    DevBuf<float> dPost;
    DevBuf<unsigned long long> dMasked;
    int rc;
    if ((rc = dOrder.alloc(n)) || (rc = dLr.alloc(AA * AA)) || (rc = dToFg.alloc(TW)) || (rc = dScale.alloc((total >> 4) + n + 2)) ||
        (rc = dPost.alloc(total + 1)) || (rc = dMasked.alloc(1)))
        return rc;
    SD_CUDA(cudaMemcpyAsync(dOrder.p, order.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    SD_CUDA(cudaMemcpyAsync(dLr.p, lr, sizeof lr, cudaMemcpyHostToDevice, st));
    SD_CUDA(cudaMemcpyAsync(dToFg.p, toFg, sizeof toFg, cudaMemcpyHostToDevice, st));
    SD_CUDA(cudaMemsetAsync(dMasked.p, 0, sizeof(unsigned long long), st));
    TantanArgs A;
    A.res = dRes; A.off = dOff; A.order = dOrder.p; A.n = n; A.lr = dLr.p; A.toFg = dToFg.p;
    A.bb = 1 - pRepeat; A.fb = pEnd; A.ff = 1 - pEnd; A.minMaskProb = minMaskProb;
    A.post = dPost.p; A.scale = dScale.p; A.masked = dMasked.p;
    tantanKernel<<<(n + 127) / 128, 128, 0, st>>>(A);
    ctx->launches++;
    SD_CUDA(cudaGetLastError());
    unsigned long long m = 0;
    SD_CUDA(cudaMemcpyAsync(&m, dMasked.p, sizeof m, cudaMemcpyDeviceToHost, st));
    SD_CUDA(cudaStreamSynchronize(st));       // order/lr/toFg are host stack/heap objects
    *maskedOut = m;
    return SD_OK;
}

}  // namespace sd
