// Small preparation kernels shared by the prefilter and alignment translation units (static: one copy per TU).
#pragma once
#include <cstdint>

#include "hostprep.h"

namespace sd {

static __device__ __forceinline__ uint32_t seqOfPosition(const uint32_t *off, uint32_t n, uint32_t gp) {
    uint32_t lo = 0, hi = n;   // find s with off[s] <= gp < off[s+1]
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (off[mid] <= gp) lo = mid; else hi = mid;
    }
    return lo;
}


static __global__ void mapResidues(const char *__restrict__ ascii, uint64_t n, const uint8_t *__restrict__ lut,
                                   uint8_t *__restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = lut[(uint8_t)ascii[i]];
}

// SubstitutionMatrix::calcLocalAaBiasCorrection (commons/SubstitutionMatrix.cpp:79-109): +-20 window; the
// float accumulator is updated through double intermediates; __dmul_rn/__dadd_rn keep nvcc from fusing.
static __global__ void compositionBias(const uint8_t *__restrict__ code, const uint32_t *__restrict__ off, uint32_t nQ,
                                       uint32_t totalRes, const int16_t *__restrict__ sub /*21x21*/,
                                       const double *__restrict__ pBack, int enabled, float *__restrict__ bias) {
    uint32_t gp = blockIdx.x * blockDim.x + threadIdx.x;
    if (gp >= totalRes) return;
    float b = 0.0f;
    if (enabled) {
        uint32_t s = seqOfPosition(off, nQ, gp);
        int i = (int)(gp - off[s]);
        int L = (int)(off[s + 1] - off[s]);
        int lo = max(0, i - 20), hi = min(L, i + 20);
        const uint8_t *seq = code + off[s];
        const int16_t *row = sub + seq[i] * AA;
        int sum = 0;
        for (int j = lo; j < hi; j++) sum += row[seq[j]];
        sum -= row[seq[i]];
        float d = (float)sum;
        d = __double2float_rn(__ddiv_rn((double)d, __dmul_rn(-1.0, (double)(float)(hi - lo))));
        for (int a = 0; a < AA; a++) d = __double2float_rn(__dadd_rn((double)d, __dmul_rn(pBack[a], (double)(float)row[a])));
        b = __fmul_rn(1.0f, d);
    }
    bias[gp] = b;
}

}  // namespace sd
