// Smith-Waterman score/coordinate passes on B200 (sm_100a): one warp per (query, target) pair,
// anti-diagonal wavefront over the target columns, R query rows per lane held in registers, the
// values that cross lane boundaries travel by warp shuffle.
//
// Reference behaviour restated (paths relative to /root/reference/lib/mmseqs/src/alignment):
//   StripedSmithWaterman.cpp:639-940   sw_sse2_byte  (8-bit, 32 interleaved segments on AVX2)
//   StripedSmithWaterman.cpp:943-1214  sw_sse2_word  (16-bit, 16 segments)
//   StripedSmithWaterman.cpp:163-187   createQueryProfile (score = mat[a][q_i] + compositionBias_i)
// The striped kernels compute exact Gotoh H with one layout-dependent rule: E (the gap that runs along
// the target) is updated from H *before* the lazy-F correction, i.e. H that only knows vertical gaps
// opened inside the same segment of segLen = ceil(L/width) rows.  Both chains are carried per cell:
//   Ft / H   exact vertical-gap chain          -> what the reference ends up storing in H
//   Fm / H1  chain reset at every segment start -> what feeds E
// so scores, end points and the 8-bit overflow decision are bit-identical to the AVX2 build.
#include <climits>

#include "sd_internal.h"
#include "sw.cuh"

namespace sd {

constexpr int SW_WARPS = 4;

struct SwKernelArgs {
    SwBatch b;
    const SwTask *tasks;
    const uint32_t *order;     // task indices, most expensive first
    SwResult *results;
    uint32_t nTasks;
    uint32_t *next;            // work counter
    int *boundary;             // per warp: 2 x 3 x boundaryStride ints (multi-stripe hand-over)
    uint32_t boundaryStride;
    unsigned long long *cells; // total DP cells computed
    int packed16;              // 1: packed 16-bit pass, 0: 32-bit pass (fallback for scores near 2^15)
};

__device__ __forceinline__ uint32_t swLane() { return threadIdx.x & 31; }

template <int R>
__device__ void swPass(const SwKernelArgs &A, const SwTask &tk, int *bnd, SwResult &res, unsigned long long &cellCount) {
    constexpr int NW = R / 4;                 // profile words per lane and column
    const uint32_t lane = swLane();
    const int nRows = tk.nRows, nCols = tk.nCols;
    const int segLen = (nRows + tk.width - 1) / tk.width;
    const int go = A.b.gapOpen, ge = A.b.gapExtend;
    const uint32_t Lq = A.b.qLen[tk.q];
    const uint32_t stride = A.b.profStride[tk.q];
    const int8_t *prof = (tk.reverse ? A.b.profRev : A.b.profFwd) + A.b.profOff[tk.q];
    const uint32_t rowOffset = tk.reverse ? (Lq - (uint32_t)nRows) : 0u;
    const uint8_t *T = A.b.tRes + A.b.tOff[tk.t];

    int best = 0, bestCol = INT_MAX, bestRow = INT_MAX;
    bool overflow = false;
    int pingpong = 0;

    for (int stripe0 = 0; stripe0 < nRows && !overflow; stripe0 += 32 * R) {
        const bool hasPrev = stripe0 > 0;
        const bool hasNext = stripe0 + 32 * R < nRows;
        const int i0 = stripe0 + (int)lane * R;
        const int nValid = min(max(nRows - i0, 0), R);
        const int *bndIn = bnd + (size_t)pingpong * 3 * A.boundaryStride;
        int *bndOut = bnd + (size_t)(pingpong ^ 1) * 3 * A.boundaryStride;
        uint32_t segMask = 0;
#pragma unroll
        for (int k = 0; k < R; k++)
            if (((i0 + k) % segLen) == 0) segMask |= 1u << k;
        // byte masks that turn the rows past the query end into -128 (they can then never hold a maximum)
        uint32_t keep[NW];
#pragma unroll
        for (int w = 0; w < NW; w++) {
            uint32_t m = 0;
#pragma unroll
            for (int b = 0; b < 4; b++)
                if (w * 4 + b < nValid) m |= 0xffu << (8 * b);
            keep[w] = m;
        }
        const uintptr_t rowAddr0 = (uintptr_t)(prof + rowOffset + (uint32_t)i0);
        const uint32_t shiftBits = (uint32_t)(rowAddr0 & 3u) * 8u;
        const bool laneHasRows = nValid > 0;

        int H[R], E[R];
#pragma unroll
        for (int k = 0; k < R; k++) { H[k] = 0; E[k] = 0; }
        int prevHup = 0;
        int outA = 0, outH = 0, outFt = 0, outFm = 0;
        int tChunk = 0, bH = 0, bFt = 0, bFm = 0;
        const int nSteps = nCols + 31;

        for (int s = 0; s < nSteps; s++) {
            if ((s & 31) == 0) {
                const int c = s + (int)lane;
                tChunk = (c < nCols) ? (int)T[tk.reverse ? (nCols - 1 - c) : c] : 0;
                if (hasPrev) {
                    bH = (c < nCols) ? bndIn[c] : 0;
                    bFt = (c < nCols) ? bndIn[A.boundaryStride + c] : 0;
                    bFm = (c < nCols) ? bndIn[2 * A.boundaryStride + c] : 0;
                }
            }
            int inA = __shfl_up_sync(0xffffffffu, outA, 1);
            int inH = __shfl_up_sync(0xffffffffu, outH, 1);
            int inFt = __shfl_up_sync(0xffffffffu, outFt, 1);
            int inFm = __shfl_up_sync(0xffffffffu, outFm, 1);
            const int a0 = __shfl_sync(0xffffffffu, tChunk, s & 31);
            int h0 = 0, ft0 = 0, fm0 = 0;
            if (hasPrev) {
                h0 = __shfl_sync(0xffffffffu, bH, s & 31);
                ft0 = __shfl_sync(0xffffffffu, bFt, s & 31);
                fm0 = __shfl_sync(0xffffffffu, bFm, s & 31);
            }
            if (lane == 0) { inA = a0; inH = h0; inFt = ft0; inFm = fm0; }
            const int c = s - (int)lane;
            if (c >= 0 && c < nCols && laneHasRows) {
                // scores of this lane's rows against letter inA: unaligned read of R bytes as aligned words
                const uint32_t *wp = (const uint32_t *)((rowAddr0 + (uintptr_t)inA * stride) & ~(uintptr_t)3);
                // (stride is a multiple of 4, so the misalignment is the same for every letter)
                uint32_t raw[NW + 1];
#pragma unroll
                for (int w = 0; w <= NW; w++) raw[w] = __ldg(wp + w);
                uint32_t sc[NW];
#pragma unroll
                for (int w = 0; w < NW; w++) {
                    const uint32_t v = __funnelshift_r(raw[w], raw[w + 1], shiftBits);
                    sc[w] = (v & keep[w]) | (0x80808080u & ~keep[w]);
                }
                int diag = prevHup;
                int ft = inFt, fm = inFm;
                int colKey = 0;
#pragma unroll
                for (int k = 0; k < R; k++) {
                    if (segMask & (1u << k)) fm = 0;
                    const int sv = (int)(signed char)((sc[k >> 2] >> (8 * (k & 3))) & 0xffu);
                    const int e = E[k];
                    const int h = __viaddmax_s32(diag, sv, e);       // max(diag + s, E); E >= 0 is the local floor
                    const int h1 = max(h, fm);
                    const int hs = max(h, ft);
                    diag = H[k];
                    H[k] = hs;
                    colKey = max(colKey, hs * 32 + (31 - k));
                    const int open1 = __viaddmax_s32_relu(h1, -go, 0);
                    E[k] = __viaddmax_s32(e, -ge, open1);
                    fm = __viaddmax_s32(fm, -ge, open1);
                    const int openS = __viaddmax_s32_relu(hs, -go, 0);
                    ft = __viaddmax_s32(ft, -ge, openS);
                }
                outA = inA; outH = H[R - 1]; outFt = ft; outFm = fm;
                prevHup = inH;
                const int colBest = colKey >> 5;
                if (colBest > best || (colBest == best && colBest > 0 && c < bestCol)) {
                    best = colBest; bestCol = c; bestRow = i0 + (31 - (colKey & 31));
                }
                if (hasNext && lane == 31) {
                    bndOut[c] = outH;
                    bndOut[A.boundaryStride + c] = outFt;
                    bndOut[2 * A.boundaryStride + c] = outFm;
                }
            }
            if ((s & 31) == 31 && tk.limit != INT_MAX) {
                if (__any_sync(0xffffffffu, best >= tk.limit)) { overflow = true; break; }
            }
        }
        cellCount += (lane == 0) ? (unsigned long long)min(nRows - stripe0, 32 * R) * (unsigned long long)nCols : 0ull;
        pingpong ^= 1;
        __syncwarp();
    }
    // reduce: highest score, then first column, then smallest row
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const int b2 = __shfl_down_sync(0xffffffffu, best, d);
        const int c2 = __shfl_down_sync(0xffffffffu, bestCol, d);
        const int r2 = __shfl_down_sync(0xffffffffu, bestRow, d);
        if (b2 > best || (b2 == best && (c2 < bestCol || (c2 == bestCol && r2 < bestRow)))) { best = b2; bestCol = c2; bestRow = r2; }
    }
    best = __shfl_sync(0xffffffffu, best, 0);
    if (!overflow && tk.limit != INT_MAX && best >= tk.limit) overflow = true;
    if (lane == 0) {
        res.score = best;
        res.col = best > 0 ? bestCol : -1;
        res.row = best > 0 ? bestRow : -1;
        res.overflow = overflow ? 1 : 0;
    }
}


// ------------------------------------------------------------------------------------------------
// Packed 16-bit pass: every lane is TWO wavefront stages (low / high half of a 32-bit register), stage
// 2l works on column step-2l, stage 2l+1 on column step-2l-1, each owning RP consecutive query rows.
// One DPX instruction (VIADDMNMX.S16x2 etc.) updates a cell of both stages.  Columns outside the target
// use a dummy profile row of -128 and rows past the query end read -128 padding, so idle stages keep an
// all-zero state without masks.  Values must stay below 2^15: callers re-run tasks whose score reaches
// SW16_GUARD with the 32-bit pass.
// ------------------------------------------------------------------------------------------------
constexpr int SW16_GUARD = 32000;
constexpr int SW_DUMMY_LETTER = 21;

// prmt.b32 in default mode: selector nibble bit 3 replicates the sign of the selected byte (PTX ISA, prmt)
__device__ __forceinline__ uint32_t prmtSext(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
__device__ __forceinline__ uint32_t max16x2(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("max.s16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t pack16(int lo, int hi) { return ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16); }
__device__ __forceinline__ int lo16(uint32_t v) { return (int)(short)(v & 0xffffu); }
__device__ __forceinline__ int hi16(uint32_t v) { return (int)(short)(v >> 16); }

template <int RP, bool KEYED>
__device__ void swPass16(const SwKernelArgs &A, const SwTask &tk, int *bnd, SwResult &res, unsigned long long &cellCount) {
    constexpr int NWA = (RP + 3) / 4;          // aligned words holding RP bytes
    const uint32_t lane = swLane();
    const int nRows = tk.nRows, nCols = tk.nCols;
    const int segLen = (nRows + tk.width - 1) / tk.width;
    const uint32_t negGo = pack16(-A.b.gapOpen, -A.b.gapOpen), negGe = pack16(-A.b.gapExtend, -A.b.gapExtend);
    const uint32_t Lq = A.b.qLen[tk.q];
    const uint32_t stride = A.b.profStride[tk.q];
    const int8_t *prof = (tk.reverse ? A.b.profRev : A.b.profFwd) + A.b.profOff[tk.q];
    const uint32_t rowOffset = tk.reverse ? (Lq - (uint32_t)nRows) : 0u;
    const uint8_t *T = A.b.tRes + A.b.tOff[tk.t];

    int bestA = 0, bestB = 0, colA = INT_MAX, colB = INT_MAX, rowA = INT_MAX, rowB = INT_MAX;
    bool overflow = false;
    int pingpong = 0;

    for (int stripe0 = 0; stripe0 < nRows && !overflow; stripe0 += 64 * RP) {
        const bool hasPrev = stripe0 > 0;
        const bool hasNext = stripe0 + 64 * RP < nRows;
        const int i0 = stripe0 + (int)lane * 2 * RP;          // stage A rows [i0, i0+RP), stage B rows [i0+RP, i0+2RP)
        const bool laneHasRows = i0 < nRows;
        const int *bndIn = bnd + (size_t)pingpong * 3 * A.boundaryStride;
        int *bndOut = bnd + (size_t)(pingpong ^ 1) * 3 * A.boundaryStride;
        // per packed row: 0xffff where the segment-local chain continues, 0 where a new segment starts
        uint32_t segKeep[RP];
#pragma unroll
        for (int k = 0; k < RP; k++) {
            const uint32_t a = ((i0 + k) % segLen) == 0 ? 0u : 0xffffu;
            const uint32_t b = ((i0 + RP + k) % segLen) == 0 ? 0u : 0xffff0000u;
            segKeep[k] = a | b;
        }
        const uintptr_t addrA = (uintptr_t)(prof + rowOffset + (uint32_t)i0);
        const uintptr_t addrB = addrA + RP;
        const uint32_t shA = (uint32_t)(addrA & 3u) * 8u, shB = (uint32_t)(addrB & 3u) * 8u;

        uint32_t H[RP], E[RP];
#pragma unroll
        for (int k = 0; k < RP; k++) { H[k] = 0; E[k] = 0; }
        uint32_t prevHup = 0;                                  // packed H of the row above, previous column
        uint32_t outL = pack16(SW_DUMMY_LETTER, SW_DUMMY_LETTER), outH = 0, outFt = 0, outFm = 0;
        int tChunk = 0, bH = 0, bFt = 0, bFm = 0;
        const int nSteps = nCols + 63;

        for (int s = 0; s < nSteps; s++) {
            if ((s & 31) == 0) {
                const int c = s + (int)lane;
                tChunk = (c < nCols) ? (int)T[tk.reverse ? (nCols - 1 - c) : c] : SW_DUMMY_LETTER;
                if (hasPrev) {
                    bH = (c < nCols) ? bndIn[c] : 0;
                    bFt = (c < nCols) ? bndIn[A.boundaryStride + c] : 0;
                    bFm = (c < nCols) ? bndIn[2 * A.boundaryStride + c] : 0;
                }
            }
            // stage A takes what the previous lane's stage B produced, stage B what this lane's stage A produced
            uint32_t upL = __shfl_up_sync(0xffffffffu, outL, 1) >> 16;
            uint32_t upH = __shfl_up_sync(0xffffffffu, outH, 1) >> 16;
            uint32_t upFt = __shfl_up_sync(0xffffffffu, outFt, 1) >> 16;
            uint32_t upFm = __shfl_up_sync(0xffffffffu, outFm, 1) >> 16;
            const int a0 = __shfl_sync(0xffffffffu, tChunk, s & 31);
            int h0 = 0, ft0 = 0, fm0 = 0;
            if (hasPrev) {
                h0 = __shfl_sync(0xffffffffu, bH, s & 31);
                ft0 = __shfl_sync(0xffffffffu, bFt, s & 31);
                fm0 = __shfl_sync(0xffffffffu, bFm, s & 31);
            }
            if (lane == 0) { upL = (uint32_t)a0; upH = (uint32_t)h0 & 0xffffu; upFt = (uint32_t)ft0 & 0xffffu; upFm = (uint32_t)fm0 & 0xffffu; }
            const uint32_t inL = upL | (outL << 16);
            const uint32_t inH = upH | (outH << 16);
            const uint32_t inFt = upFt | (outFt << 16);
            const uint32_t inFm = upFm | (outFm << 16);
            const int cA = s - 2 * (int)lane;
            if (cA >= 0 && cA <= nCols && laneHasRows) {        // stage A in [0,nCols), stage B = cA-1 in [-1,nCols)
                const uint32_t letA = inL & 0xffffu, letB = inL >> 16;
                const uint32_t *wa = (const uint32_t *)((addrA + (uintptr_t)letA * stride) & ~(uintptr_t)3);
                const uint32_t *wb = (const uint32_t *)((addrB + (uintptr_t)letB * stride) & ~(uintptr_t)3);
                uint32_t ra[NWA + 1], rb[NWA + 1];
#pragma unroll
                for (int w = 0; w <= NWA; w++) { ra[w] = __ldg(wa + w); rb[w] = __ldg(wb + w); }
                uint32_t sa[NWA], sb[NWA];
#pragma unroll
                for (int w = 0; w < NWA; w++) {
                    sa[w] = __funnelshift_r(ra[w], ra[w + 1], shA);
                    sb[w] = __funnelshift_r(rb[w], rb[w + 1], shB);
                }
                uint32_t diag = prevHup;
                uint32_t ft = inFt, fm = inFm;
                uint32_t colMax = 0;
#pragma unroll
                for (int k = 0; k < RP; k++) {
                    // sign-extended score bytes of both stages -> s16x2 (one PRMT)
                    const uint32_t sel = (uint32_t)(k & 3) | ((uint32_t)((k & 3) | 8) << 4) | ((uint32_t)(4 + (k & 3)) << 8) | ((uint32_t)((4 + (k & 3)) | 8) << 12);
                    const uint32_t sv = prmtSext(sa[k >> 2], sb[k >> 2], sel);
                    const uint32_t e = E[k];
                    fm &= segKeep[k];
                    const uint32_t h = __viaddmax_s16x2(diag, sv, e);
                    const uint32_t h1 = max16x2(h, fm);
                    const uint32_t hs = max16x2(h, ft);
                    diag = H[k];
                    H[k] = hs;
                    // KEYED (8-bit semantics, H < 2^11): the key H*16 + (15-k) orders by score, then by smallest row
                    colMax = max16x2(colMax, KEYED ? (hs * 16u + (uint32_t)((15 - k) * 0x10001)) : hs);
                    const uint32_t open1 = __viaddmax_s16x2_relu(h1, negGo, 0u);
                    E[k] = __viaddmax_s16x2(e, negGe, open1);
                    fm = __viaddmax_s16x2(fm, negGe, open1);
                    const uint32_t openS = __viaddmax_s16x2_relu(hs, negGo, 0u);
                    ft = __viaddmax_s16x2(ft, negGe, openS);
                }
                outL = inL; outH = H[RP - 1]; outFt = ft; outFm = fm;
                prevHup = inH;
                if (KEYED) {
                    // branch-free: (score, column, row) per stage; columns only grow inside a stripe and KEYED tasks have one stripe
                    const int kA = lo16(colMax), kB = hi16(colMax);
                    const int mA = kA >> 4, mB = kB >> 4;
                    const bool upA = mA > bestA && cA < nCols;
                    const bool upB = mB > bestB && cA >= 1;
                    bestA = upA ? mA : bestA; colA = upA ? cA : colA; rowA = upA ? i0 + 15 - (kA & 15) : rowA;
                    bestB = upB ? mB : bestB; colB = upB ? cA - 1 : colB; rowB = upB ? i0 + RP + 15 - (kB & 15) : rowB;
                } else {
                    const int mA = lo16(colMax), mB = hi16(colMax);
                    if (mA >= bestA && mA > 0 && cA < nCols) {
                        if (mA > bestA || cA < colA) {
                            int r = 0;
#pragma unroll
                            for (int k = RP - 1; k >= 0; k--) if (lo16(H[k]) == mA) r = k;
                            bestA = mA; colA = cA; rowA = i0 + r;
                        }
                    }
                    if (mB >= bestB && mB > 0 && cA >= 1) {
                        if (mB > bestB || cA - 1 < colB) {
                            int r = 0;
#pragma unroll
                            for (int k = RP - 1; k >= 0; k--) if (hi16(H[k]) == mB) r = k;
                            bestB = mB; colB = cA - 1; rowB = i0 + RP + r;
                        }
                    }
                }
                if (hasNext && lane == 31 && cA >= 1) {
                    bndOut[cA - 1] = hi16(outH);
                    bndOut[A.boundaryStride + cA - 1] = hi16(outFt);
                    bndOut[2 * A.boundaryStride + cA - 1] = hi16(outFm);
                }
            }
            if ((s & 31) == 31) {
                const int lim = tk.limit != INT_MAX ? tk.limit : SW16_GUARD;
                if (__any_sync(0xffffffffu, max(bestA, bestB) >= lim)) { overflow = true; break; }
            }
        }
        cellCount += (lane == 0) ? (unsigned long long)min(nRows - stripe0, 64 * RP) * (unsigned long long)nCols : 0ull;
        pingpong ^= 1;
        __syncwarp();
    }
    // stage B rows lie below stage A rows: merge the two stages, then the lanes
    int best = bestA, bestCol = colA, bestRow = rowA;
    if (bestB > best || (bestB == best && (colB < bestCol || (colB == bestCol && rowB < bestRow)))) { best = bestB; bestCol = colB; bestRow = rowB; }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const int b2 = __shfl_down_sync(0xffffffffu, best, d);
        const int c2 = __shfl_down_sync(0xffffffffu, bestCol, d);
        const int r2 = __shfl_down_sync(0xffffffffu, bestRow, d);
        if (b2 > best || (b2 == best && (c2 < bestCol || (c2 == bestCol && r2 < bestRow)))) { best = b2; bestCol = c2; bestRow = r2; }
    }
    best = __shfl_sync(0xffffffffu, best, 0);
    const int lim = tk.limit != INT_MAX ? tk.limit : SW16_GUARD;
    if (!overflow && best >= lim) overflow = true;
    if (lane == 0) {
        res.score = best;
        res.col = best > 0 ? bestCol : -1;
        res.row = best > 0 ? bestRow : -1;
        res.overflow = overflow ? (tk.limit != INT_MAX ? 1 : 2) : 0;     // 2: needs the 32-bit pass
    }
}

template <bool PACKED>
__global__ void __launch_bounds__(SW_WARPS * 32) swKernel(SwKernelArgs A) {
    const uint32_t lane = swLane();
    const uint32_t warp = blockIdx.x * SW_WARPS + (threadIdx.x >> 5);
    int *bnd = A.boundary + (size_t)warp * 6 * A.boundaryStride;
    unsigned long long cells = 0;
    for (;;) {
        uint32_t ti = 0;
        if (lane == 0) ti = atomicAdd(A.next, 1u);
        ti = __shfl_sync(0xffffffffu, ti, 0);
        if (ti >= A.nTasks) break;
        const uint32_t id = A.order ? A.order[ti] : ti;
        const SwTask tk = A.tasks[id];
        SwResult r;
        if (PACKED) {
            const int rp = (tk.nRows + 63) / 64;
            // 8-bit semantics on a single stripe: branch-free keyed maxima; everything else: generic tracking
            const bool keyed = tk.limit != INT_MAX && rp <= 16;
#define SD_SW16(N) { if (keyed) swPass16<N, true>(A, tk, bnd, r, cells); else swPass16<N, false>(A, tk, bnd, r, cells); }
            if (rp <= 1) SD_SW16(1)
            else if (rp == 2) SD_SW16(2)
            else if (rp == 3) SD_SW16(3)
            else if (rp == 4) SD_SW16(4)
            else if (rp == 5) SD_SW16(5)
            else if (rp == 6) SD_SW16(6)
            else if (rp == 7) SD_SW16(7)
            else if (rp == 8) SD_SW16(8)
            else if (rp <= 10) SD_SW16(10)
            else if (rp <= 12) SD_SW16(12)
            else SD_SW16(16)
#undef SD_SW16
        } else {
            const int rowsPerLane = (tk.nRows + 31) / 32;
            if (rowsPerLane <= 4) swPass<4>(A, tk, bnd, r, cells);
            else if (rowsPerLane <= 8) swPass<8>(A, tk, bnd, r, cells);
            else if (rowsPerLane <= 16) swPass<16>(A, tk, bnd, r, cells);
            else swPass<32>(A, tk, bnd, r, cells);
        }
        if (lane == 0) A.results[id] = r;
    }
    if (lane == 0 && cells) atomicAdd(A.cells, cells);
}

// score(query row i, letter a) = mat[a][q_i] + compositionBias_i, forward and reversed query
// (SmithWaterman::createQueryProfile StripedSmithWaterman.cpp:163-187, ssw_init :1333-1336)
__global__ void buildSwProfiles(const uint8_t *__restrict__ code, const int8_t *__restrict__ cb,
                                const uint32_t *__restrict__ qOff, uint32_t nQ, uint32_t totalRes,
                                const int8_t *__restrict__ mat /*21x21*/, const uint64_t *__restrict__ profOff,
                                const uint32_t *__restrict__ profStride, int8_t *__restrict__ fwd, int8_t *__restrict__ rev) {
    const uint32_t gp = blockIdx.x * blockDim.x + threadIdx.x;
    if (gp >= totalRes) return;
    // binary search of the owning query
    uint32_t lo = 0, hi = nQ;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (qOff[mid] <= gp) lo = mid; else hi = mid;
    }
    const uint32_t q = lo, i = gp - qOff[q], L = qOff[q + 1] - qOff[q];
    const uint32_t stride = profStride[q];
    const int c = code[gp];
    const int b = cb[gp];
    int8_t *f = fwd + profOff[q], *r = rev + profOff[q];
#pragma unroll
    for (int a = 0; a < AA; a++) {
        const int8_t v = (int8_t)(mat[a * AA + c] + b);
        f[(size_t)a * stride + i] = v;
        r[(size_t)a * stride + (L - 1 - i)] = v;
    }
    // row AA (dummy letter) and all padding stay at the memset value -128
}

// SmithWaterman::ssw_init (StripedSmithWaterman.cpp:1230-1238): int8 composition bias = double(bias -+ 0.5) truncated
__global__ void alignCompositionBias(const float *__restrict__ bias, uint32_t totalRes, int8_t *__restrict__ cb) {
    const uint32_t gp = blockIdx.x * blockDim.x + threadIdx.x;
    if (gp >= totalRes) return;
    const float b = bias[gp];
    const double v = (b < 0.0f) ? __dadd_rn((double)b, -0.5) : __dadd_rn((double)b, 0.5);
    cb[gp] = (int8_t)(int)v;
}

int launchSw(sd_context *ctx, const SwBatch &batch, const SwTask *dTasks, const uint32_t *dOrder, SwResult *dResults,
             uint32_t nTasks, int *dBoundary, uint32_t boundaryStride, uint32_t nWarpsAllocated, uint32_t *dNext,
             unsigned long long *dCells, int packed16) {
    if (nTasks == 0) return SD_OK;
    SwKernelArgs A;
    A.b = batch; A.tasks = dTasks; A.order = dOrder; A.results = dResults; A.nTasks = nTasks; A.next = dNext;
    A.boundary = dBoundary; A.boundaryStride = boundaryStride; A.cells = dCells; A.packed16 = packed16;
    SD_CUDA(cudaMemsetAsync(dNext, 0, sizeof(uint32_t), ctx->stream));
    uint32_t nCta = std::min<uint32_t>(nWarpsAllocated / SW_WARPS, (nTasks + SW_WARPS - 1) / SW_WARPS);
    if (nCta == 0) nCta = 1;
    if (packed16) swKernel<true><<<nCta, SW_WARPS * 32, 0, ctx->stream>>>(A);
    else swKernel<false><<<nCta, SW_WARPS * 32, 0, ctx->stream>>>(A);
    ctx->launches++;
    SD_CUDA(cudaGetLastError());
    return SD_OK;
}

int launchSwProfiles(sd_context *ctx, const uint8_t *code, const int8_t *cb, const uint32_t *qOff, uint32_t nQ,
                     uint32_t totalRes, const int8_t *mat, const uint64_t *profOff, const uint32_t *profStride,
                     int8_t *fwd, int8_t *rev) {
    if (totalRes == 0) return SD_OK;
    buildSwProfiles<<<(totalRes + 127) / 128, 128, 0, ctx->stream>>>(code, cb, qOff, nQ, totalRes, mat, profOff, profStride, fwd, rev);
    ctx->launches++;
    SD_CUDA(cudaGetLastError());
    return SD_OK;
}

int launchAlignCompositionBias(sd_context *ctx, const float *bias, uint32_t totalRes, int8_t *cb) {
    if (totalRes == 0) return SD_OK;
    alignCompositionBias<<<(totalRes + 255) / 256, 256, 0, ctx->stream>>>(bias, totalRes, cb);
    ctx->launches++;
    SD_CUDA(cudaGetLastError());
    return SD_OK;
}

uint32_t swWarpsPerCta() { return SW_WARPS; }

}  // namespace sd
