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
#include <cub/cub.cuh>

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
    const uint32_t *nRounds;   // packed pass: number of rounds in order[]
    uint32_t zero;             // always 0; a kernel parameter so that ptxas cannot re-materialise the constant per cell
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
                    // Ft >= Fm and go >= ge: (Fm - go) <= (Ft - ge), so the gap opened from H1 also serves the exact chain
                    ft = __viaddmax_s32(ft, -ge, open1);
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

constexpr int SW_KEYMUL = 33;      // column-maximum key = H * 33 + (31 - k): not a shift, so it is an IMAD (FMA pipe), not an LEA

// SEGMODE: where the reference's segment starts (Fm resets) fall inside a stage
//   0  anywhere (per-row runtime mask)      1  rows 0 and RP/2 of both stages (segLen = RP/2)      2  row 0 of both stages (segLen = RP)
template <int LANES, int RP, bool KEYED, int SEGMODE>
__device__ void swGroupPass(const SwKernelArgs &A, uint32_t round, int *bnd, unsigned long long &cellCount) {
    constexpr int NWA = (RP + 3) / 4;          // aligned words holding RP bytes
    constexpr uint32_t BIAS = (SEGMODE != 0) ? 32u : 0u;     // see the note at biasP
    constexpr bool STRIPED = (LANES == 32);    // only whole-warp pairs may need more than one stripe
    const uint32_t lane = swLane();
    const uint32_t gl = lane % LANES, grp = lane / LANES;
    const uint32_t gmask = (LANES == 32) ? 0xffffffffu : (((1u << LANES) - 1u) << (grp * LANES));
    const uint32_t id = A.order[(size_t)round * 4 + grp];
    const bool have = id != 0xffffffffu;
    SwTask tk;
    if (have) tk = A.tasks[id];
    else { tk.q = 0; tk.t = 0; tk.nRows = 0; tk.nCols = 0; tk.limit = INT_MAX; tk.width = 32; tk.reverse = 0; tk.pad = 0; }
    const int nRows = tk.nRows, nCols = tk.nCols;
    const int segLen = have ? (nRows + tk.width - 1) / tk.width : 1;
    const uint32_t negGo = pack16(-A.b.gapOpen, -A.b.gapOpen), negGe = pack16(-A.b.gapExtend, -A.b.gapExtend);
    const uint32_t Lq = have ? A.b.qLen[tk.q] : 0u;
    const uint32_t stride = have ? A.b.profStride[tk.q] : 0u;
    const int8_t *prof = (tk.reverse ? A.b.profRev : A.b.profFwd) + (have ? A.b.profOff[tk.q] : 0ull);
    const uint32_t rowOffset = tk.reverse ? (Lq - (uint32_t)nRows) : 0u;
    const uint8_t *T = A.b.tRes + (have ? A.b.tOff[tk.t] : 0u);
    const int lim = tk.limit != INT_MAX ? tk.limit : SW16_GUARD;

    int bestA = 0, bestB = 0, colA = INT_MAX, colB = INT_MAX, rowA = INT_MAX, rowB = INT_MAX;
    int thrA = ((int)BIAS + 1) * SW_KEYMUL, thrB = thrA;     // KEYED: smallest key that beats the current best score
    bool overflow = false;
    int pingpong = 0;
    const uint32_t zero = A.zero;               // opaque 0 (see SwKernelArgs)
    // SEGMODE != 0: all H/E/F values carry +BIAS in both halves.  The exact vertical chain is then not carried row by row:
    // inside a segment that starts at row kb with the exact F value X, F_true(k) = max(F_seg(k), X - (k-kb)*ge), and
    // X - d*ge for both halves is ONE 32-bit subtraction of a constant (the low half never borrows: X_lo >= BIAS > d), which
    // ptxas can place on the FMA pipe (IMAD) instead of one more DPX instruction per cell on the saturated ALU pipe.
    const uint32_t biasP = zero + BIAS * 0x10001u;
    const uint32_t one = zero + 1u;             // opaque 1: X * one + c stays an IMAD (FMA pipe)
    const uint32_t goP = (uint32_t)A.b.gapOpen * 0x10001u, geP = (uint32_t)A.b.gapExtend * 0x10001u;
    // every group runs the longest group's step count (shuffles are warp-wide)
    const int mySteps = have ? nCols + 2 * LANES - 1 : 0;
    const int warpSteps = (int)__reduce_max_sync(0xffffffffu, (unsigned)mySteps);
    const int nStripes = STRIPED ? (int)__reduce_max_sync(0xffffffffu, (unsigned)((nRows + 64 * RP - 1) / (64 * RP))) : 1;

    for (int stripe = 0; stripe < nStripes; stripe++) {
        const int stripe0 = stripe * 64 * RP;
        const bool hasPrev = STRIPED && stripe0 > 0;
        const bool hasNext = STRIPED && stripe0 + 64 * RP < nRows;
        const int i0 = stripe0 + (int)gl * 2 * RP;             // stage A rows [i0, i0+RP), stage B rows [i0+RP, i0+2RP)
        const bool laneHasRows = have && i0 < nRows && !overflow;
        const int *bndIn = bnd + (size_t)pingpong * 3 * A.boundaryStride;
        int *bndOut = bnd + (size_t)(pingpong ^ 1) * 3 * A.boundaryStride;
        uint32_t segKeep[SEGMODE == 0 ? RP : 1];
        if (SEGMODE == 0) {
#pragma unroll
            for (int k = 0; k < RP; k++) {
                const uint32_t a = ((i0 + k) % segLen) == 0 ? 0u : 0xffffu;
                const uint32_t b = ((i0 + RP + k) % segLen) == 0 ? 0u : 0xffff0000u;
                segKeep[SEGMODE == 0 ? k : 0] = a | b;
            }
        }
        // SEGMODE != 0 (forward 8-bit pass): staged profile, stage st starts at st * RPPAD on a 16-byte boundary
        constexpr int RPPAD = (RP + 15) & ~15;
        const uint32_t strideU = (SEGMODE != 0 && have) ? A.b.stgStride[tk.q] : stride;
        const uintptr_t addrA = (SEGMODE != 0) ? (uintptr_t)(A.b.profStg + (have ? A.b.stgOff[tk.q] : 0ull) + (uint32_t)(2 * gl) * RPPAD)
                                               : (uintptr_t)(prof + rowOffset + (uint32_t)i0);
        const uintptr_t addrB = addrA + ((SEGMODE != 0) ? RPPAD : RP);
        const uint32_t shA = (uint32_t)(addrA & 3u) * 8u, shB = (uint32_t)(addrB & 3u) * 8u;

        uint32_t H[RP], E[RP];
#pragma unroll
        for (int k = 0; k < RP; k++) { H[k] = biasP; E[k] = biasP; }
        uint32_t prevHup = biasP;
        uint32_t outL = pack16(SW_DUMMY_LETTER, SW_DUMMY_LETTER), outH = biasP, outFt = biasP, outFm = biasP;
        int tChunk = SW_DUMMY_LETTER, bH = 0, bFt = 0, bFm = 0;

        for (int s = 0; s < warpSteps; s++) {
            if ((s % LANES) == 0) {
                const int c = s + (int)gl;
                tChunk = (have && c < nCols) ? (int)T[tk.reverse ? (nCols - 1 - c) : c] : SW_DUMMY_LETTER;
                if (hasPrev) {
                    bH = (c < nCols) ? bndIn[c] : 0;
                    bFt = (c < nCols) ? bndIn[A.boundaryStride + c] : 0;
                    bFm = (c < nCols) ? bndIn[2 * A.boundaryStride + c] : 0;
                }
            }
            uint32_t upL = __shfl_up_sync(0xffffffffu, outL, 1, LANES) >> 16;
            uint32_t upH = __shfl_up_sync(0xffffffffu, outH, 1, LANES) >> 16;
            uint32_t upFt = __shfl_up_sync(0xffffffffu, outFt, 1, LANES) >> 16;
            uint32_t upFm = __shfl_up_sync(0xffffffffu, outFm, 1, LANES) >> 16;
            const int a0 = __shfl_sync(0xffffffffu, tChunk, s % LANES, LANES);
            int h0 = 0, ft0 = 0, fm0 = 0;
            if (STRIPED && nStripes > 1) {
                h0 = __shfl_sync(0xffffffffu, bH, s % LANES, LANES);
                ft0 = __shfl_sync(0xffffffffu, bFt, s % LANES, LANES);
                fm0 = __shfl_sync(0xffffffffu, bFm, s % LANES, LANES);
                if (!hasPrev) { h0 = 0; ft0 = 0; fm0 = 0; }
            }
            if (gl == 0) {
                upL = (uint32_t)a0;
                upH = (uint32_t)(h0 + (int)BIAS) & 0xffffu; upFt = (uint32_t)(ft0 + (int)BIAS) & 0xffffu; upFm = (uint32_t)(fm0 + (int)BIAS) & 0xffffu;
            }
            const uint32_t inL = upL | (outL << 16);
            const uint32_t inH = upH | (outH << 16);
            const uint32_t inFt = upFt | (outFt << 16);
            const uint32_t inFm = upFm | (outFm << 16);
            const int cA = s - 2 * (int)gl;
            if (cA >= 0 && cA <= nCols && laneHasRows && !overflow) {   // stage A column cA in [0,nCols), stage B column cA-1
                const uint32_t letA = inL & 0xffffu, letB = inL >> 16;
                uint32_t sa[NWA], sb[NWA];
                if (SEGMODE != 0) {
                    const uint4 *wa = (const uint4 *)(addrA + (uintptr_t)letA * strideU);
                    const uint4 *wb = (const uint4 *)(addrB + (uintptr_t)letB * strideU);
#pragma unroll
                    for (int w = 0; w < RPPAD / 16; w++) {
                        const uint4 va = __ldg(wa + w), vb = __ldg(wb + w);
                        if (4 * w + 0 < NWA) { sa[4 * w + 0] = va.x; sb[4 * w + 0] = vb.x; }
                        if (4 * w + 1 < NWA) { sa[4 * w + 1] = va.y; sb[4 * w + 1] = vb.y; }
                        if (4 * w + 2 < NWA) { sa[4 * w + 2] = va.z; sb[4 * w + 2] = vb.z; }
                        if (4 * w + 3 < NWA) { sa[4 * w + 3] = va.w; sb[4 * w + 3] = vb.w; }
                    }
                } else {
                    const uint32_t *wa = (const uint32_t *)((addrA + (uintptr_t)letA * stride) & ~(uintptr_t)3);
                    const uint32_t *wb = (const uint32_t *)((addrB + (uintptr_t)letB * stride) & ~(uintptr_t)3);
                    uint32_t ra[NWA + 1], rb[NWA + 1];
#pragma unroll
                    for (int w = 0; w <= NWA; w++) { ra[w] = __ldg(wa + w); rb[w] = __ldg(wb + w); }
#pragma unroll
                    for (int w = 0; w < NWA; w++) {
                        sa[w] = __funnelshift_r(ra[w], ra[w + 1], shA);
                        sb[w] = __funnelshift_r(rb[w], rb[w + 1], shB);
                    }
                }
                uint32_t diag = prevHup;
                uint32_t ft = inFt, fm = inFm;
                uint32_t colMax = 0;
                if (SEGMODE != 0) {
                    // segment-aligned stages (forward 8-bit pass, ge == 1): see the BIAS note above
                    uint32_t X = inFt;           // exact F entering the stage's first row (= a segment start)
                    int kb = 0;
                    fm = biasP;
#pragma unroll
                    for (int k = 0; k < RP; k++) {
                        const uint32_t sel = (uint32_t)(k & 3) | ((uint32_t)((k & 3) | 8) << 4) | ((uint32_t)(4 + (k & 3)) << 8) | ((uint32_t)((4 + (k & 3)) | 8) << 12);
                        const uint32_t sv = prmtSext(sa[k >> 2], sb[k >> 2], sel);
                        const uint32_t e = E[k];
                        const bool segStart = (k == 0) || (SEGMODE == 1 && k == RP / 2);
                        if (segStart && k != 0) {      // second segment of the stage: its X is the exact F entering this row
                            X = max16x2(fm, X * one - (uint32_t)(k - kb) * 0x10001u);
                            kb = k;
                            fm = biasP;
                        }
                        const int d = k - kb;
                        // X - d in both halves with one 32-bit subtraction: X_lo >= BIAS > d, so the low half never borrows
                        const uint32_t xk = d == 0 ? X : X * one - (uint32_t)d * 0x10001u;
                        const uint32_t h = __viaddmax_s16x2(diag, sv, e);
                        // the gap-open term without Fm: max(H1 - go, 0) = max(h - go, 0, Fm - go), so the Fm chain
                        // Fm' = max(Fm - ge, h - go, 0) is ONE dependent instruction per row (Fm - go <= Fm - ge)
                        const uint32_t hg = __viaddmax_s16x2(h, negGo, biasP);
                        const uint32_t fmg = fm * one - goP;          // Fm - go, E - ge: borrow-free 32-bit subtractions (IMAD)
                        const uint32_t eg = e * one - geP;
                        const uint32_t hs = segStart ? max16x2(h, xk) : __vimax3_s16x2(h, fm, xk);   // max(H1, exact F)
                        diag = H[k];
                        H[k] = hs;
                        colMax = max16x2(colMax, hs * (uint32_t)SW_KEYMUL + (uint32_t)((31 - k) * 0x10001));
                        E[k] = __vimax3_s16x2(eg, hg, fmg);
                        fm = __viaddmax_s16x2(fm, negGe, hg);
                    }
                    ft = max16x2(fm, X * one - (uint32_t)(RP - kb) * 0x10001u);     // exact F leaving the stage
                } else {
#pragma unroll
                for (int k = 0; k < RP; k++) {
                    const uint32_t sel = (uint32_t)(k & 3) | ((uint32_t)((k & 3) | 8) << 4) | ((uint32_t)(4 + (k & 3)) << 8) | ((uint32_t)((4 + (k & 3)) | 8) << 12);
                    const uint32_t sv = prmtSext(sa[k >> 2], sb[k >> 2], sel);
                    const uint32_t e = E[k];
                    fm &= segKeep[SEGMODE == 0 ? k : 0];
                    const uint32_t h = __viaddmax_s16x2(diag, sv, e);
                    const uint32_t h1 = max16x2(h, fm);
                    const uint32_t hs = max16x2(h, ft);
                    diag = H[k];
                    H[k] = hs;
                    // KEYED (8-bit semantics, H < 992): key H*33 + (31-k) orders by score, then by smallest row
                    colMax = max16x2(colMax, KEYED ? (hs * (uint32_t)SW_KEYMUL + (uint32_t)((31 - k) * 0x10001)) : hs);
                    const uint32_t open1 = __viaddmax_s16x2_relu(h1, negGo, zero);
                    E[k] = __viaddmax_s16x2(e, negGe, open1);
                    fm = __viaddmax_s16x2(fm, negGe, open1);
                    // Ft >= Fm and go >= ge: (Fm - go) <= (Ft - ge), so the gap opened from H1 also serves the exact chain
                    ft = __viaddmax_s16x2(ft, negGe, open1);
                }
                }
                outL = inL; outH = H[RP - 1]; outFt = ft; outFm = fm;
                prevHup = inH;
                if (KEYED) {
                    const int kA = lo16(colMax), kB = hi16(colMax);
                    const bool upA = kA >= thrA && cA < nCols;
                    const bool upB = kB >= thrB && cA >= 1;
                    if (upA | upB) {            // rare after the first columns: decode score and row from the key
                        const int mA = kA / SW_KEYMUL, mB = kB / SW_KEYMUL;
                        if (upA) { bestA = mA - (int)BIAS; colA = cA; rowA = i0 + 31 - (kA - mA * SW_KEYMUL); thrA = (mA + 1) * SW_KEYMUL; }
                        if (upB) { bestB = mB - (int)BIAS; colB = cA - 1; rowB = i0 + RP + 31 - (kB - mB * SW_KEYMUL); thrB = (mB + 1) * SW_KEYMUL; }
                    }
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
                if (hasNext && gl == LANES - 1 && cA >= 1) {
                    bndOut[cA - 1] = hi16(outH);
                    bndOut[A.boundaryStride + cA - 1] = hi16(outFt);
                    bndOut[2 * A.boundaryStride + cA - 1] = hi16(outFm);
                }
            }
            if ((s & 31) == 31) {
                const uint32_t hit = __ballot_sync(0xffffffffu, max(bestA, bestB) >= lim);
                if (hit & gmask) overflow = true;      // this pair is decided; its lanes idle until the round ends
                if (__all_sync(0xffffffffu, overflow || !have)) break;
            }
        }
        cellCount += (have && gl == 0 && stripe0 < nRows) ? (unsigned long long)min(nRows - stripe0, 64 * RP) * (unsigned long long)nCols : 0ull;
        pingpong ^= 1;
        __syncwarp();
    }
    int best = bestA, bestCol = colA, bestRow = rowA;
    if (bestB > best || (bestB == best && (colB < bestCol || (colB == bestCol && rowB < bestRow)))) { best = bestB; bestCol = colB; bestRow = rowB; }
#pragma unroll
    for (int d = LANES / 2; d > 0; d >>= 1) {
        const int b2 = __shfl_down_sync(0xffffffffu, best, d, LANES);
        const int c2 = __shfl_down_sync(0xffffffffu, bestCol, d, LANES);
        const int r2 = __shfl_down_sync(0xffffffffu, bestRow, d, LANES);
        if (b2 > best || (b2 == best && (c2 < bestCol || (c2 == bestCol && r2 < bestRow)))) { best = b2; bestCol = c2; bestRow = r2; }
    }
    best = __shfl_sync(0xffffffffu, best, 0, LANES);
    if (!overflow && best >= lim) overflow = true;
    if (have && gl == 0) {
        SwResult res;
        res.score = best;
        res.col = best > 0 ? bestCol : -1;
        res.row = best > 0 ? bestRow : -1;
        res.overflow = overflow ? (tk.limit != INT_MAX ? 1 : 2) : 0;     // 2: needs the 32-bit pass
        A.results[id] = res;
    }
}

// packed pass: one round = up to four pairs of the same scheduling class (order[] is laid out in rounds of 4 slots)
template <bool KEYED>
#ifndef SD_SW_LB
#define SD_SW_LB 5      // resident CTAs per SM the packed kernel is compiled for (align.cu sizes the grid with the same number)
#endif
__global__ void __launch_bounds__(SW_WARPS * 32, SD_SW_LB) swKernel16(SwKernelArgs A) {
    const uint32_t lane = swLane();
    const uint32_t warp = blockIdx.x * SW_WARPS + (threadIdx.x >> 5);
    int *bnd = A.boundary + (size_t)warp * 6 * A.boundaryStride;
    unsigned long long cells = 0;
    const uint32_t nRounds = *A.nRounds;
    for (;;) {
        uint32_t round = 0;
        if (lane == 0) round = atomicAdd(A.next, 1u);
        round = __shfl_sync(0xffffffffu, round, 0);
        if (round >= nRounds) break;
        const uint32_t id0 = A.order[(size_t)round * 4];      // slot 0 of a round is always a real task
        const int cls = swClassOf(A.tasks[id0].nRows);
        // forward 8-bit pass (KEYED): width 32, so segLen = RP/2 (8 lanes) or RP (16 lanes) for every pair of the round and
        // the segment starts sit at compile-time rows; anything else (checked, warp-uniform) takes the per-row mask
        bool segAligned = false;
        if (KEYED && cls < 20 && A.b.gapExtend == 1 && A.b.gapOpen >= 1 && A.b.gapOpen < 32) {
            const uint32_t idg = A.order[(size_t)round * 4 + (cls < 10 ? lane / 8 : lane / 16)];
            bool ok = true;
            if (idg != 0xffffffffu) {
                const SwTask tg = A.tasks[idg];
                const int sl = (tg.nRows + tg.width - 1) / tg.width;
                const int rp = cls < 10 ? 2 * (cls + 1) : cls + 1;
                ok = ((cls < 10) ? (2 * sl == rp) : (sl == rp)) && tg.reverse == 0 && tg.nRows == (int)A.b.qLen[tg.q] && A.b.stgStride[tg.q] != 0;
            }
            segAligned = __all_sync(0xffffffffu, ok);
        }
#define SD_G(L, N, M) if (KEYED && M != 0 && segAligned) swGroupPass<L, N, KEYED, KEYED ? M : 0>(A, round, bnd, cells); \
                      else swGroupPass<L, N, KEYED, 0>(A, round, bnd, cells); break;
        switch (cls) {
            case 0: SD_G(8, 2, 1) case 1: SD_G(8, 4, 1) case 2: SD_G(8, 6, 1) case 3: SD_G(8, 8, 1) case 4: SD_G(8, 10, 1)
            case 5: SD_G(8, 12, 1) case 6: SD_G(8, 14, 1) case 7: SD_G(8, 16, 1) case 8: SD_G(8, 18, 1) case 9: SD_G(8, 20, 1)
            case 10: SD_G(16, 11, 2) case 11: SD_G(16, 12, 2) case 12: SD_G(16, 13, 2) case 13: SD_G(16, 14, 2) case 14: SD_G(16, 15, 2)
            case 15: SD_G(16, 16, 2) case 16: SD_G(16, 17, 2) case 17: SD_G(16, 18, 2) case 18: SD_G(16, 19, 2) case 19: SD_G(16, 20, 2)
            case 20: SD_G(32, 12, 0) case 21: SD_G(32, 14, 0) case 22: SD_G(32, 16, 0) case 23: SD_G(32, 18, 0)
            default: SD_G(32, 20, 0)
        }
#undef SD_G
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) cells += __shfl_down_sync(0xffffffffu, cells, d);    // group leaders counted their pairs
    if (lane == 0 && cells) atomicAdd(A.cells, cells);
}

// ---- scheduling: sort tasks by (class desc, columns), lay them out in class-pure rounds of 4 slots
__global__ void swScheduleKeys(const SwTask *tasks, uint32_t n, uint32_t *keys, uint32_t *vals, uint32_t *classCount) {
    __shared__ uint32_t sCount[SW_NCLASS];      // per-block histogram: one global atomic per class and block
    for (int c = threadIdx.x; c < SW_NCLASS; c += blockDim.x) sCount[c] = 0;
    __syncthreads();
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const int cls = swClassOf(tasks[i].nRows);
        keys[i] = ((uint32_t)(SW_NCLASS - 1 - cls) << 16) | (uint32_t)min(tasks[i].nCols, 65535);
        vals[i] = i;
        atomicAdd(&sCount[SW_NCLASS - 1 - cls], 1u);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < SW_NCLASS; c += blockDim.x)
        if (sCount[c]) atomicAdd(&classCount[c], sCount[c]);
}

// classInfo: [0..20] counts (in), -> [0..20] first sorted position, [32..52] first round, [64] number of rounds
__global__ void swScheduleBases(uint32_t *classInfo) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint32_t pos = 0, round = 0;
    for (int cc = 0; cc < SW_NCLASS; cc++) {
        const uint32_t cnt = classInfo[cc];
        const uint32_t tpr = (uint32_t)swClassTasksPerRound(SW_NCLASS - 1 - cc);
        classInfo[cc] = pos;
        classInfo[32 + cc] = round;
        pos += cnt;
        round += (cnt + tpr - 1) / tpr;
    }
    classInfo[64] = round;
}

__global__ void swScheduleFill(const uint32_t *sortedKeys, const uint32_t *sortedVals, uint32_t n, const uint32_t *classInfo, uint32_t *order) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t cc = sortedKeys[p] >> 16;
    const uint32_t r = p - classInfo[cc];
    const uint32_t tpr = (uint32_t)swClassTasksPerRound(SW_NCLASS - 1 - (int)cc);
    order[(size_t)(classInfo[32 + cc] + r / tpr) * 4 + (r % tpr)] = sortedVals[p];
}

// 32-bit pass for the few pairs whose scores approach 2^15 (the packed pass flags them with overflow = 2)
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
        const int rowsPerLane = (tk.nRows + 31) / 32;
        if (rowsPerLane <= 4) swPass<4>(A, tk, bnd, r, cells);
        else if (rowsPerLane <= 8) swPass<8>(A, tk, bnd, r, cells);
        else if (rowsPerLane <= 16) swPass<16>(A, tk, bnd, r, cells);
        else swPass<32>(A, tk, bnd, r, cells);
        if (lane == 0) A.results[id] = r;
    }
    if (lane == 0 && cells) atomicAdd(A.cells, cells);
}

// score(query row i, letter a) = mat[a][q_i] + compositionBias_i, forward and reversed query
// (SmithWaterman::createQueryProfile StripedSmithWaterman.cpp:163-187, ssw_init :1333-1336)
__global__ void buildSwProfiles(const uint8_t *__restrict__ code, const int8_t *__restrict__ cb,
                                const uint32_t *__restrict__ qOff, uint32_t nQ, uint32_t totalRes,
                                const int8_t *__restrict__ mat /*21x21*/, const uint64_t *__restrict__ profOff,
                                const uint32_t *__restrict__ profStride, int8_t *__restrict__ fwd, int8_t *__restrict__ rev,
                                const uint64_t *__restrict__ stgOff, const uint32_t *__restrict__ stgStride, int8_t *__restrict__ stg) {
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
    const uint32_t sStride = stgStride[q];
    int rp, rpPad, nStages;
    swStageLayout((int)L, rp, rpPad, nStages);
    int8_t *g = stg + stgOff[q] + (sStride ? (i / (uint32_t)rp) * (uint32_t)rpPad + i % (uint32_t)rp : 0u);
#pragma unroll
    for (int a = 0; a < AA; a++) {
        const int8_t v = (int8_t)(mat[a * AA + c] + b);
        f[(size_t)a * stride + i] = v;
        r[(size_t)a * stride + (L - 1 - i)] = v;
        if (sStride) g[(size_t)a * sStride] = v;
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

int launchSw(sd_context *ctx, const SwBatch &batch, const SwTask *dTasks, SwResult *dResults, uint32_t nTasks, int *dBoundary,
             uint32_t boundaryStride, uint32_t nWarpsAllocated, uint32_t *dNext, unsigned long long *dCells, int packed16, int keyed) {
    if (nTasks == 0) return SD_OK;
    cudaStream_t st = ctx->stream;
    SwKernelArgs A;
    A.b = batch; A.tasks = dTasks; A.order = nullptr; A.results = dResults; A.nTasks = nTasks; A.next = dNext;
    A.boundary = dBoundary; A.boundaryStride = boundaryStride; A.cells = dCells; A.packed16 = packed16; A.nRounds = nullptr; A.zero = 0;
    SD_CUDA(cudaMemsetAsync(dNext, 0, sizeof(uint32_t), st));
    uint32_t nCta = std::min<uint32_t>(nWarpsAllocated / SW_WARPS, (nTasks + SW_WARPS - 1) / SW_WARPS);
    if (nCta == 0) nCta = 1;
    if (!packed16) {
        swKernel<<<nCta, SW_WARPS * 32, 0, st>>>(A);
        ctx->launches++;
        SD_CUDA(cudaGetLastError());
        return SD_OK;
    }
    // schedule: (class desc, columns) order, class-pure rounds of four slots
    ScratchBuf<uint32_t> keysA(ctx, SL_SW_KEYA), keysB(ctx, SL_SW_KEYB), valsA(ctx, SL_SW_VALA), valsB(ctx, SL_SW_VALB), order(ctx, SL_SW_ORDER);
    ScratchBuf<uint8_t> tmp(ctx, SL_SW_TMP);
    DevBuf<uint32_t> info;
    int rc;
    const size_t orderSlots = (size_t)nTasks * 4 + 4 * SW_NCLASS + 16;
    if ((rc = keysA.alloc(nTasks)) || (rc = keysB.alloc(nTasks)) || (rc = valsA.alloc(nTasks)) || (rc = valsB.alloc(nTasks)) ||
        (rc = order.alloc(orderSlots)) || (rc = info.alloc(128)))
        return rc;
    SD_CUDA(cudaMemsetAsync(info.p, 0, 128 * sizeof(uint32_t), st));
    SD_CUDA(cudaMemsetAsync(order.p, 0xff, orderSlots * sizeof(uint32_t), st));
    swScheduleKeys<<<(nTasks + 255) / 256, 256, 0, st>>>(dTasks, nTasks, keysA.p, valsA.p, info.p);
    size_t tmpBytes = 0;
    SD_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, keysA.p, keysB.p, valsA.p, valsB.p, (int64_t)nTasks, 0, 21, st));
    if ((rc = tmp.alloc(tmpBytes))) return rc;
    SD_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmpBytes, keysA.p, keysB.p, valsA.p, valsB.p, (int64_t)nTasks, 0, 21, st));
    swScheduleBases<<<1, 32, 0, st>>>(info.p);
    swScheduleFill<<<(nTasks + 255) / 256, 256, 0, st>>>(keysB.p, valsB.p, nTasks, info.p, order.p);
    ctx->launches += 3;
    SD_CUDA(cudaGetLastError());
    A.order = order.p;
    A.nRounds = info.p + 64;
    nCta = std::max<uint32_t>(1u, nWarpsAllocated / SW_WARPS);
    if (keyed) swKernel16<true><<<nCta, SW_WARPS * 32, 0, st>>>(A);
    else swKernel16<false><<<nCta, SW_WARPS * 32, 0, st>>>(A);
    ctx->launches++;
    SD_CUDA(cudaGetLastError());
    SD_CUDA(cudaStreamSynchronize(st));      // info (DevBuf) is released when this returns
    return SD_OK;
}

int launchSwProfiles(sd_context *ctx, const uint8_t *code, const int8_t *cb, const uint32_t *qOff, uint32_t nQ,
                     uint32_t totalRes, const int8_t *mat, const uint64_t *profOff, const uint32_t *profStride,
                     int8_t *fwd, int8_t *rev, const uint64_t *stgOff, const uint32_t *stgStride, int8_t *stg) {
    if (totalRes == 0) return SD_OK;
    buildSwProfiles<<<(totalRes + 127) / 128, 128, 0, ctx->stream>>>(code, cb, qOff, nQ, totalRes, mat, profOff, profStride, fwd, rev,
                                                                      stgOff, stgStride, stg);
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
