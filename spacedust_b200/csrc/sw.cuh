// Device-side interface of the Smith-Waterman stage (sw.cu), used by align.cu.
#pragma once
#include <cstdint>

namespace sd {

// One score/coordinate pass of the striped-semantics Smith-Waterman (see sw.cu).
struct SwTask {
    uint32_t q;          // query index in the batch
    uint32_t t;          // target index in the index
    int32_t nRows;       // query rows used: Lq (forward) or qEnd+1 (reverse)
    int32_t nCols;       // target columns used: Lt (forward) or tEnd+1 (reverse)
    int32_t limit;       // byte mode: first H >= limit means 8-bit overflow; word mode: INT_MAX
    int32_t width;       // 32 (sw_sse2_byte layout) or 16 (sw_sse2_word layout)
    int32_t reverse;     // 0 forward, 1 reverse (both sequences reversed on their prefixes)
    int32_t pad;
};

struct SwResult {
    int32_t score;       // maximum H (meaningless when overflow is set)
    int32_t col;         // first column (pass orientation) reaching it
    int32_t row;         // smallest row (pass orientation) holding it in that column
    int32_t overflow;    // byte mode only
};

struct SwBatch {
    // query side (per batch)
    const int8_t *profFwd;      // [q] 21 rows of stride profStride[q]: score(query row i, letter a) at a*stride + i
    const int8_t *profRev;      // same for the reversed query
    const uint64_t *profOff;    // start of query q's block
    const uint32_t *profStride; // row stride (bytes), multiple of 4
    const uint32_t *qLen;
    // forward 8-bit pass only: the same scores laid out per wavefront stage (swStageLayout), every stage's rows starting on a
    // 16-byte boundary, so a stage reads them with one or two aligned 128-bit loads (no funnel shifts, 3x fewer L1
    // wavefronts than 32-bit loads: every lane reads its own line); stride 0 = not built for this query
    const int8_t *profStg;
    const uint64_t *stgOff;
    const uint32_t *stgStride;
    // target side
    const uint8_t *tRes;        // UNMASKED residue codes
    const uint32_t *tOff;
    int gapOpen, gapExtend;
};

// Scheduling classes of the packed pass: (lanes per pair, packed rows per stage).  Short queries use 8 lanes per pair
// (4 pairs per warp, 16 wavefront stages), so every lane still owns enough rows to amortise the per-step work
// (shuffles, profile fetch, maxima) and the pipeline fill is 15 columns instead of 63.  The stage height follows the
// segment length of the 8-bit striped layout, segLen = ceil(nRows/32): 8 lanes -> RP = 2 segLen, 16 lanes -> RP = segLen,
// so that in the forward 8-bit pass every segment starts at a compile-time row of a stage (SEGMODE in sw.cu).
__host__ __device__ inline int swClassOf(int nRows) {
    if (nRows <= 320) { int rp = 2 * ((nRows + 31) / 32); if (rp < 2) rp = 2; return rp / 2 - 1; }                       // 0..9  : 8 lanes, RP 2,4..20
    if (nRows <= 640) { const int rp = (nRows + 31) / 32; return 10 + (rp - 11); }                                       // 10..19: 16 lanes, RP 11..20
    if (nRows <= 1280) { int rp = (nRows + 63) / 64; rp = (rp + 1) & ~1; return 20 + (rp - 12) / 2; }                   // 20..24: 32 lanes, RP 12..20
    return 25;                                                                                                         // stripes of 1280 rows
}
__host__ __device__ inline int swClassTasksPerRound(int cls) { return cls < 10 ? 4 : (cls < 20 ? 2 : 1); }
constexpr int SW_NCLASS = 26;

// Stage layout of the staged forward profile for a query of nRows rows: rows per stage (0 = no staged profile: 32-lane
// classes), padded rows per stage, number of stages.  Row r of letter a sits at a*stride + (r / rp) * rpPad + r % rp.
__host__ __device__ inline void swStageLayout(int nRows, int &rp, int &rpPad, int &nStages) {
    const int cls = swClassOf(nRows);
    if (cls < 10) { rp = 2 * (cls + 1); nStages = 16; }
    else if (cls < 20) { rp = cls + 1; nStages = 32; }
    else { rp = 0; nStages = 0; }
    rpPad = (rp + 15) & ~15;      // 16-byte blocks: a stage fetches its rows with one or two 128-bit loads
}

}  // namespace sd
