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
    // target side
    const uint8_t *tRes;        // UNMASKED residue codes
    const uint32_t *tOff;
    int gapOpen, gapExtend;
};

}  // namespace sd
