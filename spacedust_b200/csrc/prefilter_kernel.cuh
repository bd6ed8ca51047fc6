// The prefilter kernel (included by prefilter.cu inside namespace sd).
//
// One CTA owns one query at a time (queries handed out longest first through an atomic counter, two CTAs per SM).
// The reference's result depends on the ORDER of the index hits of a target (double-diagonal rule,
// CacheFriendlyOperations.cpp:196-207) and, when --max-seqs cuts, on the arrival order of the kept candidates
// (radix bin, then stream order).  The reference makes its per-target table cache resident by binning the hit stream
// by target id (CacheFriendlyOperations::hashIndexEntry); the same idea, laid out for an SM:
//
//   P1  produce   every warp takes a contiguous slice of the query's k-mer positions and walks it in stream order:
//                 similar k-mers (KmerGenerator order, rows of (first 3-mer neighbour a, count of second-row
//                 neighbours) flattened over positions so that every probe step has 32 busy lanes) -> occupancy
//                 bitmap {bits, rank} (16 MB, L2 resident) -> bucket rows -> index entries.  Every hit
//                 (target, diagonal, running number) is appended to the region of (slice, target bin) in a per-CTA pool
//                 (global memory, re-used by every query of the CTA: stays in the L2).  Target bins are id ranges
//                 small enough that one bin's last-diagonal table fits a warp's share of shared memory.
//   P2  consume   warp w takes bins w, w+W, ...: for a bin it reads the regions of slice 0, 1, ... in order = the
//                 reference's stream order restricted to the bin, and applies the double-hit rule on a SHARED-MEMORY
//                 table (16 bits + 1 flag bit per target).  No table traffic reaches L2/HBM.
//   P3  score     the same warp scores the bin's candidates, one candidate per lane (16-byte loads of target residues,
//                 query row in shared memory), capped at 255.
//   P4  best      first candidate carrying the target's maximum (keepMaxElement) by atomicMax of
//                 (score, -candidate index) keys in the now idle table; winners go to the CTA's kept list.
//   P5  select    CTA-wide: score histogram -> threshold, the threshold class is cut in the reference's order
//                 (id & binMask, stream order) by a radix select on (bin, running number) keys, uncapped rescoring,
//                 final order, coverage pre-filter, output.
//
// Hit-buffer overflow of the reference (QueryMatcher.cpp:43,280-320: at most maxDbMatches index entries are buffered;
// when the next bucket would reach the limit the buffered part goes through findDuplicates with a fresh table and the
// candidate lists are merged later): the chunk boundary is found from the running numbers after P1 (the bucket that
// holds hit number maxDbMatches-1), P2 restarts the last-diagonal bytes at the boundary and keeps the collapse state,
// which is exactly what findDuplicates + mergeDiagonalDuplicates produce for ONE overflow.  Two or more overflows
// (>= 2*maxDbMatches entries for one query) go through the reference's reversed merge + intermediate keepMax
// (QueryMatcher.cpp:296-303) and are not implemented: the query is reported and the call fails with SD_ELIMIT.

constexpr int PF_WARPS = 8;              // warps per CTA
constexpr int PF_THREADS = PF_WARPS * 32;
constexpr int PF_UNROLL = 4;             // probe steps in flight per warp (independent bitmap loads)
constexpr uint32_t PF_RQCAP = 128;       // row ring: (position, first-row neighbour) rows waiting to be enumerated
constexpr uint32_t PF_HQCAP = 256;       // hit-row ring: non-empty buckets {bucket row, position} waiting for their entries
constexpr uint32_t PF_HQDRAIN = 128;     // stage B runs when the ring holds this many: its rows were prefetched >= 96 pushes ago
constexpr uint32_t PF_PROFLEN = 1168;    // query positions whose 21-column diagonal profile is staged in shared memory
constexpr uint32_t PF_MAXBINSIZE = 8192; // 13 bits of target-in-bin in a pool entry
constexpr uint32_t PF_MAXBOUND = 15;     // hit-buffer overflows per query that are followed exactly (16 chunks)
constexpr uint32_t PF_SEQBITS = 28;      // running number inside a slice; slice index above
static_assert(PF_WARPS <= 16, "slice index must fit above the running number");

struct PrefilterArgs {
    // target side
    const uint8_t *tRes;
    const uint32_t *tOff;
    uint32_t nT;
    const uint2 *bm;           // per 32 k-mers: {occupancy bits, row of the first non-empty k-mer}
    const uint2 *rows;         // per non-empty k-mer: {target, 1<<31 | pos} (single entry) or {start, length}
    const uint2 *ent;          // bucket entries: (target id, first position of the k-mer in that target)
    const uint32_t *e3pair;    // sorted 3-mer neighbour rows 8000 x 8000: score (low 16 bits) | 3-mer index << 16
    const uint16_t *e3cnt;     // [8000][512] cumulative counts: e3cnt[x*512 + 255 - v] = #{y : score(x,y) >= v}
    const int8_t *diagSub;     // 21x21 BLOSUM62 x2 (-0.2)
    // query side
    QueryBatchDev q;
    uint32_t nQ;
    const uint32_t *order;     // queries, longest first
    uint32_t queryFirst;       // identity = queryFirst + q when sameDb
    int sameDb;
    // parameters
    uint32_t maxHits;          // min(max_seqs, nT)
    uint32_t minDiagScore;
    float covThr;
    int covMode;
    uint32_t binMask;          // reference radix (tie order of the --max-seqs cut)
    unsigned long long maxDbMatches;   // QueryMatcher.cpp:43
    // target bins
    uint32_t nBins, binSize;
    unsigned long long binMagic;       // ceil(2^40 / binSize)
    uint32_t tabBytes;         // shared-memory table bytes per warp (>= binSize*2 + binSize/8, multiple of 16)
    // per-CTA scratch in global memory
    uint32_t regionCap;        // pool entries per (slice, bin)
    uint2 *pool;               // [cta][PF_WARPS * nBins * regionCap]
    uint32_t candCap;          // per warp
    uint64_t *cand;            // [cta][PF_WARPS][candCap]  t<<32 | diag<<16 | score<<8
    uint32_t *candSeq;         // [cta][PF_WARPS][candCap]  slice<<28 | running number
    uint32_t keepCap;          // per CTA
    uint4 *keep;               // [cta][keepCap]  {target, diag | score<<16, seq, rescaled score}
    uint32_t *resT;            // [cta][maxHits]
    int32_t *resScore;
    uint32_t *resAux;          // diag | pass<<16 ...
    uint32_t *sortIdx;         // [cta][PF_WARPS][candCap]: candidates of a bin grouped by target (repeated-overflow path only)
    uint16_t *qpackG;          // [cta][maxLq]: staged query rows of queries longer than PF_QCAP
    uint4 *posInfo;            // [cta][maxLq]: per k-mer position {i1 | i2<<16, thr | cutoff1<<16, cutMid | cutBig<<16, similar k-mers}
    uint32_t *posPrefix;       // [cta][maxLq]: inclusive prefix of the position weights (slice balancing)
    uint32_t maxLq;
    // output
    sd_prefilter_hit *out;     // nQ * maxHits
    uint32_t *outCount;        // nQ
    unsigned long long *stats;
    uint32_t *nextQuery;
    uint32_t *errQuery;
    uint32_t *failList;        // queries whose lists did not fit (re-run with larger scratch)
    uint32_t *nFail;
    int debugSkip;             // profiling only (SD_PF_SKIP): 1 = drop stage B (index entries), 2 = drop the bitmap loads too; results are wrong
};

enum { ST_PROBES = 0, ST_MATCHES, ST_DIAGRES, ST_HITS, ST_CANDS, ST_TRUNC, ST_ERR_CAP, ST_ERR_OVERFLOW, ST_CHUNKED, ST_MULTICHUNKED,
       ST_CYC_P1, ST_CYC_P24, ST_CYC_P5, ST_CYC_SEL, ST_CYC_P0B, ST_CYC_STAGEB, ST_CYC_DRAIN, ST_CYC_BIG, ST_CYC_TAIL, ST_N_STAGEB, ST_N_HITROWS, ST_N };   // ST_CYC_*: SM cycles per phase summed over queries (thread 0's view)

__device__ __forceinline__ uint2 ldgU2(const uint2 *p) { return __ldg(p); }

// UngappedAlignment::computeSingelSequenceScores + scalarDiagonalScoring (UngappedAlignment.cpp:30-43,416-430), one
// candidate per lane.  Target residues are read in aligned 16-byte pieces (tRes is padded by 32 bytes).
// Fast path (P3, queries up to PF_PROFLEN residues): prof[pos*21 + targetCode] in shared memory, capped at 255.
__device__ __forceinline__ int diagonalScoreProf(const uint8_t *__restrict__ tRes, const uint32_t *__restrict__ tOff,
                                                 const int8_t *prof, int Lq, uint32_t t, int diag, unsigned long long &residues) {
    const uint32_t tb = __ldg(&tOff[t]);
    const int Lt = (int)(__ldg(&tOff[t + 1]) - tb);
    int qs, ts, len;
    if (diag >= 0) {
        if (diag >= Lq) return 0;
        qs = diag; ts = 0; len = min(Lt, Lq - diag);
    } else {
        if (-diag >= Lt) return 0;
        qs = 0; ts = -diag; len = min(Lt + diag, Lq);
    }
    residues += (unsigned long long)len;
    const uint32_t start = tb + (uint32_t)ts;
    const uint32_t base = start & ~15u;
    const int lead = (int)(start - base);
    int best = 0, run = 0;
    const int end = lead + len;
    for (int off = 0; off < end; off += 16) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(tRes + base + off));
        const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
        const int8_t *pr = prof + (qs + off - lead) * AA;      // profile row of byte 0 of this piece
        if (off >= lead && off + 16 <= end) {                  // whole piece inside the diagonal
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const uint32_t tc = (wv[k >> 2] >> (8 * (k & 3))) & 0xffu;
                run = max(run + (int)pr[k * AA + tc], 0);
                best = max(best, run);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const int pos = off + k;
                if (pos >= lead && pos < end) {
                    const uint32_t tc = (wv[k >> 2] >> (8 * (k & 3))) & 0xffu;
                    run = max(run + (int)pr[k * AA + tc], 0);
                    best = max(best, run);
                }
            }
        }
        if (best >= 255) break;
    }
    return best;
}

// Same score on the rare paths (queries longer than PF_PROFLEN, uncapped rescoring): one out-of-line copy, generic pointers.
// useProf: prof[pos*21 + targetCode]; otherwise qp[pos] = query code | profile correction << 8 and sub32[code*32 + targetCode].
__device__ __noinline__ int diagonalScoreSlow(const uint8_t *__restrict__ tRes, const uint32_t *__restrict__ tOff, const int8_t *prof,
                                              const int8_t *sub32, const uint16_t *qp, int useProf, int capped, int Lq, uint32_t t,
                                              int diag, unsigned long long *residues) {
    const uint32_t tb = __ldg(&tOff[t]);
    const int Lt = (int)(__ldg(&tOff[t + 1]) - tb);
    int qs, ts, len;
    if (diag >= 0) {
        if (diag >= Lq) return 0;
        qs = diag; ts = 0; len = min(Lt, Lq - diag);
    } else {
        if (-diag >= Lt) return 0;
        qs = 0; ts = -diag; len = min(Lt + diag, Lq);
    }
    *residues += (unsigned long long)len;
    const uint8_t *tc = tRes + tb + ts;
    int best = 0, run = 0;
    for (int p = 0; p < len; p++) {
        const uint32_t c = __ldg(&tc[p]);
        int sc;
        if (useProf) sc = (int)prof[(qs + p) * AA + c];
        else { const uint32_t qv = qp[qs + p]; sc = (int)sub32[(qv & 0xffu) * 32u + c] + (int)(int8_t)(qv >> 8); }
        run = max(run + sc, 0);
        best = max(best, run);
        if (capped && best >= 255) break;
    }
    return best;
}

__device__ __forceinline__ void prefetchL2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// P1, stage B: 32 (or the last n) non-empty buckets from the hit-row ring -> their index entries -> (target bin, running number)
// appended to this warp's regions.  Out of line: it is large.  Latency is taken out of the dependent chain bucket row -> entries
// by working one batch ahead: the rows of THIS batch were loaded by the previous call (rowBuf; their entry lists prefetched into
// the L2 then), and this call loads the rows of the next batch (nAhead entries wait behind this one; their rows were prefetched
// when they were pushed) before it appends.
// Returns the new running number; bit 31 set = a region overflowed.
__device__ __noinline__ uint32_t prefilterStageB(const uint2 *__restrict__ rows, const uint2 *__restrict__ ent, unsigned long long binMagic,
                                                 uint32_t binSize, uint32_t regionCap, const uint2 *HQ, uint32_t hqH, uint32_t n, uint32_t nAhead,
                                                 uint2 *rowBuf, uint32_t rowBufValid, uint32_t *cntW, uint2 *poolW, uint32_t seqBase) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t ltMask = (1u << lane) - 1u;
    __syncwarp();
    uint32_t rx = 0, ry = 0, m = 0, pq = 0;
    bool overflow = false;
    if (lane < n) {
        const uint2 h = HQ[(hqH + lane) & (PF_HQCAP - 1)];      // {bucket row, position}
        const uint2 row = (lane < rowBufValid) ? rowBuf[lane] : __ldg(&rows[h.x]);     // rowBufValid: rows loaded by the previous call
        pq = h.y;
        rx = row.x;
        ry = row.y;
        m = (ry >> 31) ? 1u : ry;                // >= 1: only non-empty buckets are queued
    }
    // rows of the next batch (in flight while this batch is appended)
    uint2 nextRow = make_uint2(0u, 0x80000000u);
    const bool haveNext = lane < min(nAhead, 32u);
    if (haveNext) nextRow = __ldg(&rows[HQ[(hqH + 32u + lane) & (PF_HQCAP - 1)].x]);
    uint32_t totalM;
    const uint32_t exclM = warpExclusiveScan(m, totalM);
    for (uint32_t ec = 0; ec < totalM; ec += 32) {
        const uint32_t eidx = ec + lane;
        const bool valid = eidx < totalM;
        const uint32_t owner = warpOwnerOfSlot(exclM, m != 0, ec, lane);
        const uint32_t oExcl = __shfl_sync(0xffffffffu, exclM, owner);
        const uint32_t oX = __shfl_sync(0xffffffffu, rx, owner);
        const uint32_t oY = __shfl_sync(0xffffffffu, ry, owner);
        const uint32_t oP = __shfl_sync(0xffffffffu, pq, owner);
        uint32_t t = 0, j = 0, k = 0;
        if (valid) {
            k = eidx - oExcl;
            if (oY >> 31) { t = oX; j = oY & 0xffffu; }
            else { const uint2 en = __ldg(&ent[oX + k]); t = en.x; j = en.y; }
        }
        const uint32_t bin = (uint32_t)(((unsigned long long)t * binMagic) >> 40);
        const uint32_t grp = __match_any_sync(0xffffffffu, valid ? bin : (0x80000000u | lane));
        const uint32_t leader = (uint32_t)__ffs(grp) - 1u;
        uint32_t old = 0;
        if (valid && lane == leader) { old = cntW[bin]; cntW[bin] = old + (uint32_t)__popc(grp); }
        old = __shfl_sync(0xffffffffu, old, leader);
        if (valid) {
            const uint32_t slot = old + (uint32_t)__popc(grp & ltMask);
            const uint32_t tl = t - bin * binSize;
            const uint32_t diag16 = (oP - j) & 0xffffu;
            if (slot < regionCap)
                poolW[(size_t)bin * regionCap + slot] = make_uint2(tl | (diag16 << 13) | ((k == 0 ? 1u : 0u) << 29), seqBase + lane);
            else overflow = true;
        }
        seqBase += min(32u, totalM - ec);
        __syncwarp();
    }
    // next batch: keep its rows, prefetch its entry lists (first 128 bytes; longer lists stream behind them)
    if (haveNext) {
        rowBuf[lane] = nextRow;
        if (!(nextRow.y >> 31)) {
            prefetchL2(&ent[nextRow.x]);
            if (nextRow.y > 8u) prefetchL2(&ent[nextRow.x + 8u]);
        }
    }
    __syncwarp();
    if (__any_sync(0xffffffffu, overflow) || seqBase >= (1u << PF_SEQBITS)) seqBase |= 0x80000000u;
    return seqBase;
}

// Repeated hit-buffer overflow (QueryMatcher.cpp:296-303,322-333), per target.  With k >= 2 overflows the reference's candidate
// lists D_0..D_k (one findDuplicates result per chunk) are folded like this:
//   overflow j = 2..k:  S = F ++ D_{j-1}  (F = D_0 before the first fold, afterwards the single survivor of the previous fold);
//                       walk S BACKWARDS, keep an element if it carries a score or its 8-bit diagonal differs from the element
//                       after it (mergeDiagonalKeepScoredHitsDuplicates); score the unscored ones; the survivor is the first
//                       maximum in that reversed order (keepMaxElement).  A score of 0 counts as "no score" again.
//   end:                S = F ++ D_k forwards, keep an element if its 8-bit diagonal differs from the element before it
//                       (mergeDiagonalDuplicates); score the unscored ones; first maximum wins.
// The candidates of the bin (stream order, chunk in the low byte, capped score already in place) are grouped by target with a
// counting sort on the idle table words, then one lane folds one target.  Winners go to the CTA's kept list with their arrival
// order (block of the origin chunk, direction).  cands / residues: the reference's counters for this path.
__device__ __noinline__ void prefilterMergeChunks(const uint64_t *candW, const uint32_t *candSeqW, uint32_t nC, uint32_t *sortIdx,
                                                  uint32_t *tab32, uint32_t tabWords, uint32_t binBase, uint32_t binSize, uint32_t k,
                                                  const uint32_t *blockRank, const uint32_t *blockDir, const uint32_t *__restrict__ tOff, int Lq,
                                                  uint4 *keepC, uint32_t keepCap, uint32_t *nKeep, uint32_t *hist,
                                                  unsigned long long *candsOut, unsigned long long *residuesOut) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t ltMask = (1u << lane) - 1u;
    unsigned long long cands = 0, residues = 0;
    for (uint32_t lo = 0; lo < binSize; lo += tabWords) {
        for (uint32_t i = lane; i < tabWords; i += 32) tab32[i] = 0u;
        __syncwarp();
        for (uint32_t c = lane; c < nC; c += 32) {
            const uint32_t tl = (uint32_t)(candW[c] >> 32) - binBase - lo;
            if (tl < tabWords) atomicAdd(&tab32[tl], 1u);
        }
        __syncwarp();
        uint32_t carry = 0;
        for (uint32_t i0 = 0; i0 < tabWords; i0 += 32) {           // exclusive prefix: start of every target's run
            const uint32_t v = (i0 + lane < tabWords) ? tab32[i0 + lane] : 0u;
            uint32_t x = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
                if (lane >= (uint32_t)d) x += y;
            }
            if (i0 + lane < tabWords) tab32[i0 + lane] = x - v + carry;
            carry += __shfl_sync(0xffffffffu, x, 31);
        }
        __syncwarp();
        for (uint32_t c0 = 0; c0 < nC; c0 += 32) {                 // stable scatter; afterwards tab32[tl] = END of the run
            const uint32_t c = c0 + lane;
            uint32_t tl = 0xffffffffu;
            if (c < nC) tl = (uint32_t)(candW[c] >> 32) - binBase - lo;
            const bool valid = tl < tabWords;
            const uint32_t grp = __match_any_sync(0xffffffffu, valid ? tl : (0x80000000u | lane));
            const uint32_t leader = (uint32_t)__ffs(grp) - 1u;
            uint32_t base = 0;
            if (valid && lane == leader) { base = tab32[tl]; tab32[tl] = base + (uint32_t)__popc(grp); }
            base = __shfl_sync(0xffffffffu, base, leader);
            if (valid) sortIdx[base + (uint32_t)__popc(grp & ltMask)] = c;
            __syncwarp();
        }
        const uint32_t nTargets = min(tabWords, binSize - lo);
        for (uint32_t t0 = 0; t0 < nTargets; t0 += 32) {
            const uint32_t tl = t0 + lane;
            int best = -1;
            uint32_t bestScore = 0;
            if (tl < nTargets) {
                const uint32_t b = tl ? tab32[tl - 1] : 0u, e = tab32[tl];
                if (e > b) {
                    const uint32_t t = binBase + lo + tl;
                    const int Lt = (int)(__ldg(&tOff[t + 1]) - __ldg(&tOff[t]));
                    auto elem = [&](uint32_t i) -> uint64_t { return candW[sortIdx[i]]; };
                    auto diagLen = [&](uint64_t v) -> uint32_t {                  // residues the diagonal scorer reads
                        const int diag = (int)(int16_t)(uint16_t)(v >> 16);
                        if (diag >= 0) return diag >= Lq ? 0u : (uint32_t)min(Lt, Lq - diag);
                        return -diag >= Lt ? 0u : (uint32_t)min(Lt + diag, Lq);
                    };
                    int f = -1;                        // survivor of the folds so far (position in the run)
                    uint32_t fScore = 0;
                    uint32_t p = b;
                    for (uint32_t j = 2; j <= k; j++) {
                        uint32_t s0b = p, s0e = p;
                        if (j == 2) { while (s0e < e && (uint32_t)(elem(s0e) & 0xffu) == 0u) s0e++; p = s0e; }
                        const uint32_t cb = p;
                        uint32_t ce = p;
                        while (ce < e && (uint32_t)(elem(ce) & 0xffu) == j - 1u) ce++;
                        p = ce;
                        bool first = true;
                        uint32_t table = 0;
                        int bst = -1;
                        uint32_t bstScore = 0;
                        auto visit = [&](uint32_t i, bool isF) {
                            const uint64_t v = elem(i);
                            const uint32_t d8 = (uint32_t)(v >> 16) & 0xffu;
                            const bool scored = isF && fScore != 0;
                            const bool keep = first || scored || d8 != table;
                            table = d8;
                            first = false;
                            if (keep) {
                                const uint32_t sc = scored ? fScore : ((uint32_t)(v >> 8) & 0xffu);
                                if (!scored) residues += diagLen(v);
                                if (bst < 0 || sc > bstScore) { bst = (int)i; bstScore = sc; }
                            }
                        };
                        for (uint32_t x = ce; x-- > cb;) visit(x, false);
                        if (j == 2) { for (uint32_t x = s0e; x-- > s0b;) visit(x, false); }
                        else if (f >= 0) visit((uint32_t)f, true);
                        if (bst >= 0) { f = bst; fScore = bstScore; }
                    }
                    {
                        bool first = true;
                        uint32_t prevD8 = 0;
                        auto visit = [&](uint32_t i, bool isF) {
                            const uint64_t v = elem(i);
                            const uint32_t d8 = (uint32_t)(v >> 16) & 0xffu;
                            const bool keep = first || d8 != prevD8;
                            prevD8 = d8;
                            first = false;
                            if (keep) {
                                cands++;
                                const bool scored = isF && fScore != 0;
                                const uint32_t sc = scored ? fScore : ((uint32_t)(v >> 8) & 0xffu);
                                if (!scored) residues += diagLen(v);
                                if (best < 0 || sc > bestScore) { best = (int)i; bestScore = sc; }
                            }
                        };
                        if (f >= 0) visit((uint32_t)f, true);
                        for (uint32_t x = p; x < e; x++) visit(x, false);
                    }
                }
            }
            const bool win = best >= 0;
            const uint32_t wb = __ballot_sync(0xffffffffu, win);
            uint32_t base = 0;
            if (lane == 0 && wb) base = atomicAdd(nKeep, (uint32_t)__popc(wb));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (win) {
                const uint32_t ci = sortIdx[best];
                const uint64_t v = candW[ci];
                const uint32_t chunk = (uint32_t)(v & 0xffu);
                const uint32_t g = candSeqW[ci];
                const uint32_t pos = base + (uint32_t)__popc(wb & ltMask);
                if (pos < keepCap)
                    keepC[pos] = make_uint4((uint32_t)(v >> 32), ((uint32_t)(v >> 16) & 0xffffu) | (bestScore << 16), blockDir[chunk] ? g : ~g, blockRank[chunk] << 8);
                atomicAdd(&hist[bestScore], 1u);
            }
            __syncwarp();
        }
    }
    for (int d = 16; d > 0; d >>= 1) cands += __shfl_xor_sync(0xffffffffu, cands, d);
    if (lane == 0) *candsOut += cands;          // the caller keeps this counter on lane 0 and reduces the residues itself
    *residuesOut += residues;
}

// Shared memory: static part below + dynamic [counters: PF_WARPS*nBins u32][tables: PF_WARPS * tabBytes]
struct PfShared {
    int8_t sub32[AA * 32];
    uint32_t hist[256];
    union {      // the rings live in P1, the diagonal profile from P2 on
        struct { uint32_t rq[PF_WARPS][PF_RQCAP]; uint2 hq[PF_WARPS][PF_HQCAP]; uint2 rowBuf[PF_WARPS][32]; } ring;
        int8_t prof[PF_PROFLEN * AA];
    } u;
    uint32_t sliceTotal[PF_WARPS];
    uint32_t scanTmp[PF_WARPS];
    unsigned long long totalWeight;
    long long prof1[6];
    long long clkP0b;
    unsigned long long sliceBase[PF_WARPS + 1];
    unsigned long long stat[8];
    unsigned long long selPrefix;
    uint32_t q, fail, nKeep, nRes, nAt, thrScore, selCount, selNeed, nOut;
    uint32_t nBound, boundSearch;
    uint32_t boundSlice[PF_MAXBOUND], boundSeq[PF_MAXBOUND];     // chunk c+1 starts at (slice, running number)
    uint32_t blockRank[PF_MAXBOUND + 1], blockDir[PF_MAXBOUND + 1];   // final arrival order of the surviving element by origin chunk
    int maxSelf;
};

__global__ void __launch_bounds__(PF_THREADS, 3) prefilterKernel(PrefilterArgs a) {
    __shared__ PfShared S;
    extern __shared__ __align__(16) uint8_t sDyn[];

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31;
    const uint32_t w = tid >> 5;
    const uint32_t ltMask = (1u << lane) - 1u;
    const uint32_t nBins = a.nBins;
    uint32_t *cntAll = reinterpret_cast<uint32_t *>(sDyn);                    // [slice][bin]
    uint32_t *cntW = cntAll + w * nBins;
    uint8_t *tabW = sDyn + (((size_t)PF_WARPS * nBins * 4 + 15) & ~(size_t)15) + (size_t)w * a.tabBytes;
    uint16_t *tab16 = reinterpret_cast<uint16_t *>(tabW);                     // [binSize]  last d8 | last emitted d8 << 8
    uint32_t *tabFlag = reinterpret_cast<uint32_t *>(tabW + (((size_t)a.binSize * 2 + 15) & ~(size_t)15));   // emitted bits
    uint32_t *tab32 = reinterpret_cast<uint32_t *>(tabW);                     // P4 view
    const uint32_t tabWords = a.tabBytes / 4;

    uint2 *poolC = a.pool + (size_t)blockIdx.x * PF_WARPS * nBins * a.regionCap;
    uint64_t *candW = a.cand + ((size_t)blockIdx.x * PF_WARPS + w) * a.candCap;
    uint32_t *candSeqW = a.candSeq + ((size_t)blockIdx.x * PF_WARPS + w) * a.candCap;
    uint4 *keepC = a.keep + (size_t)blockIdx.x * a.keepCap;
    uint32_t *resT = a.resT + (size_t)blockIdx.x * a.maxHits;
    int32_t *resScore = a.resScore + (size_t)blockIdx.x * a.maxHits;
    uint32_t *resAux = a.resAux + (size_t)blockIdx.x * a.maxHits;
    uint16_t *qpackG = a.qpackG + (size_t)blockIdx.x * a.maxLq;

    for (uint32_t i = tid; i < AA * 32; i += PF_THREADS) S.sub32[i] = ((i & 31) < AA) ? a.diagSub[(i >> 5) * AA + (i & 31)] : (int8_t)0;

    for (;;) {
        __syncthreads();
        if (tid == 0) {
            const uint32_t qi = atomicAdd(a.nextQuery, 1u);
            S.q = (qi < a.nQ) ? a.order[qi] : 0xffffffffu;
            S.fail = 0; S.nKeep = 0; S.nRes = 0; S.nAt = 0; S.nBound = 0; S.nOut = 0;
            for (int i = 0; i < 8; i++) S.stat[i] = 0;
        }
        __syncthreads();
        const uint32_t q = S.q;
        if (q == 0xffffffffu) break;
        const uint32_t qBase = a.q.off[q];
        const int Lq = (int)(a.q.off[q + 1] - qBase);
        const uint32_t identity = a.sameDb ? (a.queryFirst + q) : 0xffffffffu;
        const bool useProf = Lq <= (int)PF_PROFLEN;
        const uint16_t *qp = qpackG;

        // ---------------- P0: counters, histogram, staged query rows
        for (uint32_t i = tid; i < PF_WARPS * nBins; i += PF_THREADS) cntAll[i] = 0;
        for (uint32_t i = tid; i < 256; i += PF_THREADS) S.hist[i] = 0;
        if (!useProf) {
            for (int p = (int)tid; p < Lq; p += PF_THREADS)
                qpackG[p] = (uint16_t)((uint32_t)a.q.code[qBase + p] | ((uint32_t)(uint8_t)a.q.corr[qBase + p] << 8));
        }
        unsigned long long stProbes = 0, stDiagRes = 0;
        const long long clk0 = clock64();

        // ---------------- P0b: per-position set-up and EXACT similar-k-mer counts, balanced slices
        // (KmerGenerator.cpp:129-137,186-216: the list of position p is {(a, c): s1[a] + s2[c] >= thr}; its size is
        // sum over score levels v of #{a: s1[a] = v} * #{c: s2[c] >= thr - v}, read from the cumulative count table)
        const int nPos = max(Lq - SPAN + 1, 0);
        uint4 *posInfo = a.posInfo + (size_t)blockIdx.x * a.maxLq;
        uint32_t *posPrefix = a.posPrefix + (size_t)blockIdx.x * a.maxLq;        // inclusive prefix of the slice weights
        {
            uint32_t carry = 0;
            for (int p0 = 0; p0 < nPos; p0 += PF_THREADS) {
                const int p = p0 + (int)tid;
                uint32_t wgt = 0;
                if (p < nPos) {
                    uint4 info = make_uint4(0u, 0xffffu, 0u, 0u);               // threshold -1: no k-mer here
                    const int thr = a.q.thr[qBase + p];
                    if (thr >= 0) {
                        const uint8_t *qc = a.q.code + qBase + p;
                        const uint32_t i1 = qc[0] + 20u * qc[1] + 400u * qc[3];
                        const uint32_t i2 = qc[5] + 20u * qc[8] + 400u * qc[9];
                        const int top1 = (int)(int16_t)(__ldg(&a.e3pair[(size_t)i1 * N3]) & 0xffffu);
                        const uint32_t *row2 = a.e3pair + (size_t)i2 * N3;
                        const int top2 = (int)(int16_t)(__ldg(&row2[0]) & 0xffffu);
                        if (top1 + top2 >= thr) {                                // else nothing reaches the threshold
                            const int cutoff1 = thr - top2;                      // first-row elements with s1 >= cutoff1 have a partner
                            const int cutMid = thr - (int)(int16_t)(__ldg(&row2[1]) & 0xffffu);     // ... at least 2 partners
                            const int cutBig = thr - (int)(int16_t)(__ldg(&row2[31]) & 0xffffu);    // ... at least 32 partners
                            uint32_t prevC = 0, probes = 0;
                            for (int v = top1; v >= cutoff1; v--) {
                                const uint32_t c1 = countGE(a.e3cnt, i1, v);
                                if (c1 != prevC) probes += (c1 - prevC) * countGE(a.e3cnt, i2, thr - v);
                                prevC = c1;
                            }
                            // the first row is sorted by score: rows [0, nBig) have >= 32 partners, [nBig, nHead) 2..31, [nHead, nRows) one
                            const uint32_t nRows = prevC;
                            const uint32_t nHead = countGE(a.e3cnt, i1, cutMid), nBig = countGE(a.e3cnt, i1, cutBig);
                            info = make_uint4(i1 | (i2 << 16), ((uint32_t)thr & 0xffffu) | (nRows << 16), nBig | (nHead << 16), probes);
                            wgt = probes + 64u;
                        }
                    }
                    posInfo[p] = info;
                }
                // block-wide inclusive scan of wgt
                uint32_t x = wgt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
                    if (lane >= (uint32_t)d) x += y;
                }
                if (lane == 31) S.scanTmp[w] = x;
                __syncthreads();
                uint32_t add = carry;
                for (uint32_t i = 0; i < w; i++) add += S.scanTmp[i];
                uint32_t tot = 0;
                for (uint32_t i = 0; i < PF_WARPS; i++) tot += S.scanTmp[i];
                if (p < nPos) posPrefix[p] = x + add;
                carry += tot;
                __syncthreads();
            }
            if (tid == 0) { S.totalWeight = carry; S.clkP0b = clock64() - clk0; }
        }
        __syncthreads();

        // ---------------- P1: hit stream of this warp's slice of positions
        {
            uint32_t *RQ = S.u.ring.rq[w];                 // queued rows of the current position: first-row neighbour | partners << 16
            uint2 *HQ = S.u.ring.hq[w];
            uint2 *rowBuf = S.u.ring.rowBuf[w];
            uint32_t rowBufValid = 0;               // rows of the ring's first 32 entries already loaded (and their entries prefetched)
            uint32_t rqH = 0, rqN = 0, hqH = 0, hqN = 0;
            uint32_t seqBase = 0;
            bool overflow = false;
            uint2 *poolW = poolC + (size_t)w * nBins * a.regionCap;
            long long cycStageB = 0;
            uint32_t nStageB = 0, nHitRows = 0;

            // slice of positions with 1/PF_WARPS of the query's similar k-mers
            int pLo, pHi;
            {
                const unsigned long long tw = S.totalWeight;
                auto bound = [&](const uint32_t k) -> int {      // first position whose exclusive prefix reaches tw*k/PF_WARPS
                    if (k == 0) return 0;
                    if (k >= PF_WARPS) return nPos;
                    const uint32_t target = (uint32_t)(tw * k / PF_WARPS);
                    int lo = 0, hi = nPos;
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        const uint32_t excl = mid == 0 ? 0u : posPrefix[mid - 1];
                        if (excl >= target) hi = mid; else lo = mid + 1;
                    }
                    return lo;
                };
                pLo = bound(w);
                pHi = bound(w + 1);
            }

            // ONE loop produces the similar k-mers of the slice in stream order, 32*PF_UNROLL at a time, from four generators
            // (a small state machine: the code of the occupancy test, the hit-row ring and stage B exists once):
            //   HEAD   next chunk of 32 first-row neighbours with at least 2 partners: rows with >= 32 partners are handed to BIG,
            //          rows with 2..31 partners are queued;
            //   BIG    one first-row neighbour, lanes across its second-row partners;
            //   DRAIN  the queued rows, flattened: lane-balanced 32 similar k-mers per step;
            //   TAIL   first-row neighbours whose only partner is the best second-row neighbour, lanes across the first row.
            // Positions are walked with a two-deep software pipeline: the set-up record of position p+2 and the first chunk of the
            // first row (and the best second-row neighbour) of position p+1 are loaded while position p is processed.
            enum { M_NEXTPOS, M_HEAD, M_BIG, M_DRAIN, M_TAIL };
            int mode = M_NEXTPOS, afterBig = M_NEXTPOS, afterDrain = M_NEXTPOS;
            uint4 infoB = make_uint4(0u, 0xffffu, 0u, 0u), infoC = infoB;      // p+1, p+2
            uint32_t prB = 0, topB = 0;
            int p = pLo - 1;
            if (pLo < pHi) infoB = posInfo[pLo];
            if (pLo + 1 < pHi) infoC = posInfo[pLo + 1];
            if ((int16_t)(infoB.y & 0xffffu) >= 0) {
                prB = __ldg(&a.e3pair[(size_t)(infoB.x & 0xffffu) * N3 + lane]);
                topB = __ldg(&a.e3pair[(size_t)(infoB.x >> 16) * N3]);
            }
            int thr = 0;
            uint32_t i2 = 0, nRows = 0, nBig = 0, nHead = 0, x2top = 0, pr0 = 0, a0 = 0;
            const uint32_t *row1 = a.e3pair, *row2 = a.e3pair;
            uint32_t xaL = 0, nbL = 0, bigMask = 0;                             // HEAD -> BIG
            uint32_t xb = 0, nbb = 0, c0 = 0;                                   // BIG
            uint32_t dRow = 0, dExcl = 0, dTotal = 0, pc = 0;                   // DRAIN
            for (;;) {
                uint32_t kmerW[PF_UNROLL];
                uint32_t nSteps = 0;
#pragma unroll
                for (int j = 0; j < PF_UNROLL; j++) kmerW[j] = 0xffffffffu;
                if (mode == M_NEXTPOS) {
                    p++;
                    if (p >= pHi) break;
                    const uint4 info = infoB;
                    pr0 = prB;
                    const uint32_t top2w = topB;
                    infoB = infoC;
                    if (p + 2 < pHi) infoC = posInfo[p + 2];
                    if (p + 1 < pHi && (int16_t)(infoB.y & 0xffffu) >= 0) {
                        prB = __ldg(&a.e3pair[(size_t)(infoB.x & 0xffffu) * N3 + lane]);
                        topB = __ldg(&a.e3pair[(size_t)(infoB.x >> 16) * N3]);
                    }
                    thr = (int)(int16_t)(info.y & 0xffffu);
                    if (thr < 0) continue;
                    stProbes += (lane == 0) ? info.w : 0;
                    i2 = info.x >> 16;
                    nRows = info.y >> 16; nBig = info.z & 0xffffu; nHead = info.z >> 16;
                    row1 = a.e3pair + (size_t)(info.x & 0xffffu) * N3;
                    row2 = a.e3pair + (size_t)i2 * N3;
                    x2top = top2w >> 16;
                    a0 = 0;
                    mode = nHead ? M_HEAD : M_TAIL;
                    continue;
                } else if (mode == M_HEAD) {
                    const uint32_t ai = a0 + lane;
                    const uint32_t pr = (a0 == 0) ? pr0 : __ldg(&row1[min(ai, N3 - 1)]);
                    xaL = pr >> 16;
                    const bool isBig = ai < nBig;
                    const bool isMid = !isBig && ai < nHead;
                    nbL = 1;
                    // KmerGenerator.cpp:201: second-row elements with s2 >= thr - s1
                    if (isBig || isMid) nbL = countGE(a.e3cnt, i2, thr - (int)(int16_t)(pr & 0xffffu));
                    bigMask = __ballot_sync(0xffffffffu, isBig);
                    const uint32_t midMask = __ballot_sync(0xffffffffu, isMid);
                    if (isMid) RQ[(rqH + rqN + __popc(midMask & ltMask)) & (PF_RQCAP - 1)] = xaL | (nbL << 16);
                    rqN += __popc(midMask);
                    a0 = min(a0 + 32u, nHead);
                    // stream order: the big rows of this chunk, then (when the head is complete or the queue is full) the queued rows
                    int next = M_HEAD;
                    if (a0 == nHead) { next = rqN ? M_DRAIN : M_TAIL; afterDrain = M_TAIL; }
                    else if (rqN >= PF_RQCAP - 32) { next = M_DRAIN; afterDrain = M_HEAD; }
                    afterBig = next;
                    mode = bigMask ? M_BIG : next;
                    c0 = 0; nbb = 0;
                    pc = 0; dTotal = 0;
                    continue;
                } else if (mode == M_BIG) {
                    if (c0 >= nbb) {
                        if (!bigMask) { mode = afterBig; pc = 0; dTotal = 0; continue; }
                        const int src = __ffs(bigMask) - 1;
                        bigMask &= bigMask - 1;
                        xb = __shfl_sync(0xffffffffu, xaL, src);
                        nbb = __shfl_sync(0xffffffffu, nbL, src);
                        c0 = 0;
                    }
#pragma unroll
                    for (int j = 0; j < PF_UNROLL; j++) {
                        const uint32_t c = c0 + 32u * j + lane;
                        if (c < nbb) kmerW[j] = kmerSlot(xb, __ldg(&row2[c]) >> 16);
                    }
                    nSteps = (min(nbb - c0, 32u * PF_UNROLL) + 31u) / 32u;
                    c0 += 32 * PF_UNROLL;
                } else if (mode == M_DRAIN) {
                    if (pc >= dTotal) {
                        if (rqN == 0) { mode = afterDrain; continue; }
                        // calls of similar size (a 33-row queue is drained as 17 + 16, not 32 + 1)
                        const uint32_t calls = (rqN + 31u) / 32u;
                        const uint32_t n = (rqN + calls - 1u) / calls;
                        __syncwarp();
                        dRow = (lane < n) ? RQ[(rqH + lane) & (PF_RQCAP - 1)] : 0u;
                        rqH += n;
                        rqN -= n;
                        dExcl = warpExclusiveScan(dRow >> 16, dTotal);
                        pc = 0;
                    }
                    const uint32_t nb = dRow >> 16;
#pragma unroll
                    for (int j = 0; j < PF_UNROLL; j++) {
                        const uint32_t base = pc + 32u * j;
                        if (base < dTotal) {
                            const uint32_t owner = warpOwnerOfSlot(dExcl, nb != 0, base, lane);
                            const uint32_t oExcl = __shfl_sync(0xffffffffu, dExcl, owner);
                            const uint32_t oRow = __shfl_sync(0xffffffffu, dRow, owner);
                            if (base + lane < dTotal) kmerW[j] = kmerSlot(oRow & 0xffffu, __ldg(&row2[base + lane - oExcl]) >> 16);   // KmerGenerator.cpp:207
                        }
                    }
                    nSteps = (min(dTotal - pc, 32u * PF_UNROLL) + 31u) / 32u;
                    pc += 32 * PF_UNROLL;
                } else {    // M_TAIL
                    if (a0 >= nRows) { mode = M_NEXTPOS; continue; }
#pragma unroll
                    for (int j = 0; j < PF_UNROLL; j++) {
                        const uint32_t ai = a0 + 32u * j + lane;
                        if (ai < nRows) kmerW[j] = kmerSlot(__ldg(&row1[ai]) >> 16, x2top);
                    }
                    nSteps = (min(nRows - a0, 32u * PF_UNROLL) + 31u) / 32u;
                    a0 += 32 * PF_UNROLL;
                }
                // ---- common: occupancy words of up to 32*PF_UNROLL similar k-mers, non-empty buckets to the hit-row ring, stage B
                uint2 wd[PF_UNROLL];
#pragma unroll
                for (int j = 0; j < PF_UNROLL; j++) {
                    wd[j] = make_uint2(0u, 0u);
                    if (kmerW[j] != 0xffffffffu) wd[j] = (a.debugSkip & 2) ? make_uint2(kmerW[j] * 2654435761u & 0x11111111u, 0u) : ldgU2(&a.bm[kmerW[j] >> 5]);
                }
#pragma unroll
                for (int j = 0; j < PF_UNROLL; j++) {
                    if ((uint32_t)j < nSteps) {
                        const bool has = ((wd[j].x >> (kmerW[j] & 31u)) & 1u) != 0;       // idle lanes carry bits = 0
                        const uint32_t hb = __ballot_sync(0xffffffffu, has);
                        if (has) {
                            const uint32_t ci = wd[j].y + __popc(wd[j].x & ((1u << (kmerW[j] & 31u)) - 1u));
                            prefetchL2(&a.rows[ci]);
                            HQ[(hqH + hqN + __popc(hb & ltMask)) & (PF_HQCAP - 1)] = make_uint2(ci, (uint32_t)p);
                        }
                        hqN += __popc(hb);
                    }
                }
                while (hqN >= PF_HQDRAIN) {
                    const long long t0 = clock64();
                    nStageB++; nHitRows += 32;
                    if (!(a.debugSkip & 3)) {
                        const uint32_t r = prefilterStageB(a.rows, a.ent, a.binMagic, a.binSize, a.regionCap, HQ, hqH, 32u, hqN - 32u, rowBuf, rowBufValid, cntW, poolW, seqBase);
                        overflow = overflow || (r >> 31) != 0;
                        seqBase = r & 0x7fffffffu;
                        rowBufValid = 32;
                    }
                    hqH += 32;
                    hqN -= 32;
                    cycStageB += clock64() - t0;
                }
            }
            while (hqN > 0 && !(a.debugSkip & 3)) {
                const uint32_t n = min(hqN, 32u);
                const uint32_t r = prefilterStageB(a.rows, a.ent, a.binMagic, a.binSize, a.regionCap, HQ, hqH, n, hqN - n, rowBuf, rowBufValid, cntW, poolW, seqBase);
                overflow = overflow || (r >> 31) != 0;
                seqBase = r & 0x7fffffffu;
                rowBufValid = min(hqN - n, 32u);
                hqH += n;
                hqN -= n;
            }
            __syncwarp();
            if (lane == 0) S.sliceTotal[w] = seqBase;
            if (tid == 0) {      // warp 0's view of where P1's cycles go
                S.prof1[0] = cycStageB; S.prof1[1] = 0; S.prof1[2] = 0; S.prof1[3] = 0; S.prof1[4] = nStageB; S.prof1[5] = nHitRows;
            }
            if (__any_sync(0xffffffffu, overflow)) { if (lane == 0) atomicOr(&S.fail, 1u); }
        }
        __syncthreads();

        if (useProf) {
            // UngappedAlignment::createProfile: BLOSUM62 x2 + per-position correction, one 21-column row per position
            // (the rings of P1 are dead now: same shared memory)
            for (int i = (int)tid; i < Lq * AA; i += PF_THREADS) {
                const int p = i / AA, c = i - p * AA;
                S.u.prof[i] = (int8_t)((int)S.sub32[(uint32_t)a.q.code[qBase + p] * 32u + c] + (int)a.q.corr[qBase + p]);
            }
        }
        // ---------------- chunk boundaries of the reference's hit buffer (see the header comment)
        // Chunk c+1 opens with the bucket that holds hit number G_c + maxDbMatches - 1 (G_c = first hit of chunk c): the reference
        // flushes when buffered + bucket size reaches the limit, and the bucket goes to the new chunk (QueryMatcher.cpp:280-314).
        if (tid == 0) {
            unsigned long long acc = 0;
            for (int s = 0; s < PF_WARPS; s++) { S.sliceBase[s] = acc; acc += S.sliceTotal[s]; }
            S.sliceBase[PF_WARPS] = acc;
            S.stat[ST_MATCHES] = acc;
            S.nBound = 0;
            S.boundSearch = (acc >= a.maxDbMatches) ? 1u : 0u;
            S.selPrefix = 0ull;                                             // G of the current chunk
        }
        __syncthreads();
        while (S.boundSearch && !S.fail) {
            if (tid == 0) {
                const unsigned long long x = S.selPrefix + a.maxDbMatches - 1ull;
                int s = 0;
                while (S.sliceBase[s + 1] <= x) s++;
                S.boundSlice[S.nBound] = (uint32_t)s;
                S.boundSeq[S.nBound] = (uint32_t)(x - S.sliceBase[s]);      // turned into the bucket start below
                S.selCount = 0;
            }
            __syncthreads();
            {   // largest bucket start (first-entry flag) <= that hit in its slice, over all bins
                const uint32_t bs = S.boundSlice[S.nBound], bx = S.boundSeq[S.nBound];
                uint32_t best = 0;
                for (uint32_t bin = w; bin < nBins; bin += PF_WARPS) {
                    const uint32_t n = min(cntAll[bs * nBins + bin], a.regionCap);
                    const uint2 *reg = poolC + ((size_t)bs * nBins + bin) * a.regionCap;
                    for (uint32_t i = lane; i < n; i += 32) {
                        const uint2 e = reg[i];
                        if (((e.x >> 29) & 1u) && e.y <= bx) best = max(best, e.y);
                    }
                }
                for (int d = 16; d > 0; d >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, d));
                if (lane == 0) atomicMax(&S.selCount, best);
            }
            __syncthreads();
            if (tid == 0) {
                const uint32_t k = S.nBound;
                S.boundSeq[k] = S.selCount;
                S.selPrefix = S.sliceBase[S.boundSlice[k]] + S.selCount;
                S.nBound = k + 1;
                S.boundSearch = (S.sliceBase[PF_WARPS] - S.selPrefix >= a.maxDbMatches) ? 1u : 0u;
                if (S.boundSearch && S.nBound >= PF_MAXBOUND) S.fail |= 2u;
            }
            __syncthreads();
        }
        if (tid == 0 && S.nBound >= 2) {
            // Arrival order of what survives the repeated merges (ties of the --max-seqs cut): every overflow from the second on
            // REVERSES the merged list (mergeDiagonalKeepScoredHitsDuplicates walks backwards, CacheFriendlyOperations.cpp:132-142)
            // before keepMax; the last chunk is appended in stream order.  blocks[] = origin chunks in final order.
            const uint32_t k = S.nBound;                                    // overflows; chunks 0..k
            uint32_t blk[PF_MAXBOUND + 1], dir[PF_MAXBOUND + 1], nb = 1;
            blk[0] = 0; dir[0] = 1;                                         // F_1 = chunk 0, stream order
            for (uint32_t j = 2; j <= k; j++) {                             // F_j = reversed(F_{j-1} ++ chunk j-1)
                uint32_t nblk[PF_MAXBOUND + 1], ndir[PF_MAXBOUND + 1];
                nblk[0] = j - 1; ndir[0] = 0;
                for (uint32_t i = 0; i < nb; i++) { nblk[1 + i] = blk[nb - 1 - i]; ndir[1 + i] = dir[nb - 1 - i] ^ 1u; }
                nb++;
                for (uint32_t i = 0; i < nb; i++) { blk[i] = nblk[i]; dir[i] = ndir[i]; }
            }
            for (uint32_t i = 0; i < nb; i++) { S.blockRank[blk[i]] = i; S.blockDir[blk[i]] = dir[i]; }
            S.blockRank[k] = nb; S.blockDir[k] = 1;
        }
        __syncthreads();

        const long long clk1 = clock64();
        // ---------------- P2-P4: per warp, bin by bin
        if (!S.fail) {
            const uint32_t nBound = S.nBound;
            const bool multi = nBound >= 2;            // repeated overflow: chunk-local lists, folded per target below
            uint32_t *sortIdxW = a.sortIdx + ((size_t)blockIdx.x * PF_WARPS + w) * a.candCap;
            bool overflow = false;
            unsigned long long stCands = 0;
            for (uint32_t bin = w; bin < nBins; bin += PF_WARPS) {
                const uint32_t binBase = bin * a.binSize;
                for (uint32_t i = lane; i < tabWords; i += 32) tab32[i] = 0u;
                __syncwarp();
                uint32_t nC = 0;
                // ---- P2: double-hit rule in stream order (CacheFriendlyOperations::findDuplicates restated per target)
                auto rule = [&](const uint2 *reg, const uint32_t s, const uint32_t lo, const uint32_t hi, const uint32_t chunk) {
                    uint2 nxt = make_uint2(0u, 0u);
                    if (lo + lane < hi) nxt = reg[lo + lane];
                    for (uint32_t e0 = lo; e0 < hi; e0 += 32) {
                        const bool validE = e0 + lane < hi;
                        const uint32_t ex = nxt.x, ey = nxt.y;
                        if (e0 + 32 + lane < hi) nxt = reg[e0 + 32 + lane];       // next step's entries are in flight during this one
                        const uint32_t tl = validE ? (ex & 0x1fffu) : (0x80000000u | lane);     // unique dummy for idle lanes
                        const uint32_t diag16 = (ex >> 13) & 0xffffu;
                        const uint32_t d8 = diag16 & 0xffu;
                        const uint32_t grp = __match_any_sync(0xffffffffu, tl);
                        uint32_t old = 0, oldFlag = 0;
                        if (validE) { old = tab16[tl]; oldFlag = (tabFlag[tl >> 5] >> (tl & 31u)) & 1u; }
                        const uint32_t lowerG = grp & ltMask;
                        const int prevLane = lowerG ? (31 - __clz(lowerG)) : -1;
                        uint32_t prevD8 = __shfl_sync(0xffffffffu, d8, prevLane >= 0 ? prevLane : (int)lane);
                        if (prevLane < 0) prevD8 = old & 0xffu;
                        const bool emit = validE && (d8 == prevD8);
                        const uint32_t emitBallot = __ballot_sync(0xffffffffu, emit);
                        const uint32_t grpE = grp & emitBallot;
                        const uint32_t lowerE = grpE & ltMask;
                        const int prevELane = lowerE ? (31 - __clz(lowerE)) : -1;
                        uint32_t prevED8 = __shfl_sync(0xffffffffu, d8, prevELane >= 0 ? prevELane : (int)lane);
                        bool prevEExists = true;
                        if (prevELane < 0) { prevED8 = old >> 8; prevEExists = oldFlag != 0; }
                        const bool isCand = emit && !(prevEExists && prevED8 == d8);
                        const int lastELane = grpE ? (31 - __clz(grpE)) : -1;
                        const uint32_t lastED8 = __shfl_sync(0xffffffffu, d8, lastELane >= 0 ? lastELane : (int)lane);
                        if (validE && (grp >> lane) == 1u) {        // last hit of this target in the step
                            tab16[tl] = (uint16_t)(d8 | ((lastELane >= 0 ? lastED8 : (old >> 8)) << 8));
                            if (lastELane >= 0 && !oldFlag) atomicOr(&tabFlag[tl >> 5], 1u << (tl & 31u));
                        }
                        const uint32_t candBallot = __ballot_sync(0xffffffffu, isCand);
                        if (isCand) {
                            const uint32_t pos = nC + __popc(candBallot & ltMask);
                            if (pos < a.candCap) {
                                candW[pos] = ((uint64_t)(binBase + tl) << 32) | ((uint64_t)diag16 << 16) | chunk;
                                candSeqW[pos] = (s << PF_SEQBITS) | ey;
                            }
                        }
                        nC += __popc(candBallot);
                        __syncwarp();
                    }
                };
                uint32_t chunk = 0;                                     // chunk of the entries that come next
                for (uint32_t s = 0; s < PF_WARPS; s++) {
                    const uint32_t n = cntAll[s * nBins + bin];
                    const uint2 *reg = poolC + ((size_t)s * nBins + bin) * a.regionCap;
                    uint32_t lo = 0;
                    for (;;) {
                        uint32_t hi = n;
                        const bool boundaryHere = chunk < nBound && S.boundSlice[chunk] <= s;
                        if (boundaryHere) {
                            hi = lo;
                            if (S.boundSlice[chunk] == s) {
                                // entries are ordered by running number: count those before the boundary
                                const uint32_t bseq = S.boundSeq[chunk];
                                uint32_t c = 0;
                                for (uint32_t i = lane; i < n; i += 32) c += (reg[i].y < bseq) ? 1u : 0u;
                                for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
                                hi = max(c, lo);
                            }
                        }
                        rule(reg, s, lo, hi, multi ? chunk : 0u);
                        lo = hi;
                        if (!boundaryHere) break;
                        // the next chunk starts: the reference runs findDuplicates on a cleared table.  One overflow: the collapse
                        // state (last emitted diagonal) lives on through mergeDiagonalDuplicates, only the last-diagonal bytes
                        // restart.  Repeated overflow: chunk-local lists, everything restarts.
                        if (multi) { for (uint32_t i = lane; i < tabWords; i += 32) tab32[i] = 0u; }
                        else { for (uint32_t i = lane; i < a.binSize; i += 32) tab16[i] &= 0xff00u; }
                        __syncwarp();
                        chunk++;
                    }
                }
                if (nC > a.candCap) { overflow = true; nC = a.candCap; }
                if (!multi) stCands += (lane == 0) ? nC : 0;
                __syncwarp();
                // ---- P3: diagonal scores, one candidate per lane, capped at 255 (UngappedAlignment::align)
                for (uint32_t c0 = 0; c0 < nC; c0 += 32) {
                    const uint32_t c = c0 + lane;
                    if (c < nC) {
                        const uint64_t v = candW[c];
                        const uint32_t ct = (uint32_t)(v >> 32);
                        const int cd = (int)(int16_t)(uint16_t)(v >> 16);
                        unsigned long long resHere = 0;
                        const int sc = useProf ? diagonalScoreProf(a.tRes, a.tOff, S.u.prof, Lq, ct, cd, resHere)
                                               : diagonalScoreSlow(a.tRes, a.tOff, S.u.prof, S.sub32, qp, 0, 1, Lq, ct, cd, &resHere);
                        if (!multi) stDiagRes += resHere;     // repeated overflow: the fold below counts what the reference scores
                        candW[c] = v | ((uint64_t)min(sc, 255) << 8);
                    }
                }
                __syncwarp();
                if (multi) {
                    prefilterMergeChunks(candW, candSeqW, nC, sortIdxW, tab32, tabWords, binBase, a.binSize, nBound, S.blockRank, S.blockDir,
                                         a.tOff, Lq, keepC, a.keepCap, &S.nKeep, S.hist, &stCands, &stDiagRes);
                    continue;
                }
                // ---- P4: keepMaxScoreElementOnly: first candidate (stream order) carrying the target's maximum
                for (uint32_t lo = 0; lo < a.binSize; lo += tabWords) {
                    for (uint32_t i = lane; i < tabWords; i += 32) tab32[i] = 0u;
                    __syncwarp();
                    for (uint32_t c = lane; c < nC; c += 32) {
                        const uint64_t v = candW[c];
                        const uint32_t tl = (uint32_t)(v >> 32) - binBase - lo;
                        if (tl < tabWords) atomicMax(&tab32[tl], ((uint32_t)((v >> 8) & 0xffu) << 24) | (0xffffffu - c));
                    }
                    __syncwarp();
                    for (uint32_t c0 = 0; c0 < nC; c0 += 32) {
                        const uint32_t c = c0 + lane;
                        bool win = false;
                        uint64_t v = 0;
                        if (c < nC) {
                            v = candW[c];
                            const uint32_t tl = (uint32_t)(v >> 32) - binBase - lo;
                            win = tl < tabWords && tab32[tl] == (((uint32_t)((v >> 8) & 0xffu) << 24) | (0xffffffu - c));
                        }
                        const uint32_t wb = __ballot_sync(0xffffffffu, win);
                        uint32_t base = 0;
                        if (lane == 0 && wb) base = atomicAdd(&S.nKeep, (uint32_t)__popc(wb));
                        base = __shfl_sync(0xffffffffu, base, 0);
                        if (win) {
                            const uint32_t pos = base + __popc(wb & ltMask);
                            const uint32_t sc = (uint32_t)(v >> 8) & 0xffu;
                            if (pos < a.keepCap) keepC[pos] = make_uint4((uint32_t)(v >> 32), ((uint32_t)(v >> 16) & 0xffffu) | (sc << 16), candSeqW[c], 0u);
                            atomicAdd(&S.hist[sc], 1u);
                        }
                    }
                    __syncwarp();
                }
            }
            if (__any_sync(0xffffffffu, overflow)) { if (lane == 0) atomicOr(&S.fail, 1u); }
            for (int d = 16; d > 0; d >>= 1) stDiagRes += __shfl_down_sync(0xffffffffu, stDiagRes, d);
            if (lane == 0) {
                atomicAdd(&S.stat[ST_DIAGRES], stDiagRes);
                atomicAdd(&S.stat[ST_CANDS], stCands);
            }
        }
        if (lane == 0) atomicAdd(&S.stat[ST_PROBES], stProbes);
        __syncthreads();
        if (tid == 0 && S.nKeep > a.keepCap) S.fail |= 1u;
        __syncthreads();
        if (S.fail) {
            if (tid == 0) {
                a.outCount[q] = 0;
                if (S.fail & 2u) { atomicAdd(&a.stats[ST_ERR_OVERFLOW], 1ull); atomicMin(a.errQuery, q); }
                else { a.failList[atomicAdd(a.nFail, 1u)] = q; atomicAdd(&a.stats[ST_ERR_CAP], 1ull); }
            }
            continue;
        }

        // ---------------- P5: threshold (QueryMatcher::computeScoreThreshold), selection, final order, output
        auto scoreUncapped = [&](const uint32_t t, const int diag, unsigned long long &res) -> int {
            return diagonalScoreSlow(a.tRes, a.tOff, S.u.prof, S.sub32, qp, useProf ? 1 : 0, 0, Lq, t, diag, &res);
        };
        const long long clk2 = clock64();
        long long clkSel = 0;
        const uint32_t nKeep = S.nKeep;
        if (tid == 0) {
            uint32_t found = 0, sc;
            for (sc = 255; sc > 0; sc--) {
                found += S.hist[sc];
                if (found >= a.maxHits) break;
            }
            S.thrScore = max(sc, a.minDiagScore);
            const uint32_t first = (identity != 0xffffffffu) ? 1u : 0u;
            if (first) { resT[0] = identity; resScore[0] = 65535; resAux[0] = 0; }
            S.nRes = first;
        }
        __syncthreads();
        const uint32_t thrScore = S.thrScore;
        const bool needRescale = thrScore >= 255u;   // >= max-seqs elements at the 8-bit cap: rescoreHits path
        int maxSelf = 0;
        if (needRescale) {
            // scoreIsTruncated (QueryMatcher.cpp:157-171): the cap no longer orders the top class.  rescoreHits (:525-544):
            // uncapped diagonal score minus 255, rescaled to 0..255 against the query's self score; stable re-sort by that
            // value; getResult maps it back with rescaleScore = maxSelf (:407-410).
            if (tid == 0) {
                int best = 0, run = 0;
                for (int p = 0; p < Lq; p++) {
                    const uint32_t qcode = a.q.code[qBase + p];
                    run += (int)S.sub32[qcode * 32u + qcode] + (int)a.q.corr[qBase + p];
                    run = max(run, 0);
                    best = max(best, run);
                }
                S.maxSelf = min(max(best - 255, 1), 65535);
            }
            __syncthreads();
            maxSelf = S.maxSelf;
            const float fltMaxSelf = (float)maxSelf;
            for (uint32_t i = tid; i < nKeep; i += PF_THREADS) {
                uint4 v = keepC[i];
                if ((v.y >> 16) == 255u) {
                    unsigned long long dummy = 0;
                    const uint32_t full = (uint32_t)scoreUncapped(v.x, (int)(int16_t)(uint16_t)(v.y & 0xffffu), dummy);
                    const uint32_t newScore = full - 255u;
                    const float sc = (float)min(newScore, 65535u);
                    const float scaled = __fmul_rn(__fdiv_rn(sc, fltMaxSelf), 255.0f);
                    v.w = (v.w & 0xff00u) | ((uint32_t)(int)__dadd_rn((double)scaled, 0.5) & 0xffu);
                    keepC[i] = v;
                }
            }
            __syncthreads();
        } else {
            // everything strictly above the threshold survives (fewer than maxHits elements by construction)
            for (uint32_t i0 = 0; i0 < nKeep; i0 += PF_THREADS) {
                const uint32_t i = i0 + tid;
                bool sel = false;
                uint4 v = make_uint4(0, 0, 0, 0);
                if (i < nKeep) { v = keepC[i]; sel = (v.y >> 16) > thrScore && v.x != identity; }
                if (sel) {
                    const uint32_t pos = atomicAdd(&S.nRes, 1u);
                    if (pos < a.maxHits) { resT[pos] = v.x; resScore[pos] = (int)(v.y >> 16); resAux[pos] = v.y & 0xffffu; }
                }
            }
            __syncthreads();
        }
        // the threshold class fills what is left in the reference's order: (rescaled score desc,) id & binMask, stream order
        const uint32_t classScore = needRescale ? 255u : thrScore;
        {
            uint32_t c = 0;
            for (uint32_t i = tid; i < nKeep; i += PF_THREADS) {
                const uint4 v = keepC[i];
                c += ((v.y >> 16) == classScore && v.x != identity) ? 1u : 0u;
            }
            for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
            if (lane == 0 && c) atomicAdd(&S.nAt, c);
        }
        __syncthreads();
        const uint32_t nResAbove = min(S.nRes, a.maxHits);
        const uint32_t room = a.maxHits - nResAbove;
        const uint32_t nAt = S.nAt;
        unsigned long long selLimit = ~0ull;             // keys <= selLimit are taken
        auto keyOf = [&](const uint4 &v) -> unsigned long long {
            return ((unsigned long long)(255u - (needRescale ? (v.w & 0xffu) : 255u)) << 47) | ((unsigned long long)(v.x & a.binMask) << 36) |
                   ((unsigned long long)((v.w >> 8) & 0xfu) << 32) | v.z;
        };
        if (nAt > room) {
            clkSel = clock64();
            // radix select of the room-th smallest key (keys are distinct: the running number is)
            if (room == 0) selLimit = 0ull;              // nothing fits; key 0 cannot occur together with room == 0 ... guard below
            else {
                if (tid == 0) { S.selPrefix = 0ull; S.selNeed = room; }
                __syncthreads();
                for (int bit = needRescale ? 54 : 46; bit >= 0; bit--) {
                    if (tid == 0) S.selCount = 0;
                    __syncthreads();
                    const unsigned long long prefix = S.selPrefix;
                    uint32_t c = 0;
                    for (uint32_t i = tid; i < nKeep; i += PF_THREADS) {
                        const uint4 v = keepC[i];
                        if ((v.y >> 16) == classScore && v.x != identity) {
                            const unsigned long long k = keyOf(v);
                            c += ((k >> bit) == (prefix >> bit)) ? 1u : 0u;      // matches the prefix with this bit 0
                        }
                    }
                    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
                    if (lane == 0 && c) atomicAdd(&S.selCount, c);
                    __syncthreads();
                    if (tid == 0) {
                        if (S.selNeed > S.selCount) { S.selNeed -= S.selCount; S.selPrefix = prefix | (1ull << bit); }
                    }
                    __syncthreads();
                }
                selLimit = S.selPrefix;
            }
            clkSel = clock64() - clkSel;
        }
        for (uint32_t i0 = 0; i0 < nKeep; i0 += PF_THREADS) {
            const uint32_t i = i0 + tid;
            bool sel = false;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (i < nKeep) {
                v = keepC[i];
                sel = (v.y >> 16) == classScore && v.x != identity && room > 0 && keyOf(v) <= selLimit;
            }
            if (sel) {
                const uint32_t pos = atomicAdd(&S.nRes, 1u);
                if (pos < a.maxHits) {
                    resT[pos] = v.x;
                    resScore[pos] = needRescale ? (int)(255u + ((v.w & 0xffu) * (uint32_t)maxSelf) / 255u) : (int)(v.y >> 16);
                    resAux[pos] = v.y & 0xffffu;
                }
            }
        }
        __syncthreads();
        const uint32_t nRes = min(S.nRes, a.maxHits);
        // scores that hit the 8-bit cap are replaced by the uncapped diagonal score (QueryMatcher.cpp:410-413)
        if (!needRescale) {
            for (uint32_t r = tid; r < nRes; r += PF_THREADS) {
                if (resScore[r] >= 255 && resScore[r] != 65535) {
                    unsigned long long dummy = 0;
                    resScore[r] = scoreUncapped(resT[r], (int)(int16_t)(uint16_t)resAux[r], dummy);
                }
            }
        }
        __syncthreads();
        // ---------------- final order (score desc, id asc; identity stays first) + coverage pre-filter
        const uint32_t first = (identity != 0xffffffffu) ? 1u : 0u;
        sd_prefilter_hit *outQ = a.out + (size_t)q * a.maxHits;
        for (uint32_t r = tid; r < nRes; r += PF_THREADS) {
            const uint32_t t = resT[r];
            const int s = resScore[r];
            uint32_t rank = 0;
            if (r >= first) {
                rank = first;
                for (uint32_t j = first; j < nRes; j++) {
                    const int sj = resScore[j];
                    const uint32_t tj = resT[j];
                    rank += (sj > s || (sj == s && tj < t)) ? 1u : 0u;
                }
            }
            // coverage pre-filter (Prefiltering.cpp:856-863, Util::canBeCovered Util.cpp:477-494)
            bool pass = true;
            if (a.covThr > 0.0f && (a.covMode == 0 || a.covMode == 2 || a.covMode == 5)) {
                const float ql = (float)Lq;
                const float tl = (float)(a.tOff[t + 1] - a.tOff[t]);
                if (a.covMode == 0) pass = (__fdiv_rn(ql, tl) >= a.covThr) && (__fdiv_rn(tl, ql) >= a.covThr);
                else if (a.covMode == 2) pass = __fdiv_rn(tl, ql) >= a.covThr;
                else pass = __fdiv_rn(fminf(tl, ql), fmaxf(tl, ql)) >= a.covThr;
            }
            outQ[rank].query = q;
            outQ[rank].target = t;
            outQ[rank].score = pass ? s : -1;      // -1 marks "filtered by coverage"; compacted below
            outQ[rank].diag = (int16_t)(uint16_t)resAux[r];
            outQ[rank].pad = 0;
        }
        __syncthreads();
        // compact the coverage-filtered entries in place, keeping order (warp 0)
        if (w == 0) {
            uint32_t nOut = 0;
            for (uint32_t r0 = 0; r0 < nRes; r0 += 32) {
                const uint32_t r = r0 + lane;
                sd_prefilter_hit h;
                bool ok = false;
                if (r < nRes) { h = outQ[r]; ok = h.score >= 0; }
                const uint32_t b = __ballot_sync(0xffffffffu, ok);
                __syncwarp();
                if (ok) outQ[nOut + __popc(b & ltMask)] = h;
                nOut += __popc(b);
                __syncwarp();
            }
            if (lane == 0) {
                a.outCount[q] = nOut;
                const uint32_t eligible = (needRescale ? 0u : (nResAbove - first)) + nAt;
                atomicAdd(&a.stats[ST_PROBES], S.stat[ST_PROBES]);
                atomicAdd(&a.stats[ST_MATCHES], S.stat[ST_MATCHES]);
                atomicAdd(&a.stats[ST_DIAGRES], S.stat[ST_DIAGRES]);
                atomicAdd(&a.stats[ST_HITS], (unsigned long long)nRes);
                atomicAdd(&a.stats[ST_CANDS], S.stat[ST_CANDS]);
                if (eligible + first > nRes) atomicAdd(&a.stats[ST_TRUNC], 1ull);
                if (S.nBound) atomicAdd(&a.stats[ST_CHUNKED], 1ull);
                if (S.nBound >= 2) atomicAdd(&a.stats[ST_MULTICHUNKED], 1ull);
                const long long clk3 = clock64();
                atomicAdd(&a.stats[ST_CYC_P1], (unsigned long long)(clk1 - clk0));
                atomicAdd(&a.stats[ST_CYC_P24], (unsigned long long)(clk2 - clk1));
                atomicAdd(&a.stats[ST_CYC_P5], (unsigned long long)(clk3 - clk2));
                atomicAdd(&a.stats[ST_CYC_SEL], (unsigned long long)clkSel);
                atomicAdd(&a.stats[ST_CYC_P0B], (unsigned long long)S.clkP0b);
                atomicAdd(&a.stats[ST_CYC_STAGEB], (unsigned long long)S.prof1[0]);
                atomicAdd(&a.stats[ST_CYC_DRAIN], (unsigned long long)S.prof1[1]);
                atomicAdd(&a.stats[ST_CYC_BIG], (unsigned long long)S.prof1[2]);
                atomicAdd(&a.stats[ST_CYC_TAIL], (unsigned long long)S.prof1[3]);
                atomicAdd(&a.stats[ST_N_STAGEB], (unsigned long long)S.prof1[4]);
                atomicAdd(&a.stats[ST_N_HITROWS], (unsigned long long)S.prof1[5]);
            }
        }
    }
}
