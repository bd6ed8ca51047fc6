// Streaming variant of the prefilter kernel (included by prefilter.cu inside namespace sd) for target sets whose last-diagonal
// table stays small (sd_prefilter_core picks it by the number of targets): one WARP owns one query and walks its similar-k-mer
// probes in the reference's stream order, 32 probes / 32 index entries per step; the order-dependent double-hit rule is evaluated
// with __match_any_sync inside the step and with a per-warp table in global memory (epoch-tagged, one word per target) between
// steps.  The hit stream is never materialised, so the per-entry cost is the lowest of the two kernels; the table traffic grows
// with the number of targets (every entry is a sector read + write once the tables outgrow the L2), which is where the binned
// kernel (prefilter_kernel.cuh: shared-memory tables) takes over.  Queries that reach the reference's hit-buffer limit
// (QueryMatcher.cpp:43,280) are not finished here: they are listed and re-run by the binned kernel, which restates that path.
//
// Three decoupled stages connected by per-warp FIFO rings in shared memory (stream order = FIFO order):
//   A  32 similar k-mers per step -> occupancy bitmap {bits, rank} (L2 resident) -> non-empty ones to ring 1
//   B  32 buckets per step: bucket row (prefetched by A) -> entries -> (target, diagonal) to ring 2
//   C  32 index hits per step: double-diagonal rule on the per-warp table (prefetched by B)
// then diagonal scores (warp-cooperative), best diagonal per target, --max-seqs cut, final order (all by the owning warp).

constexpr int PS_WARPS = 8;              // warps per CTA
constexpr int PS_CAP2 = 1024;            // staged prefix of the second 3-mer's neighbour row
constexpr int PS_UNROLL = 4;             // probe steps in flight per warp (independent bitmap loads)
constexpr uint32_t PS_Q1CAP = 256;       // ring 1 (non-empty buckets): up to 32*PS_UNROLL pushes between drains
constexpr uint32_t PS_Q2CAP = 128;       // ring 2 (index hits)
constexpr uint32_t PS_QDRAIN = 96;       // a stage runs when its ring holds this many: popped items were prefetched >= 64 pushes ago
constexpr size_t PS_DYN_SMEM = (size_t)PS_WARPS * (PS_Q1CAP + PS_Q2CAP) * (sizeof(uint32_t) + sizeof(uint16_t));
static_assert(PS_QDRAIN - 1 + 32 * PS_UNROLL <= PS_Q1CAP && PS_QDRAIN - 1 + 32 <= PS_Q2CAP, "ring capacity");


// L2 cache policies: the streamed, never re-used data (bucket rows, entries, the per-warp tables) is marked evict-first so
// that it does not push the re-used data (occupancy bitmap, count table, target residues) out of the L2.
__device__ __forceinline__ uint64_t l2Policy(int evictFirst) {
    uint64_t pol;
    if (evictFirst) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2PolicyKeep() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint32_t ldKeep(const uint32_t *p, uint64_t pol) {
    uint32_t v;
    asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ uint32_t ldStream(const uint32_t *p, uint64_t pol) {
    uint32_t v;
    asm volatile("ld.global.cg.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol) : "memory");
    return v;
}
__device__ __forceinline__ uint2 ldStream2(const uint2 *p, uint64_t pol) {
    uint2 v;
    asm volatile("ld.global.cg.L2::cache_hint.v2.u32 {%0, %1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(pol) : "memory");
    return v;
}
__device__ __forceinline__ void stStream(uint32_t *p, uint32_t v, uint64_t pol) {
    asm volatile("st.global.cg.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(pol) : "memory");
}

struct StreamArgs {
    // target side
    const uint8_t *tRes;
    const uint32_t *tOff;
    uint32_t nT;
    const uint2 *bm;           // per 32 k-mers: {occupancy bits, row of the first non-empty k-mer}
    const uint2 *rows;         // per non-empty k-mer: {target, 1<<31 | pos} (single entry) or {start, length}
    const uint2 *ent;          // bucket entries: (target id, first position of the k-mer in that target)
    const uint32_t *e3pair;    // sorted 3-mer neighbour rows: score | index << 16
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
    uint32_t binMask;
    unsigned long long maxDbMatches;   // QueryMatcher.cpp:43: queries that reach it are handed to the binned kernel
    // scratch
    uint32_t *table;           // per warp: nT words
    uint32_t candCap;
    uint64_t *cand;            // per warp: candCap   (t<<32 | diag<<16 | score<<8)
    uint64_t *keep;            // per warp: candCap   (t<<32 | diag<<16 | score<<8), emission order kept
    uint32_t *resT;            // per warp: maxHits
    int32_t *resScore;
    uint16_t *resDiag;
    // output
    sd_prefilter_hit *out;     // nQ * maxHits
    uint32_t *outCount;        // nQ
    unsigned long long *stats; // [0]=probes [1]=matches [2]=diagResidues [3]=hits [4]=candidates [5]=truncated [6]=errors
    uint32_t *nextQuery;
    uint32_t *errQuery;
    uint32_t *failList;        // queries whose candidate list did not fit candCap (re-run with a larger list)
    uint32_t *nFail;
    int evictFirst;            // L2 evict-first policy for the streamed data
    int prefetch;              // bit 0: table words, bit 1: bucket rows
    int debugSkip;             // profiling only (SD_PF_SKIP): bit 0 drops stage C, bit 1 drops stage B, bit 2 drops the bitmap loads -- results are then wrong
};


// UngappedAlignment::computeSingelSequenceScores + scalarDiagonalScoring (UngappedAlignment.cpp:30-43,416-430)
__device__ int diagonalScore(const StreamArgs &a, const int8_t *sSub, uint32_t qBase, int Lq, uint32_t t, int diag,
                             uint32_t &residues) {
    const uint32_t tb = a.tOff[t];
    const int Lt = (int)(a.tOff[t + 1] - tb);
    int qs, ts, len;
    if (diag >= 0) {
        if (diag >= Lq) return 0;
        qs = diag; ts = 0; len = min(Lt, Lq - diag);
    } else {
        if (-diag >= Lt) return 0;
        qs = 0; ts = -diag; len = min(Lt + diag, Lq);
    }
    const uint8_t *qc = a.q.code + qBase + qs;
    const int8_t *qk = a.q.corr + qBase + qs;
    const uint8_t *tc = a.tRes + tb + ts;
    int best = 0, run = 0;
    for (int p = 0; p < len; p++) {
        run += (int)sSub[qc[p] * AA + tc[p]] + (int)qk[p];
        run = max(run, 0);
        best = max(best, run);
    }
    residues += (uint32_t)len;
    return best;
}

// same score, computed by a whole warp: lane l takes residues l, l+32, ...; with S the prefix sums of the per-residue scores,
// run_i = S_i - min(0, min_{j<=i} S_j) and the result is max_i run_i
__device__ __forceinline__ int diagonalScoreWarp(const StreamArgs &a, const int8_t *sSub, uint32_t qBase, int Lq, uint32_t t,
                                                 int diag, uint32_t lane, unsigned long long &residues) {
    const uint32_t tb = __ldg(&a.tOff[t]);
    const int Lt = (int)(__ldg(&a.tOff[t + 1]) - tb);
    int qs, ts, len;
    if (diag >= 0) {
        if (diag >= Lq) return 0;
        qs = diag; ts = 0; len = min(Lt, Lq - diag);
    } else {
        if (-diag >= Lt) return 0;
        qs = 0; ts = -diag; len = min(Lt + diag, Lq);
    }
    const uint8_t *qc = a.q.code + qBase + qs;
    const int8_t *qk = a.q.corr + qBase + qs;
    const uint8_t *tc = a.tRes + tb + ts;
    int best = 0, carrySum = 0, carryMin = 0;
    for (int p0 = 0; p0 < len; p0 += 32) {
        const int p = p0 + (int)lane;
        int x = 0;
        if (p < len) x = (int)sSub[__ldg(&qc[p]) * AA + __ldg(&tc[p])] + (int)__ldg(&qk[p]);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= (uint32_t)d) x += y;
        }
        x += carrySum;
        int m = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, m, d);
            if (lane >= (uint32_t)d) m = min(m, y);
        }
        m = min(m, carryMin);
        best = max(best, x - m);
        carrySum = __shfl_sync(0xffffffffu, x, 31);
        carryMin = __shfl_sync(0xffffffffu, m, 31);
    }
    if (lane == 0) residues += (unsigned long long)len;
    return __reduce_max_sync(0xffffffffu, best);
}

__global__ void __launch_bounds__(PS_WARPS * 32, 4) prefilterStreamKernel(StreamArgs a) {
    __shared__ uint16_t sX2[PS_WARPS][PS_CAP2];
    __shared__ uint32_t sHist[PS_WARPS][256];
    __shared__ int8_t sSub[AA * AA];
    extern __shared__ __align__(16) uint8_t sDyn[];    // stage rings: [warp][Q1c | Q2t | Q1p | Q2d]

    for (int i = threadIdx.x; i < AA * AA; i += blockDim.x) sSub[i] = a.diagSub[i];
    __syncthreads();

    const uint32_t lane = laneId();
    const uint32_t wInBlock = threadIdx.x >> 5;
    const uint32_t warp = blockIdx.x * PS_WARPS + wInBlock;
    const uint32_t ltMask = (1u << lane) - 1u;
    uint32_t *tab = a.table + (size_t)warp * a.nT;
    uint64_t *cand = a.cand + (size_t)warp * a.candCap;
    uint64_t *keep = a.keep + (size_t)warp * a.candCap;
    uint32_t *resT = a.resT + (size_t)warp * a.maxHits;
    int32_t *resScore = a.resScore + (size_t)warp * a.maxHits;
    uint16_t *resDiag = a.resDiag + (size_t)warp * a.maxHits;
    uint16_t *X2 = sX2[wInBlock];
    uint32_t *hist = sHist[wInBlock];
    uint32_t *Q1c = reinterpret_cast<uint32_t *>(sDyn) + (size_t)wInBlock * (PS_Q1CAP + PS_Q2CAP);
    uint32_t *Q2t = Q1c + PS_Q1CAP;
    uint16_t *Q1p = reinterpret_cast<uint16_t *>(sDyn + (size_t)PS_WARPS * (PS_Q1CAP + PS_Q2CAP) * sizeof(uint32_t)) + (size_t)wInBlock * (PS_Q1CAP + PS_Q2CAP);
    uint16_t *Q2d = Q1p + PS_Q1CAP;

    const uint64_t pol = l2Policy(a.evictFirst);
    const uint64_t polKeep = a.evictFirst ? l2PolicyKeep() : l2Policy(0);
    uint32_t epoch = 0;   // table words start zeroed; live epochs are 1..255

    for (;;) {
        uint32_t qi = 0;
        if (lane == 0) qi = atomicAdd(a.nextQuery, 1u);
        qi = __shfl_sync(0xffffffffu, qi, 0);
        if (qi >= a.nQ) break;
        // per-query counters: added to the global statistics only when the query completes (a query whose candidate
        // list overflows is re-run by a second launch and counted there)
        unsigned long long stProbes = 0, stMatches = 0, stDiagRes = 0, stHits = 0, stCands = 0, stTrunc = 0;
        const uint32_t q = a.order[qi];
        const uint32_t qBase = a.q.off[q];
        const int Lq = (int)(a.q.off[q + 1] - qBase);
        const uint32_t identity = a.sameDb ? (a.queryFirst + q) : 0xffffffffu;

        // two epochs per query: E for the hit stream, E+1 for the best-diagonal pass
        if (epoch + 2 > 255) {
            for (uint32_t i = lane; i < a.nT; i += 32) __stcg(&tab[i], 0u);
            __syncwarp();
            epoch = 0;
        }
        const uint32_t E = epoch + 1;
        epoch += 2;

        uint32_t nCand = 0;
        bool candOverflow = false;

        // ---------------- hit stream: positions -> similar k-mers -> buckets -> double-hit rule
        // Three stages, decoupled by two small per-warp FIFO rings in shared memory so that every stage runs with full
        // lanes and every HBM access was prefetched into L2 one stage earlier (stream order is the FIFO order):
        //   A  32 similar k-mers per step -> occupancy bitmap + rank (L2 resident) -> non-empty ones to ring 1
        //   B  32 buckets per step: compact offset row (prefetched by A) -> entries -> (target, diagonal) to ring 2
        //   C  32 index hits per step: double-diagonal rule on the per-warp table (prefetched by B)
        uint32_t q1h = 0, q1n = 0, q2h = 0, q2n = 0;

        auto stageC = [&](const uint32_t n) {
            __syncwarp();
            const bool validE = lane < n;
            uint32_t t = 0x80000000u | lane;     // unique dummy for idle lanes
            uint32_t diag16 = 0;
            if (validE) {
                const uint32_t slot = (q2h + lane) & (PS_Q2CAP - 1);
                t = Q2t[slot];
                diag16 = Q2d[slot];
            }
            q2h += n;
            q2n -= n;
            if (a.debugSkip & 1) return;
            const uint32_t d8 = diag16 & 0xffu;
            // --- CacheFriendlyOperations::findDuplicates, restated per target in stream order
            const uint32_t grp = __match_any_sync(0xffffffffu, t);
            uint32_t old = 0;
            if (validE) old = ldStream(&tab[t], pol);
            if ((old >> 24) != E) old = 0;
            const uint32_t lowerG = grp & ltMask;
            const int prevLane = lowerG ? (31 - __clz(lowerG)) : -1;
            uint32_t prevD8 = __shfl_sync(0xffffffffu, d8, prevLane >= 0 ? prevLane : (int)lane);
            if (prevLane < 0) prevD8 = (old >> 16) & 0xffu;
            const bool emit = validE && (d8 == prevD8);
            const uint32_t emitBallot = __ballot_sync(0xffffffffu, emit);
            const uint32_t grpE = grp & emitBallot;
            const uint32_t lowerE = grpE & ltMask;
            const int prevELane = lowerE ? (31 - __clz(lowerE)) : -1;
            uint32_t prevED8 = __shfl_sync(0xffffffffu, d8, prevELane >= 0 ? prevELane : (int)lane);
            bool prevEExists = true;
            if (prevELane < 0) { prevED8 = old & 0xffu; prevEExists = ((old >> 8) & 1u) != 0; }
            const bool isCand = emit && !(prevEExists && prevED8 == d8);
            const int lastELane = grpE ? (31 - __clz(grpE)) : -1;
            const uint32_t lastED8 = __shfl_sync(0xffffffffu, d8, lastELane >= 0 ? lastELane : (int)lane);
            if (validE && (grp >> lane) == 1u) {
                uint32_t w = (E << 24) | (d8 << 16);
                if (lastELane >= 0) w |= (1u << 8) | lastED8;
                else w |= old & 0x1ffu;
                stStream(&tab[t], w, pol);
            }
            const uint32_t candBallot = __ballot_sync(0xffffffffu, isCand);
            if (isCand) {
                const uint32_t pos = nCand + __popc(candBallot & ltMask);
                if (pos < a.candCap) cand[pos] = ((uint64_t)t << 32) | ((uint64_t)diag16 << 16);
            }
            nCand += __popc(candBallot);
        };

        auto stageB = [&](const uint32_t n) {
            __syncwarp();
            uint32_t rx = 0, ry = 0, m = 0, pq = 0;
            if (a.debugSkip & 2) { q1h += n; q1n -= n; return; }
            if (lane < n) {
                const uint32_t slot = (q1h + lane) & (PS_Q1CAP - 1);
                const uint2 row = ldStream2(&a.rows[Q1c[slot]], pol);
                pq = Q1p[slot];
                rx = row.x;
                ry = row.y;
                m = (ry >> 31) ? 1u : ry;                // >= 1: only non-empty buckets are queued
            }
            q1h += n;
            q1n -= n;
            uint32_t totalM;
            const uint32_t exclM = warpExclusiveScan(m, totalM);
            stMatches += (lane == 0) ? totalM : 0;
            for (uint32_t ec = 0; ec < totalM; ec += 32) {
                const uint32_t eidx = ec + lane;
                const bool validE = eidx < totalM;
                const uint32_t ownerM = warpOwnerOfSlot(exclM, m != 0, ec, lane);
                const uint32_t oExcl = __shfl_sync(0xffffffffu, exclM, ownerM);
                const uint32_t oX = __shfl_sync(0xffffffffu, rx, ownerM);
                const uint32_t oY = __shfl_sync(0xffffffffu, ry, ownerM);
                const uint32_t oP = __shfl_sync(0xffffffffu, pq, ownerM);
                if (validE) {
                    uint2 en = make_uint2(oX, oY & 0xffffu);
                    if (!(oY >> 31)) en = ldStream2(&a.ent[oX + (eidx - oExcl)], pol);
                    if (a.prefetch & 1) prefetchL2(&tab[en.x]);
                    const uint32_t slot = (q2h + q2n + lane) & (PS_Q2CAP - 1);
                    Q2t[slot] = en.x;
                    Q2d[slot] = (uint16_t)(oP - en.y);
                }
                q2n += min(32u, totalM - ec);
                while (q2n >= PS_QDRAIN) stageC(32u);
            }
        };

        for (int p0 = 0; p0 + SPAN <= Lq; p0 += 32) {
            // ---- set-up of 32 positions at once (one per lane): thresholds, 3-mer rows, reachable prefix of row 2
            int thrL = -1, top2L = 0;
            uint32_t i1L = 0, i2L = 0, nb2L = 0;
            {
                const int pl = p0 + (int)lane;
                if (pl + SPAN <= Lq) thrL = a.q.thr[qBase + pl];
                if (thrL >= 0) {
                    const uint8_t *qc = a.q.code + qBase + pl;
                    i1L = qc[0] + 20u * qc[1] + 400u * qc[3];
                    i2L = qc[5] + 20u * qc[8] + 400u * qc[9];
                    const int top1 = (int)(int16_t)(__ldg(&a.e3pair[(size_t)i1L * N3]) & 0xffffu);
                    top2L = (int)(int16_t)(__ldg(&a.e3pair[(size_t)i2L * N3]) & 0xffffu);
                    if (top1 < thrL - top2L) thrL = -1;           // KmerGenerator.cpp:129-137: nothing reaches the threshold
                    else {
                        nb2L = countGE(a.e3cnt, i2L, thrL - top1);
                        if (nb2L == 0) thrL = -1;
                        else {
                            const uint32_t *x2 = a.e3pair + (size_t)i2L * N3;
                            for (uint32_t c = 0; c < nb2L && c < (uint32_t)PS_CAP2; c += 32) prefetchL2(&x2[c]);
                        }
                    }
                }
            }
            uint32_t active = __ballot_sync(0xffffffffu, thrL >= 0);
            while (active) {
                const int src = __ffs(active) - 1;
                active &= active - 1;
                const int p = p0 + src;
                const int thr = __shfl_sync(0xffffffffu, thrL, src);
                const uint32_t i1 = __shfl_sync(0xffffffffu, i1L, src);
                const uint32_t i2 = __shfl_sync(0xffffffffu, i2L, src);
                const uint32_t nb2 = __shfl_sync(0xffffffffu, nb2L, src);
                const int cutoff1 = thr - __shfl_sync(0xffffffffu, top2L, src);
                const uint32_t *r1 = a.e3pair + (size_t)i1 * N3;
                const uint32_t *x2 = a.e3pair + (size_t)i2 * N3;
                // stage the prefix of row 2 that any first-row element can reach
                __syncwarp();
                for (uint32_t c = lane; c < nb2 && c < (uint32_t)PS_CAP2; c += 32) X2[c] = (uint16_t)(__ldg(&x2[c]) >> 16);
                __syncwarp();

                for (uint32_t a0 = 0; a0 < N3; a0 += 32) {
                    const uint32_t ai = a0 + lane;
                    const uint32_t pr1 = (ai < N3) ? __ldg(&r1[ai]) : 0x8000u;
                    const int sa = (int)(int16_t)(pr1 & 0xffffu);
                    const bool okA = sa >= cutoff1;
                    const uint32_t okBallot = __ballot_sync(0xffffffffu, okA);
                    if (okBallot == 0) break;
                    uint32_t nb = 0;
                    uint32_t xa = 0;
                    if (okA) {
                        xa = pr1 >> 16;
                        nb = countGE(a.e3cnt, i2, thr - sa);      // KmerGenerator.cpp:201: second-row elements with s2 >= thr - s1
                    }
                    // nb is non-increasing over the lanes (row 1 is sorted), zeros only at the tail
                    uint32_t totalProbes;
                    const uint32_t exclA = warpExclusiveScan(nb, totalProbes);
                    stProbes += (lane == 0) ? totalProbes : 0;

                    for (uint32_t pc = 0; pc < totalProbes; pc += 32 * PS_UNROLL) {
                        uint32_t w[PS_UNROLL], rk[PS_UNROLL], kmerW[PS_UNROLL];
#pragma unroll
                        for (int j = 0; j < PS_UNROLL; j++) {
                            const uint32_t base = pc + 32u * j;
                            w[j] = 0;
                            rk[j] = 0;
                            kmerW[j] = 0;
                            if (base < totalProbes) {
                                const uint32_t ownerA = warpOwnerOfSlot(exclA, nb != 0, base, lane);
                                const uint32_t ownerExcl = __shfl_sync(0xffffffffu, exclA, ownerA);
                                const uint32_t ownerX = __shfl_sync(0xffffffffu, xa, ownerA);
                                if (base + lane < totalProbes) {
                                    const uint32_t c = base + lane - ownerExcl;
                                    const uint32_t xc = (c < (uint32_t)PS_CAP2) ? (uint32_t)X2[c] : (__ldg(&x2[c]) >> 16);
                                    kmerW[j] = kmerSlot(ownerX, xc);            // KmerGenerator.cpp:207
                                    const uint2 wv = (a.debugSkip & 4) ? make_uint2(kmerW[j] * 2654435761u & 0x11111111u, 0u) : __ldg(&a.bm[kmerW[j] >> 5]);
                                    w[j] = wv.x;
                                    rk[j] = wv.y;
                                }
                            }
                        }
#pragma unroll
                        for (int j = 0; j < PS_UNROLL; j++) {
                            if (pc + 32u * j < totalProbes) {
                                const uint32_t bit = kmerW[j] & 31u;
                                const bool has = ((w[j] >> bit) & 1u) != 0;
                                const uint32_t ci = rk[j] + __popc(w[j] & ((1u << bit) - 1u));
                                if (has && (a.prefetch & 2)) prefetchL2(&a.rows[ci]);
                                const uint32_t hb = __ballot_sync(0xffffffffu, has);
                                if (has) {
                                    const uint32_t slot = (q1h + q1n + __popc(hb & ltMask)) & (PS_Q1CAP - 1);
                                    Q1c[slot] = ci;
                                    Q1p[slot] = (uint16_t)p;
                                }
                                q1n += __popc(hb);
                            }
                        }
                        while (q1n >= PS_QDRAIN) stageB(32u);
                    }
                    if (okBallot != 0xffffffffu) break;
                }
            }
        }
        while (q1n > 0) stageB(min(q1n, 32u));
        while (q2n > 0) stageC(min(q2n, 32u));
        __syncwarp();
        if (nCand > a.candCap) { candOverflow = true; nCand = a.candCap; }
        stCands += (lane == 0) ? nCand : 0;
        __syncwarp();

        // ---------------- diagonal scores (UngappedAlignment::align), capped at 255
        // The whole warp scores one diagonal at a time: lanes take consecutive residues (coalesced reads of query and
        // target), the running max(0, run + s) becomes prefix sums minus their running minimum (warp scans).
        for (uint32_t c0 = 0; c0 < nCand; c0 += 32) {
            const uint32_t c = c0 + lane;
            const uint64_t v = (c < nCand) ? cand[c] : 0ull;
            const uint32_t nHere = min(32u, nCand - c0);
            int myScore = 0;
            for (uint32_t j = 0; j < nHere; j++) {
                const uint64_t vj = __shfl_sync(0xffffffffu, v, j);
                const int sc = diagonalScoreWarp(a, sSub, qBase, Lq, (uint32_t)(vj >> 32), (int)(int16_t)(uint16_t)(vj >> 16), lane, stDiagRes);
                if (lane == j) myScore = sc;
            }
            if (c < nCand) {
                const int sc = min(myScore, 255);
                cand[c] = v | ((uint64_t)sc << 8);
                // best score per target: epoch E+1 words order above every epoch-E word
                atomicMax(&tab[(uint32_t)(v >> 32)], ((E + 1) << 24) | (uint32_t)sc);
            }
        }
        __threadfence_block();
        __syncwarp();

        // ---------------- keepMaxScoreElementOnly: first candidate (stream order) carrying the target's max
        for (uint32_t i = lane; i < 256; i += 32) hist[i] = 0;
        __syncwarp();
        uint32_t nKeep = 0;
        for (uint32_t c0 = 0; c0 < nCand; c0 += 32) {
            const uint32_t c = c0 + lane;
            const bool valid = c < nCand;
            uint64_t v = 0;
            uint32_t t = 0x80000000u | lane;
            if (valid) { v = cand[c]; t = (uint32_t)(v >> 32); }
            const uint32_t s = (uint32_t)(v >> 8) & 0xffu;
            uint32_t w = 0;
            if (valid) w = __ldcg(&tab[t]);
            const bool isMax = valid && ((w & 0x100u) == 0) && ((w & 0xffu) == s);
            const uint32_t grp = __match_any_sync(0xffffffffu, t);
            const uint32_t maxBallot = __ballot_sync(0xffffffffu, isMax);
            const uint32_t grpMax = grp & maxBallot;
            const bool winner = isMax && ((grpMax & ltMask) == 0);
            if (winner) __stcg(&tab[t], w | 0x100u);
            const uint32_t winBallot = __ballot_sync(0xffffffffu, winner);
            if (winner) {
                keep[nKeep + __popc(winBallot & ltMask)] = v;
                atomicAdd(&hist[s], 1u);
            }
            nKeep += __popc(winBallot);
            __syncwarp();
        }

        // ---------------- score threshold (QueryMatcher::computeScoreThreshold) and selection
        uint32_t thrScore = 0;
        if (lane == 0) {
            uint32_t found = 0, sc;
            for (sc = 255; sc > 0; sc--) {
                found += hist[sc];
                if (found >= a.maxHits) break;
            }
            thrScore = max(sc, a.minDiagScore);
        }
        thrScore = __shfl_sync(0xffffffffu, thrScore, 0);
        const bool needRescale = thrScore >= 255u;   // >= max-seqs elements at the 8-bit cap: rescoreHits path below
        uint32_t nRes = 0;
        if (identity != 0xffffffffu) {
            if (lane == 0) { resT[0] = identity; resScore[0] = 65535; resDiag[0] = 0; }
            nRes = 1;
        }
        bool dropped = false;
        if (needRescale) {
            // scoreIsTruncated (QueryMatcher.cpp:157-171): the cap no longer orders the top class.  rescoreHits (:525-544):
            // uncapped diagonal score minus 255, rescaled to 0..255 against the query's self score; stable re-sort by that
            // value; getResult maps it back with rescaleScore = maxSelf (:407-410).  Rare, so plain lane-parallel code.
            // 1. the capped class in the reference's order (bin-major, arrival order inside a bin) -> cand[] (free by now)
            uint32_t nL = 0;
            for (uint32_t bin = 0; bin <= a.binMask; bin++) {
                for (uint32_t c0 = 0; c0 < nKeep; c0 += 32) {
                    const uint32_t c = c0 + lane;
                    uint64_t v = 0;
                    bool sel = false;
                    if (c < nKeep) {
                        v = keep[c];
                        sel = ((uint32_t)(v >> 8) & 0xffu) == 255u && ((uint32_t)(v >> 32) & a.binMask) == bin;
                    }
                    const uint32_t b = __ballot_sync(0xffffffffu, sel);
                    if (sel) cand[nL + __popc(b & ltMask)] = v & ~0xffffull;      // (target, diagonal), score byte cleared
                    nL += __popc(b);
                }
            }
            __syncwarp();
            // 2. self score of the query on the main diagonal (scoreSingleSequence(query, 0, 0)), uncapped
            int maxSelf = 0;
            if (lane == 0) {
                const uint8_t *qc = a.q.code + qBase;
                const int8_t *qk = a.q.corr + qBase;
                int best = 0, run = 0;
                for (int p = 0; p < Lq; p++) {
                    run += (int)sSub[qc[p] * AA + qc[p]] + (int)qk[p];
                    run = max(run, 0);
                    best = max(best, run);
                }
                maxSelf = min(max(best - 255, 1), 65535);
            }
            maxSelf = __shfl_sync(0xffffffffu, maxSelf, 0);
            const float fltMaxSelf = (float)maxSelf;
            // 3. rescaled 8-bit scores + their histogram
            for (uint32_t i = lane; i < 256; i += 32) hist[i] = 0;
            __syncwarp();
            for (uint32_t c = lane; c < nL; c += 32) {
                const uint64_t v = cand[c];
                uint32_t dummy = 0;
                const uint32_t full = (uint32_t)diagonalScore(a, sSub, qBase, Lq, (uint32_t)(v >> 32), (int)(int16_t)(uint16_t)(v >> 16), dummy);
                const uint32_t newScore = full - 255u;
                const float sc = (float)min(newScore, 65535u);
                const float scaled = __fmul_rn(__fdiv_rn(sc, fltMaxSelf), 255.0f);
                const uint32_t cnt = (uint32_t)(int)__dadd_rn((double)scaled, 0.5) & 0xffu;
                cand[c] = v | (uint64_t)cnt;
                atomicAdd(&hist[cnt], 1u);
            }
            __syncwarp();
            // 4. stable selection by rescaled score, highest first
            for (int s2 = 255; s2 >= 0; s2--) {
                if (hist[s2] == 0) continue;
                for (uint32_t c0 = 0; c0 < nL; c0 += 32) {
                    const uint32_t c = c0 + lane;
                    uint64_t v = 0;
                    bool sel = false;
                    if (c < nL) {
                        v = cand[c];
                        sel = (uint32_t)(v & 0xffu) == (uint32_t)s2 && (uint32_t)(v >> 32) != identity;
                    }
                    const uint32_t b = __ballot_sync(0xffffffffu, sel);
                    if (sel) {
                        const uint32_t pos = nRes + __popc(b & ltMask);
                        if (pos < a.maxHits) {
                            resT[pos] = (uint32_t)(v >> 32);
                            resScore[pos] = (int)(255u + ((uint32_t)s2 * (uint32_t)maxSelf) / 255u);
                            resDiag[pos] = (uint16_t)(v >> 16);
                        }
                    }
                    nRes += __popc(b);
                    if (nRes > a.maxHits) { dropped = true; nRes = a.maxHits; }
                }
            }
            __syncwarp();
        } else {
        // everything strictly above the threshold survives
        for (uint32_t c0 = 0; c0 < nKeep; c0 += 32) {
            const uint32_t c = c0 + lane;
            uint64_t v = 0;
            bool sel = false;
            if (c < nKeep) {
                v = keep[c];
                const uint32_t s = (uint32_t)(v >> 8) & 0xffu;
                sel = s > thrScore && (uint32_t)(v >> 32) != identity;
            }
            const uint32_t b = __ballot_sync(0xffffffffu, sel);
            if (sel) {
                const uint32_t pos = nRes + __popc(b & ltMask);
                if (pos < a.maxHits) {
                    resT[pos] = (uint32_t)(v >> 32);
                    resScore[pos] = (int)((v >> 8) & 0xffu);
                    resDiag[pos] = (uint16_t)(v >> 16);
                }
            }
            nRes += __popc(b);
        }
        nRes = min(nRes, a.maxHits);
        // the threshold class fills what is left, in (id & binMask, stream order) order
        for (uint32_t bin = 0; bin <= a.binMask; bin++) {
            for (uint32_t c0 = 0; c0 < nKeep; c0 += 32) {
                const uint32_t c = c0 + lane;
                uint64_t v = 0;
                bool sel = false;
                if (c < nKeep) {
                    v = keep[c];
                    const uint32_t s = (uint32_t)(v >> 8) & 0xffu;
                    const uint32_t t = (uint32_t)(v >> 32);
                    sel = s == thrScore && t != identity && (t & a.binMask) == bin;
                }
                const uint32_t b = __ballot_sync(0xffffffffu, sel);
                if (sel) {
                    const uint32_t pos = nRes + __popc(b & ltMask);
                    if (pos < a.maxHits) {
                        resT[pos] = (uint32_t)(v >> 32);
                        resScore[pos] = (int)((v >> 8) & 0xffu);
                        resDiag[pos] = (uint16_t)(v >> 16);
                    }
                }
                nRes += __popc(b);
                if (nRes > a.maxHits) { dropped = true; nRes = a.maxHits; }
            }
        }
        }
        __syncwarp();
        if (dropped) stTrunc += (lane == 0) ? 1 : 0;
        // scores that hit the 8-bit cap are replaced by the uncapped diagonal score (QueryMatcher.cpp:410-413)
        for (uint32_t r = lane; r < nRes && !needRescale; r += 32) {
            if (resScore[r] >= 255 && resScore[r] != 65535) {
                uint32_t dummy = 0;
                resScore[r] = diagonalScore(a, sSub, qBase, Lq, resT[r], (int)(int16_t)resDiag[r], dummy);
            }
        }
        __syncwarp();
        stHits += (lane == 0) ? nRes : 0;

        // ---------------- final order (score desc, id asc; identity stays first) + coverage pre-filter
        const uint32_t first = (identity != 0xffffffffu) ? 1u : 0u;
        sd_prefilter_hit *outQ = a.out + (size_t)q * a.maxHits;
        uint32_t nOut = 0;
        // rank sort: rank = number of elements that come before
        for (uint32_t r0 = 0; r0 < nRes; r0 += 32) {
            const uint32_t r = r0 + lane;
            uint32_t rank = 0xffffffffu;
            uint32_t t = 0;
            int s = 0;
            uint16_t dg = 0;
            if (r < nRes) {
                t = resT[r]; s = resScore[r]; dg = resDiag[r];
                if (r < first) rank = 0;
                else {
                    rank = first;
                    for (uint32_t j = first; j < nRes; j++) {
                        const int sj = resScore[j];
                        const uint32_t tj = resT[j];
                        rank += (sj > s || (sj == s && tj < t)) ? 1u : 0u;
                    }
                }
            }
            // coverage pre-filter (Prefiltering.cpp:856-863, Util::canBeCovered Util.cpp:477-494)
            bool pass = r < nRes;
            if (pass && a.covThr > 0.0f && (a.covMode == 0 || a.covMode == 2 || a.covMode == 5)) {
                const float ql = (float)Lq;
                const float tl = (float)(a.tOff[t + 1] - a.tOff[t]);
                if (a.covMode == 0) pass = (__fdiv_rn(ql, tl) >= a.covThr) && (__fdiv_rn(tl, ql) >= a.covThr);
                else if (a.covMode == 2) pass = __fdiv_rn(tl, ql) >= a.covThr;
                else pass = __fdiv_rn(fminf(tl, ql), fmaxf(tl, ql)) >= a.covThr;
            }
            if (r < nRes) {
                // park (rank, pass) in the scratch result arrays' tail: write straight to the output slot by rank
                outQ[rank].query = q;
                outQ[rank].target = t;
                outQ[rank].score = pass ? s : -1;      // -1 marks "filtered by coverage"; compacted below
                outQ[rank].diag = (int16_t)dg;
                outQ[rank].pad = 0;
            }
        }
        __syncwarp();
        __threadfence_block();
        // compact the coverage-filtered entries in place, keeping order
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
        // diagResidues is accumulated per lane
        for (int d = 16; d > 0; d >>= 1) stDiagRes += __shfl_down_sync(0xffffffffu, stDiagRes, d);
        if (lane == 0) {
            if (stMatches >= a.maxDbMatches) {
                // the reference's hit buffer overflows for this query (QueryMatcher.cpp:280): the binned kernel restates that path
                a.outCount[q] = 0;
                a.failList[atomicAdd(a.nFail, 1u)] = q | 0x80000000u;
            } else if (candOverflow) {
                a.outCount[q] = 0;
                a.failList[atomicAdd(a.nFail, 1u)] = q;
                atomicAdd(&a.stats[ST_ERR_CAP], 1ull);
                atomicMin(a.errQuery, q);
            } else {
                a.outCount[q] = nOut;
                atomicAdd(&a.stats[ST_PROBES], stProbes);
                atomicAdd(&a.stats[ST_MATCHES], stMatches);
                atomicAdd(&a.stats[ST_DIAGRES], stDiagRes);
                atomicAdd(&a.stats[ST_HITS], stHits);
                atomicAdd(&a.stats[ST_CANDS], stCands);
                if (stTrunc) atomicAdd(&a.stats[ST_TRUNC], stTrunc);
            }
        }
        __syncwarp();
    }
}
