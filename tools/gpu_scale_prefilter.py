#!/usr/bin/env python3
"""Scale check: prefilter on N x G synthetic proteomes all-vs-all; a seeded sample of queries is verified against the oracle."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
import spacedust_b200 as sd  # noqa: E402
from spacedust_b200 import synth  # noqa: E402


def main():
    n, g, nsample = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    t0 = time.time()
    prot = synth.generate(n, g)
    buf, off, _ = synth.pack(prot)
    nq = len(off) - 1
    print("generated %d proteins, %d residues in %.1fs" % (nq, int(off[-1]), time.time() - t0), flush=True)
    ctx = sd.Context(0)
    gp = sd.default_prefilter_params(same_db=1)
    t0 = time.time()
    idx = sd.TargetIndex(ctx, buf, off, gp)
    print("index build %.2fs entries %d" % (time.time() - t0, idx.entries), flush=True)
    for rep in range(2):
        t0 = time.time()
        hits, st = sd.prefilter(ctx, idx, buf, off, gp)
        print("prefilter wall %.2fs" % (time.time() - t0), st, len(hits), flush=True)
    algo = 8 * st["probes"] + 6 * st["matches"] + st["diag_residues"] + 21 * st["query_residues"] + 10 * st["hits"]
    print("algorithmic bytes %.3f GB -> %.1f GB/s over kernel time" % (algo / 1e9, algo / 1e9 / (st["kernel_ms"] / 1e3)))
    if nsample > 0:
        rng = np.random.default_rng(1)
        sample = np.sort(rng.choice(nq, size=min(nsample, nq), replace=False))
        qb, qo = helpers.pack_seqs([bytes(buf[int(off[i]):int(off[i + 1])]).decode() for i in sample])
        t0 = time.time()
        op = helpers.default_pref_params(sameDB=0, binSize=2)
        ohits, ost = helpers.orc_prefilter(buf, off, qb, qo, op)
        print("oracle on %d sampled queries: %.1fs" % (len(sample), time.time() - t0), ost, flush=True)
        bad = 0
        for j, q in enumerate(sample):
            o = ohits[ohits["query"] == j]
            g_ = hits[hits["query"] == q]
            g_ = g_[g_["target"] != q]           # oracle ran without the identity rule for the sampled subset
            # identity consumes one --max-seqs slot on the GPU side: compare only when no truncation happened
            o2 = o[o["target"] != q]
            same = len(o2) == len(g_) and all(np.array_equal(o2[f], g_[f]) for f in ("target", "score", "diag"))
            if not same:
                bad += 1
                if bad < 5:
                    print(" DIFF query", q, len(o2), len(g_))
        print("sampled queries differing: %d of %d (differences are expected only for truncated queries)" % (bad, len(sample)))


if __name__ == "__main__":
    main()
