#!/usr/bin/env python3
"""First-contact GPU diagnostic: CUDA prefilter vs oracle on small seeded inputs; prints every difference
instead of stopping at the first (GPU round trips are expensive)."""
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
    n, g = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (3, 120)
    prot = synth.generate(n, g, seed=7)
    buf, off, _ = synth.pack(prot)
    t0 = time.time()
    op = helpers.default_pref_params()
    ohits, ost, omask = helpers.orc_prefilter(buf, off, buf, off, op, want_masked=True)
    print("oracle: %.1fs" % (time.time() - t0), ost, len(ohits))
    ctx = sd.Context(0)
    gp = sd.default_prefilter_params(same_db=1, bin_size=2)
    t0 = time.time()
    idx = sd.TargetIndex(ctx, buf, off, gp)
    print("index build %.2fs entries %d" % (time.time() - t0, idx.entries))
    gmask = idx.masked_residues()
    print("mask diff positions:", int((gmask != omask).sum()), "masked oracle", int((omask == 20).sum()), "gpu", int((gmask == 20).sum()))
    t0 = time.time()
    ghits, gst = sd.prefilter(ctx, idx, buf, off, gp)
    print("gpu prefilter %.2fs" % (time.time() - t0), gst, len(ghits))
    for a, b in (("probes", "probes"), ("matches", "matches"), ("candidates", "candidates"), ("diagResidues", "diag_residues"), ("hits", "hits")):
        print("  %-12s oracle %12d gpu %12d %s" % (a, ost[a], gst[b], "OK" if ost[a] == gst[b] else "DIFF"))
    nq = len(off) - 1
    bad = 0
    for q in range(nq):
        o = ohits[ohits["query"] == q]
        g_ = ghits[ghits["query"] == q]
        same = len(o) == len(g_) and all(np.array_equal(o[f], g_[f]) for f in ("target", "score", "diag"))
        if not same:
            bad += 1
            if bad <= 8:
                print("query", q, "L", int(off[q + 1] - off[q]))
                print("   oracle", [(int(x["target"]), int(x["score"]), int(x["diag"])) for x in o[:12]])
                print("   gpu   ", [(int(x["target"]), int(x["score"]), int(x["diag"])) for x in g_[:12]])
    print("queries differing: %d of %d" % (bad, nq))
    # second call on the same context/index must give the same answer (epoch / scratch reuse)
    ghits2, _ = sd.prefilter(ctx, idx, buf, off, gp)
    print("repeat identical:", len(ghits2) == len(ghits) and all(np.array_equal(ghits[f], ghits2[f]) for f in ("query", "target", "score", "diag")))
    print("launches", ctx.launches)


if __name__ == "__main__":
    main()
