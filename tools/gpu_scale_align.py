#!/usr/bin/env python3
"""Scale check: prefilter + align on N x G synthetic proteomes all-vs-all; a sample of queries verified against the oracle."""
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
    prot = synth.generate(n, g)
    buf, off, _ = synth.pack(prot)
    nq = len(off) - 1
    print("generated %d proteins, %d residues" % (nq, int(off[-1])), flush=True)
    ctx = sd.Context(0)
    gp = sd.default_prefilter_params(same_db=1)
    t0 = time.time()
    idx = sd.TargetIndex(ctx, buf, off, gp)
    print("index build %.2fs" % (time.time() - t0), flush=True)
    t0 = time.time()
    hits, st = sd.prefilter(ctx, idx, buf, off, gp)
    print("prefilter wall %.2fs kernel %.1f ms hits %d" % (time.time() - t0, st["kernel_ms"], len(hits)), flush=True)
    ap = sd.default_align_params(same_db=1)
    for rep in range(2):
        t0 = time.time()
        al, pool, ast = sd.align(ctx, idx, buf, off, hits["query"], hits["target"], ap)
        print("align wall %.2fs" % (time.time() - t0), ast, len(al), flush=True)
    cells = ast["cells_forward"] + ast["cells_reverse"]
    print("SW GCUPS (fwd+rev cells / sw kernel time): %.1f" % (cells / 1e9 / (ast["sw_ms"] / 1e3)))
    if nsample > 0:
        rng = np.random.default_rng(1)
        sample = np.sort(rng.choice(nq, size=min(nsample, nq), replace=False))
        sel = np.isin(hits["query"], sample)
        remap = {int(q): i for i, q in enumerate(sample)}
        pq = np.array([remap[int(q)] for q in hits["query"][sel]], dtype=np.uint32)
        pt = hits["target"][sel]
        qb, qo = helpers.pack_seqs([bytes(buf[int(off[i]):int(off[i + 1])]).decode() for i in sample])
        tkeys = np.arange(nq, dtype=np.uint32)
        t0 = time.time()
        oal, obt, ost = helpers.orc_align(buf, off, tkeys, qb, qo, sample.astype(np.uint32), pq, pt, helpers.default_align_params(sameDB=1))
        print("oracle align on %d sampled queries (%d pairs): %.1fs" % (len(sample), len(pq), time.time() - t0), ost, flush=True)
        g_ = al[np.isin(al["query"], sample)]
        ok = len(g_) == len(oal)
        if ok:
            for fo, fg in (("target", "target"), ("score", "score"), ("bits", "bits"), ("qStart", "q_start"), ("qEnd", "q_end"),
                           ("tStart", "t_start"), ("tEnd", "t_end"), ("alnLen", "aln_len"), ("seqId", "seq_id"), ("evalue", "evalue")):
                if not np.array_equal(oal[fo], g_[fg]):
                    ok = False
                    print(" field differs:", fo)
        print("sample parity:", "OK" if ok else "MISMATCH", len(g_), len(oal))


if __name__ == "__main__":
    main()
