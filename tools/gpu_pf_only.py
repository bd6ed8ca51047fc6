#!/usr/bin/env python3
"""Prefilter only on N x G synthetic proteomes all-vs-all (no oracle): kernel time, counters, algorithmic GB/s.
usage: gpu_pf_only.py N G [reps] [query_genes]   (query_genes: only the first so many genes are queries, e.g. one rank's block)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spacedust_b200 as sd  # noqa: E402
from spacedust_b200 import synth  # noqa: E402


def main():
    n, g = int(sys.argv[1]), int(sys.argv[2])
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    prot = synth.generate(n, g)
    buf, off, _ = synth.pack(prot)
    ctx = sd.Context(0)
    gp = sd.default_prefilter_params(same_db=1)
    t0 = time.time()
    idx = sd.TargetIndex(ctx, buf, off, gp)
    print("index build %.3fs entries %d" % (time.time() - t0, idx.entries), flush=True)
    nqg = int(sys.argv[4]) if len(sys.argv) > 4 else len(off) - 1
    batch = sd.QueryBatch(ctx, buf[:int(off[nqg])], off[:nqg + 1])
    for rep in range(reps):
        hits, st = sd.prefilter_batch(ctx, idx, batch, gp)
        algo = 8 * st["probes"] + 6 * st["matches"] + st["diag_residues"] + 21 * st["query_residues"] + 10 * st["hits"]
        print("kernel %.2f ms  total %.2f ms  algorithmic %.2f GB -> %.1f GB/s  hits %d truncated %d"
              % (st["kernel_ms"], st["total_ms"], algo / 1e9, algo / 1e9 / (st["kernel_ms"] / 1e3), len(hits), st["truncated"]), flush=True)


if __name__ == "__main__":
    main()
