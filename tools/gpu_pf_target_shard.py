#!/usr/bin/env python3
"""Ablation for DESIGN.md section 5: the prefilter stage of BASELINE's target-sharded layout, measured on ONE GPU acting as rank 0
of `world`: index over 1/world of the targets, ALL queries probed against it (what every GPU would do before the all-gather of
per-query hit lists), next to the layout that is built (whole index, 1/world of the queries).
usage: gpu_pf_target_shard.py N G world"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spacedust_b200 as sd  # noqa: E402
from spacedust_b200 import synth  # noqa: E402


def run(ctx, tbuf, toff, qbuf, qoff, same_db, tag):
    gp = sd.default_prefilter_params(same_db=same_db)
    idx = sd.TargetIndex(ctx, tbuf, toff, gp)
    batch = sd.QueryBatch(ctx, qbuf, qoff)
    for rep in range(2):
        hits, st = sd.prefilter_batch(ctx, idx, batch, gp, 0)
    algo = 8 * st["probes"] + 6 * st["matches"] + st["diag_residues"] + 21 * st["query_residues"] + 10 * st["hits"]
    print("%s: targets %d entries %d queries %d | kernel %.2f ms total %.2f ms | probes %.3g index entries %.3g hits %d (%.1f MB as 12-byte records) "
          "| algorithmic %.2f GB" % (tag, len(toff) - 1, idx.entries, len(qoff) - 1, st["kernel_ms"], st["total_ms"], st["probes"], st["matches"],
                                     len(hits), 12 * len(hits) / 1e6, algo / 1e9), flush=True)
    batch.close()
    idx.close()


def main():
    n, g, world = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    prot = synth.generate(n, g)
    buf, off, _ = synth.pack(prot)
    nt = len(off) - 1
    ctx = sd.Context(0)
    cut = nt // world
    run(ctx, buf, off, buf, off, 1, "single GPU (whole index, all queries)")
    run(ctx, buf, off, buf[:int(off[cut])], off[:cut + 1], 1, "built layout, one rank of %d (whole index, 1/%d of the queries)" % (world, world))
    run(ctx, buf[:int(off[cut])], off[:cut + 1], buf, off, 0, "target-sharded, one rank of %d (1/%d of the index, all queries)" % (world, world))


if __name__ == "__main__":
    main()
