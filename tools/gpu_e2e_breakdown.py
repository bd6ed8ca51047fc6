#!/usr/bin/env python3
"""Wall-clock breakdown of the end-to-end step of bench.py (index build, query upload, search, teardown)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spacedust_b200 as sd  # noqa: E402
from spacedust_b200 import shard, synth  # noqa: E402

n, g = int(sys.argv[1]), int(sys.argv[2])
prot = synth.generate(n, g, seed=20261017)
buf, off, _ = synth.pack(prot)
meta = synth.set_metadata(prot)
ctx = sd.Context(0)
pp = sd.default_prefilter_params(same_db=1)
ap = sd.default_align_params(same_db=1)
cp = sd.default_cluster_params(filter_self_match=1)
ex = shard.ResultExchange(device=torch.device("cuda", 0))
tb = torch.from_numpy(buf.copy()).pin_memory().numpy()
for it in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    idx = sd.TargetIndex(ctx, tb, off, pp)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    b = sd.QueryBatch(ctx, tb, off)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    r = sd.clustersearch(ctx, idx, b, meta, meta, pp, ap, cp, 0, out=ex)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    b.close(); idx.close()
    torch.cuda.synchronize(); t4 = time.perf_counter()
    st = r["prefilter_stats"]["total_ms"] + r["align_stats"]["total_ms"] + r["cluster_stats"]["total_ms"]
    print("iter %d: index %.1f  batch %.1f  search %.1f (stages %.1f)  close %.1f  total %.1f ms" %
          (it, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), st, 1e3 * (t4 - t3), 1e3 * (t4 - t0)), flush=True)
idx = sd.TargetIndex(ctx, tb, off, pp)
b = sd.QueryBatch(ctx, tb, off)
for it in range(3):
    torch.cuda.synchronize(); t2 = time.perf_counter()
    r = sd.clustersearch(ctx, idx, b, meta, meta, pp, ap, cp, 0, out=ex)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    st = r["prefilter_stats"]["total_ms"] + r["align_stats"]["total_ms"] + r["cluster_stats"]["total_ms"]
    print("resident %d: search %.1f (stages %.1f: pf %.1f al %.1f cl %.1f)" % (it, 1e3 * (t3 - t2), st, r["prefilter_stats"]["total_ms"],
          r["align_stats"]["total_ms"], r["cluster_stats"]["total_ms"]), flush=True)
