#!/usr/bin/env python3
"""Debug helper: prefilter with a lowered hit-buffer limit, GPU vs oracle, per-query diff.  usage: gpu_debug_overflow.py LIMIT [stream|binned]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
import spacedust_b200 as sd  # noqa: E402
from spacedust_b200 import synth  # noqa: E402

limit = int(sys.argv[1])
if len(sys.argv) > 2 and sys.argv[2] == "binned":
    os.environ["SD_PF_STREAM_MAX"] = "0"
os.environ["SD_PF_MAXDBMATCHES"] = str(limit)
prot = synth.generate(3, 120, seed=7)
buf, off, _ = synth.pack(prot)
op = helpers.default_pref_params(maxDbMatches=limit)
ohits, ost = helpers.orc_prefilter(buf, off, buf, off, op)
ctx = sd.Context(0)
gp = sd.default_prefilter_params(same_db=1, bin_size=2)
idx = sd.TargetIndex(ctx, buf, off, gp)
ghits, gst = sd.prefilter(ctx, idx, buf, off, gp)
print("oracle", ost)
print("gpu   ", {k: gst[k] for k in ("probes", "matches", "diag_residues", "hits", "candidates", "truncated")})
bad = 0
for q in range(len(off) - 1):
    o = ohits[ohits["query"] == q]
    g = ghits[ghits["query"] == q]
    same = len(o) == len(g) and all(np.array_equal(o[f], g[f]) for f in ("target", "score", "diag"))
    if not same:
        bad += 1
        if bad <= 3:
            # per-query oracle stats
            qb, qo = helpers.pack_seqs([bytes(buf[int(off[q]):int(off[q + 1])]).decode()])
            _, s1 = helpers.orc_prefilter(buf, off, qb, qo, helpers.default_pref_params(sameDB=0, maxDbMatches=limit))
            print("query", q, "len", int(off[q + 1] - off[q]), "oracle hits", len(o), "gpu hits", len(g), "oracle per-query", s1)
            print("  oracle", [(int(x["target"]), int(x["score"]), int(x["diag"])) for x in o][:12])
            print("  gpu   ", [(int(x["target"]), int(x["score"]), int(x["diag"])) for x in g][:12])
print("queries differing:", bad)
