#!/usr/bin/env python3
"""Regenerate the committed golden fixtures by running the UNMODIFIED reference binary
(oracle/_ref/spacedust, built by oracle/build_ref.py) on seeded synthetic proteomes.
Run in the build container:  python tests/golden/make_golden.py"""
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import helpers  # noqa: E402
from spacedust_b200 import synth  # noqa: E402


def prefilter_fixture(n_proteomes=2, genes=60, seed=3):
    d = tempfile.mkdtemp(prefix="sdgold_")
    prot = synth.generate(n_proteomes, genes, seed=seed)
    paths = synth.write_fasta(prot, os.path.join(d, "faa"))
    db = os.path.join(d, "setDB")
    helpers.run_ref(["createsetdb"] + paths + [db, os.path.join(d, "tmp"), "-v", "1"])
    helpers.run_ref(["prefilter", db, db, os.path.join(d, "pref"), "-s", "5.7", "--max-seqs", "300", "-c", "0.8",
                     "--cov-mode", "2", "--min-ungapped-score", "15", "--threads", "1", "-v", "1"])
    keys, buf, off = helpers.load_seqdb(db)
    key2idx = {int(k): i for i, k in enumerate(keys)}
    ref = helpers.read_db(os.path.join(d, "pref"))
    q, t, s, dg = [], [], [], []
    for i, k in enumerate(keys):
        for line in ref[int(k)].split(b"\n"):
            if not line:
                continue
            a, b, c = line.split(b"\t")
            q.append(i); t.append(key2idx[int(a)]); s.append(int(b)); dg.append(int(c))
    np.savez_compressed(os.path.join(HERE, "pref_syn%dx%d.npz" % (n_proteomes, genes)), n_proteomes=n_proteomes,
                        genes=genes, seed=seed, query=np.array(q, dtype=np.uint32), target=np.array(t, dtype=np.uint32),
                        score=np.array(s, dtype=np.int32), diag=np.array(dg, dtype=np.int16))
    shutil.rmtree(d, ignore_errors=True)
    print("pref fixture: %d hits" % len(q))


if __name__ == "__main__":
    if not helpers.have_ref():
        sys.exit("oracle/_ref/spacedust missing: run python oracle/build_ref.py")
    prefilter_fixture()
