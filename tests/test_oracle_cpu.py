"""CPU tests: the oracle restatement against the unmodified reference binary (oracle/_ref) and
against the committed golden fixtures."""
import os
import shutil
import tempfile

import numpy as np
import pytest

import helpers
from spacedust_b200 import synth


@pytest.fixture(scope="module")
def ref_setdb(synth_small):
    if not helpers.have_ref():
        pytest.skip("oracle/_ref/spacedust not built")
    d = tempfile.mkdtemp(prefix="sdref_")
    paths = synth.write_fasta(synth_small, os.path.join(d, "faa"))
    helpers.run_ref(["createsetdb"] + paths + [os.path.join(d, "setDB"), os.path.join(d, "tmp"), "-v", "1"])
    yield d
    shutil.rmtree(d, ignore_errors=True)


def test_oracle_prefilter_vs_reference(ref_setdb, synth_small):
    d = ref_setdb
    helpers.run_ref(["prefilter", os.path.join(d, "setDB"), os.path.join(d, "setDB"), os.path.join(d, "pref"),
                     "-s", "5.7", "--max-seqs", "300", "-c", "0.8", "--cov-mode", "2", "--min-ungapped-score", "15",
                     "--threads", "2", "-v", "1"])
    keys, buf, off = helpers.load_seqdb(os.path.join(d, "setDB"))
    pbuf, poff, _ = synth.pack(synth_small)
    assert np.array_equal(buf, pbuf) and np.array_equal(off, poff)       # createsetdb keeps file order
    hits, st = helpers.orc_prefilter(buf, off, buf, off, helpers.default_pref_params())
    mine = helpers.hits_to_lines(hits, keys, keys)
    ref = helpers.read_db(os.path.join(d, "pref"))
    for k in keys:
        r = [l for l in ref[int(k)].split(b"\n") if l]
        assert r == mine[int(k)], "query %d" % k


def test_oracle_prefilter_golden():
    """Committed fixture: reference `prefilter` output for a seeded synthetic set (tests/golden/make_golden.py)."""
    g = os.path.join(helpers.GOLDEN, "pref_syn2x60.npz")
    if not os.path.exists(g):
        pytest.skip("golden fixture missing")
    z = np.load(g)
    prot = synth.generate(int(z["n_proteomes"]), int(z["genes"]), seed=int(z["seed"]))
    buf, off, _ = synth.pack(prot)
    hits, _ = helpers.orc_prefilter(buf, off, buf, off, helpers.default_pref_params())
    assert np.array_equal(hits["query"], z["query"])
    assert np.array_equal(hits["target"], z["target"])
    assert np.array_equal(hits["score"], z["score"])
    assert np.array_equal(hits["diag"], z["diag"])
