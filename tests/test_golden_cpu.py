"""CPU tests: the oracle restatement against the committed golden fixtures (tests/golden/, produced by the UNMODIFIED
reference binary through tests/golden/make_golden.py) and against the reference's own known answers
(util/run_regression.sh:23-26) when the reference tree is mounted."""
import gzip
import json
import os

import numpy as np
import pytest

import helpers
from spacedust_b200 import synth


def _load(name):
    p = os.path.join(helpers.GOLDEN, name)
    if not os.path.exists(p):
        pytest.skip("golden fixture %s missing" % name)
    with gzip.open(p, "rt") as f:
        return json.load(f)


def test_oracle_align_golden():
    """oracle prefilter -> oracle align == the reference's `result` DB lines (text identical, per query, in order)."""
    g = _load("align_syn2x60.json.gz")
    prot = synth.generate(g["n_proteomes"], g["genes"], seed=g["seed"])
    buf, off, _ = synth.pack(prot)
    keys = np.arange(len(off) - 1, dtype=np.uint32)
    hits, _ = helpers.orc_prefilter(buf, off, buf, off, helpers.default_pref_params())
    alns, bt, st = helpers.orc_align(buf, off, keys, buf, off, keys, hits["query"], hits["target"], helpers.default_align_params())
    mine = helpers.format_alignments(alns, bt, keys)
    for k in keys:
        assert [l.decode() for l in mine[int(k)]] == g["result"][str(int(k))], "query %d" % k
    assert st["passed"] == sum(len(v) for v in g["result"].values())


def test_oracle_clustersearch_golden(synth_small):
    """Whole oracle pipeline == canonicalised final TSV of the reference's clustersearch --filter-self-match 1."""
    g = _load("clusters_syn3x120.json.gz")
    assert (g["n_proteomes"], g["genes"], g["seed"]) == (3, 120, 7)      # = the synth_small fixture
    o = helpers.oracle_clustersearch(synth_small, filter_self_match=True)
    ref = [(a, b, c, d, e, tuple(l)) for a, b, c, d, e, l in g["clusters"]]
    assert helpers.canonical_clusters(o["clusters"]) == ref
    # P-values within 1e-6 relative is the BASELINE bar; the printed %.3E strings are in fact identical
    assert len(ref) == 35


EXAMPLES = "/root/reference/examples"


@pytest.mark.skipif(not os.path.exists(os.path.join(EXAMPLES, "NC_000913.faa")), reason="reference examples not mounted (GPU box)")
def test_oracle_known_answer_on_reference_examples():
    """util/run_regression.sh:23-26: clustersearch of examples/NC_000913.faa + NC_000915.faa with --filter-self-match 1
    reports 308 hit lines in 108 clusters, 2 of them with a clustering P-value < 1e-20."""
    prot = []
    for acc in ("NC_000913", "NC_000915"):
        names, seqs = helpers.read_fasta(os.path.join(EXAMPLES, acc + ".faa"))
        prot.append({"acc": acc, "names": names, "seqs": [s.encode() for s in seqs]})
    o = helpers.oracle_clustersearch(prot, filter_self_match=True)
    clusters = o["clusters"]
    assert len(clusters) == 108
    assert sum(len(lines) for _, lines in clusters) == 308
    assert sum(1 for hdr, _ in clusters if float(hdr.split("\t")[2]) < 1e-20) == 2
