"""GPU parity at BASELINE scale (config[1]: 20 x 3000 synthetic proteomes = 60 000 genes, the bench workload): one pass over the
whole set, then seeded samples against the oracle and -- for the clusters of two query proteomes -- against the unmodified
reference binary (oracle/_ref travels to the GPU box).  Also the reference's own example (config[0]) through the front-end."""
import os
import shutil
import subprocess
import tempfile

import numpy as np
import pytest

import helpers
import spacedust_b200 as sd
from spacedust_b200 import synth

pytestmark = pytest.mark.gpu

N_PROTEOMES, GENES, SEED = 20, 3000, 20261017          # = bench.py


@pytest.fixture(scope="module")
def c2():
    prot = synth.generate(N_PROTEOMES, GENES, seed=SEED)
    buf, off, _ = synth.pack(prot)
    return prot, buf, off


@pytest.fixture(scope="module")
def c2_index(gpu_ctx, c2):
    _, buf, off = c2
    idx = sd.TargetIndex(gpu_ctx, buf, off, sd.default_prefilter_params(same_db=1, bin_size=2))
    yield idx
    idx.close()


def _sub(buf, off, ids):
    return helpers.pack_seqs([bytes(buf[int(off[i]):int(off[i + 1])]).decode() for i in ids])


@pytest.mark.parametrize("variant", ["stream", "binned"])
def test_c2_prefilter_sample_vs_oracle(gpu_ctx, c2, c2_index, variant, monkeypatch):
    """All 60 000 queries through the prefilter kernel; 500 seeded queries hit for hit against the oracle (which sees the same
    60 000 targets); the P/M/R counters of the sample are checked through a second GPU run over just the sample."""
    _, buf, off = c2
    if variant == "binned":
        monkeypatch.setenv("SD_PF_STREAM_MAX", "0")
    gp = sd.default_prefilter_params(same_db=1, bin_size=2)
    hits, st = sd.prefilter(gpu_ctx, c2_index, buf, off, gp)
    nq = len(off) - 1
    rng = np.random.default_rng(2)
    lens = np.diff(off).astype(np.int64)
    sample = np.unique(np.concatenate([rng.choice(nq, size=480, replace=False), np.argsort(lens)[-20:]]))    # + the longest queries
    qb, qo = _sub(buf, off, sample)
    ohits, ost = helpers.orc_prefilter(buf, off, qb, qo, helpers.default_pref_params(sameDB=0, binSize=2))
    # the sample again on the GPU without the identity rule: same mode as the oracle run, counters comparable
    g2 = sd.default_prefilter_params(same_db=0, bin_size=2)
    shits, sst = sd.prefilter(gpu_ctx, c2_index, qb, qo, g2)
    for k_o, k_g in (("probes", "probes"), ("matches", "matches"), ("candidates", "candidates"), ("diagResidues", "diag_residues"),
                     ("hits", "hits"), ("truncated", "truncated")):
        assert ost[k_o] == sst[k_g], (k_o, ost[k_o], sst[k_g])
    assert len(shits) == len(ohits)
    for f in ("query", "target", "score", "diag"):
        assert np.array_equal(shits[f], ohits[f]), f
    # and the full same-DB run: per sampled query the non-identity hits agree wherever --max-seqs did not cut (the identity hit
    # takes one of the 300 places)
    starts = np.searchsorted(hits["query"], sample, side="left")
    ends = np.searchsorted(hits["query"], sample, side="right")
    ostarts = np.searchsorted(ohits["query"], np.arange(len(sample)), side="left")
    oends = np.searchsorted(ohits["query"], np.arange(len(sample)), side="right")
    checked = 0
    for j, q in enumerate(sample):
        g = hits[starts[j]:ends[j]]
        assert len(g) > 0 and g["target"][0] == q and g["score"][0] == 65535
        g = g[1:]
        o = ohits[ostarts[j]:oends[j]]
        o = o[o["target"] != q]
        if len(o) + 1 < 300:
            assert len(o) == len(g), q
            for f in ("target", "score", "diag"):
                assert np.array_equal(o[f], g[f]), (q, f)
            checked += 1
    assert checked > 400


def test_c2_alignments_sample_vs_oracle(gpu_ctx, c2, c2_index):
    """Prefilter pairs of 60 seeded queries (against all 60 000 targets) through the Smith-Waterman / traceback kernels: result lines
    (score, E-value, coordinates, CIGAR) equal the oracle's, with the E-value search space of the whole target DB."""
    _, buf, off = c2
    nq = len(off) - 1
    rng = np.random.default_rng(3)
    lens = np.diff(off).astype(np.int64)
    sample = np.unique(np.concatenate([rng.choice(nq, size=56, replace=False), np.argsort(lens)[-4:]]))
    qb, qo = _sub(buf, off, sample)
    gp = sd.default_prefilter_params(same_db=0, bin_size=2)
    hits, _ = sd.prefilter(gpu_ctx, c2_index, qb, qo, gp)
    ap = sd.default_align_params(same_db=0)
    alns, pool, ast = sd.align(gpu_ctx, c2_index, qb, qo, hits["query"], hits["target"], ap)
    keys = np.arange(nq, dtype=np.uint32)
    qkeys = np.arange(len(sample), dtype=np.uint32)
    oal, obt, ost = helpers.orc_align(buf, off, keys, qb, qo, qkeys, hits["query"], hits["target"], helpers.default_align_params(sameDB=0))
    mine = {}
    for a in alns:
        mine.setdefault(int(a["query"]), []).append(sd.format_alignment(a, a["target"], pool).rstrip(b"\n"))
    ref = helpers.format_alignments(oal, obt, qkeys)
    assert ast["alignments"] == ost["alignments"] > 5000 and ost["passed"] == len(alns) > 200
    for k in qkeys:
        assert ref[int(k)] == mine.get(int(k), []), "query %d" % sample[int(k)]


def test_c2_clusters_two_sets_vs_reference(gpu_ctx, c2, c2_index):
    """Two query proteomes against the 20-proteome target setDB: final clusters (membership, per-hit P-values, cluster P-values as
    printed) equal the unmodified reference's clustersearch TSV."""
    if not helpers.have_ref():
        pytest.skip("oracle/_ref/spacedust not built")
    prot, buf, off = c2
    meta = synth.set_metadata(prot)
    d = tempfile.mkdtemp(prefix="sdscale_")
    try:
        paths = synth.write_fasta(prot, os.path.join(d, "faa"))
        helpers.run_ref(["createsetdb"] + paths + [os.path.join(d, "tDB"), os.path.join(d, "tmp_t"), "-v", "1"])
        helpers.run_ref(["createsetdb"] + paths[:2] + [os.path.join(d, "qDB"), os.path.join(d, "tmp_q"), "-v", "1"])
        out = os.path.join(d, "res.tsv")
        helpers.run_ref(["clustersearch", os.path.join(d, "qDB"), os.path.join(d, "tDB"), out, os.path.join(d, "tmp_cs"),
                         "--filter-self-match", "1", "--threads", str(os.cpu_count() or 1), "-v", "1"])
        src = helpers.read_source(os.path.join(d, "tDB"))
        name_to_key = {v: k for k, v in helpers.read_lookup(os.path.join(d, "tDB")).items()}
        ref = helpers.parse_result_tsv(out, src, name_to_key)
        nq = int(meta["set_size"][0] + meta["set_size"][1])
        q_meta = {"set_of": meta["set_of"][:nq], "pos": meta["pos"][:nq], "strand": meta["strand"][:nq], "set_size": meta["set_size"]}
        batch = sd.QueryBatch(gpu_ctx, buf[:int(off[nq])], off[:nq + 1])
        try:
            res = sd.clustersearch(gpu_ctx, c2_index, batch, q_meta, meta, sd.default_prefilter_params(same_db=1),
                                   sd.default_align_params(same_db=1), sd.default_cluster_params(filter_self_match=1))
        finally:
            batch.close()
        mine = helpers.canonical_clusters(helpers.gpu_clusters(res))
        assert len(ref) > 500
        assert mine == ref
    finally:
        shutil.rmtree(d, ignore_errors=True)


def test_c1_reference_example_through_frontend(gpu_ctx):
    """BASELINE config[0]: the reference's own example (two E. coli / H. pylori proteomes, fixture copies under tests/golden/) through
    `sdb200 createsetdb` + `sdb200 clustersearch` on the GPU: 308 hit lines, 108 clusters, 2 with P < 1e-20 (util/run_regression.sh:23-26)."""
    faa = [os.path.join(helpers.GOLDEN, "examples", n) for n in ("NC_000913.faa.gz", "NC_000915.faa.gz")]
    if not all(os.path.exists(f) for f in faa):
        pytest.skip("example fixtures missing")
    import gzip
    d = tempfile.mkdtemp(prefix="sdc1_")
    try:
        paths = []
        for f in faa:
            p = os.path.join(d, os.path.basename(f)[:-3])
            with gzip.open(f, "rb") as src, open(p, "wb") as dst:
                shutil.copyfileobj(src, dst)
            paths.append(p)
        sdb = os.path.join(os.path.dirname(sd.LIB_PATH), "sdb200")
        subprocess.run([sdb, "createsetdb"] + paths + [os.path.join(d, "setDB"), os.path.join(d, "tmp")], check=True)
        out = os.path.join(d, "res.tsv")
        subprocess.run([sdb, "clustersearch", os.path.join(d, "setDB"), os.path.join(d, "setDB"), out, os.path.join(d, "tmp2"),
                        "--filter-self-match", "1"], check=True)
        lines = open(out).read().split("\n")
        hits = [l for l in lines if l.startswith(">")]
        clusters = [l for l in lines if l.startswith("#")]
        assert len(hits) == 308 and len(clusters) == 108
        assert sum(1 for l in clusters if float(l.split("\t")[3]) < 1e-20) == 2
        golden = os.path.join(helpers.GOLDEN, "examples", "result_canonical.tsv")
        if os.path.exists(golden):
            canon = sorted("\t".join(l.split("\t")[1:]) if l.startswith("#") else l for l in lines if l)
            assert canon == sorted(open(golden).read().split("\n")[:-1])
    finally:
        shutil.rmtree(d, ignore_errors=True)
