"""GPU: the sdb200 front-end on the reference's DB files.  A set DB written by the reference's createsetdb is searched by
`sdb200 prefilter`, `sdb200 align` and `sdb200 clustersearch`; the prefilter / result DBs must equal the reference's
entry by entry (text identical) and the final TSV must equal the reference's clustersearch TSV (canonical cluster order;
the reference's own cluster numbering is thread-scheduling dependent)."""
import gzip
import json
import os
import shutil
import subprocess
import tempfile

import numpy as np
import pytest

import helpers
from spacedust_b200 import synth

pytestmark = pytest.mark.gpu
SDB200 = os.path.join(helpers.ROOT, "spacedust_b200", "sdb200")


def _run(args):
    r = subprocess.run([SDB200] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    return r.stdout.decode()


@pytest.fixture(scope="module")
def workdir(synth_small):
    d = tempfile.mkdtemp(prefix="sdfe_")
    paths = synth.write_fasta(synth_small, os.path.join(d, "faa"))
    _run(["createsetdb"] + paths + [os.path.join(d, "setDB"), os.path.join(d, "tmp")])
    yield d
    shutil.rmtree(d, ignore_errors=True)


def _tsv_canonical(d, tsv):
    db = os.path.join(d, "setDB")
    lookup = helpers.read_lookup(db)
    return helpers.parse_result_tsv(tsv, helpers.read_source(db), {v: k for k, v in lookup.items()})


def test_clustersearch_tsv_matches_golden(workdir):
    """Native createsetdb -> sdb200 clustersearch -> TSV == committed reference TSV fixture (no reference binary needed)."""
    d = workdir
    db = os.path.join(d, "setDB")
    out = os.path.join(d, "mine.tsv")
    _run(["clustersearch", db, db, out, os.path.join(d, "tmp_mine"), "--filter-self-match", "1", "--write-dbs", "1"])
    with gzip.open(os.path.join(helpers.GOLDEN, "clusters_syn3x120.json.gz"), "rt") as f:
        g = json.load(f)
    ref = [(a, b, c, e, n, tuple(l)) for a, b, c, e, n, l in g["clusters"]]
    assert _tsv_canonical(d, out) == ref
    # clusters / clusters_h DBs: one entry per cluster, header "qSet tSet pCO pMH n" (ClusterHits.cpp:455-483)
    hdr = helpers.read_db_entries(os.path.join(d, "tmp_mine", "clusters_h"))
    body = helpers.read_db_entries(os.path.join(d, "tmp_mine", "clusters"))
    assert len(hdr) == len(body) == len(ref)
    mine = [(h.decode().rstrip("\n"), [l.decode() for l in b.split(b"\n") if l]) for (_, h), (_, b) in zip(hdr, body)]
    assert helpers.canonical_clusters(mine) == ref


def test_modules_match_reference_dbs(workdir):
    if not helpers.have_ref():
        pytest.skip("oracle/_ref/spacedust not built")
    d = workdir
    db = os.path.join(d, "setDB")
    flags_p = ["-s", "5.7", "--max-seqs", "300", "-c", "0.8", "--cov-mode", "2", "--min-ungapped-score", "15"]
    flags_a = ["-a", "1", "--alignment-mode", "2", "-e", "10", "--min-aln-len", "30", "-c", "0.8", "--cov-mode", "2"]
    helpers.run_ref(["prefilter", db, db, os.path.join(d, "pref_ref")] + flags_p + ["--threads", "2", "-v", "1"])
    helpers.run_ref(["align", db, db, os.path.join(d, "pref_ref"), os.path.join(d, "res_ref")] + flags_a + ["--threads", "2", "-v", "1"])
    _run(["prefilter", db, db, os.path.join(d, "pref_mine")] + flags_p + ["--threads", "2"])
    # our align on the REFERENCE's prefilter DB and on our own
    _run(["align", db, db, os.path.join(d, "pref_ref"), os.path.join(d, "res_mine")] + flags_a)
    _run(["align", db, db, os.path.join(d, "pref_mine"), os.path.join(d, "res_mine2")] + flags_a)
    assert helpers.read_db(os.path.join(d, "pref_mine")) == helpers.read_db(os.path.join(d, "pref_ref"))
    ref = helpers.read_db(os.path.join(d, "res_ref"))
    assert helpers.read_db(os.path.join(d, "res_mine")) == ref
    assert helpers.read_db(os.path.join(d, "res_mine2")) == ref
    assert sum(len(v) for v in ref.values()) > 1000
    # dbtype files
    assert open(os.path.join(d, "pref_mine.dbtype"), "rb").read() == open(os.path.join(d, "pref_ref.dbtype"), "rb").read()
    assert open(os.path.join(d, "res_mine.dbtype"), "rb").read() == open(os.path.join(d, "res_ref.dbtype"), "rb").read()


def test_clustersearch_tsv_matches_reference_run():
    """Different seed and sizes than the fixtures: reference createsetdb + clustersearch vs sdb200 on the same set DB."""
    if not helpers.have_ref():
        pytest.skip("oracle/_ref/spacedust not built")
    d = tempfile.mkdtemp(prefix="sdfe_")
    try:
        paths = synth.write_fasta(synth.generate(4, 200, seed=99), os.path.join(d, "faa"))
        db = os.path.join(d, "setDB")
        helpers.run_ref(["createsetdb"] + paths + [db, os.path.join(d, "tmp"), "-v", "1"])
        helpers.run_ref(["clustersearch", db, db, os.path.join(d, "ref.tsv"), os.path.join(d, "tmp_ref"), "--filter-self-match", "1",
                         "--threads", "4", "-v", "1"])
        _run(["clustersearch", db, db, os.path.join(d, "mine.tsv"), os.path.join(d, "tmp_mine"), "--filter-self-match", "1"])
        ref = _tsv_canonical(d, os.path.join(d, "ref.tsv"))
        assert _tsv_canonical(d, os.path.join(d, "mine.tsv")) == ref
        assert len(ref) > 20
    finally:
        shutil.rmtree(d, ignore_errors=True)


def test_modules_default_flags_match_reference(workdir):
    """`prefilter` / `align` WITHOUT flags: the stand-alone modules start from the reference's module defaults (-s 4, -c 0, -e 0.001,
    Parameters.cpp:2364-2425), not from the clustersearch workflow's (-s 5.7, -c 0.8 --cov-mode 2, -e 10)."""
    if not helpers.have_ref():
        pytest.skip("oracle/_ref/spacedust not built")
    d = workdir
    db = os.path.join(d, "setDB")
    helpers.run_ref(["prefilter", db, db, os.path.join(d, "pref_ref_def"), "--threads", "2", "-v", "1"])
    helpers.run_ref(["align", db, db, os.path.join(d, "pref_ref_def"), os.path.join(d, "res_ref_def"), "-a", "1", "--threads", "2", "-v", "1"])
    _run(["prefilter", db, db, os.path.join(d, "pref_mine_def")])
    _run(["align", db, db, os.path.join(d, "pref_mine_def"), os.path.join(d, "res_mine_def"), "-a", "1"])
    assert helpers.read_db(os.path.join(d, "pref_mine_def")) == helpers.read_db(os.path.join(d, "pref_ref_def"))
    ref = helpers.read_db(os.path.join(d, "res_ref_def"))
    assert helpers.read_db(os.path.join(d, "res_mine_def")) == ref
    assert sum(len(v) for v in ref.values()) > 300


def test_persisted_index(gpu_ctx, synth_small, workdir):
    """sd_target_index_save / _load (and `sdb200 createindex` + clustersearch picking <db>.sdidx up): same hits, same final TSV; a
    file built with another sensitivity is refused."""
    import spacedust_b200 as sd
    buf, off, _ = synth.pack(synth_small)
    pp = sd.default_prefilter_params(same_db=1, bin_size=2)
    idx = sd.TargetIndex(gpu_ctx, buf, off, pp)
    path = os.path.join(workdir, "index.bin")
    try:
        hits, st = sd.prefilter(gpu_ctx, idx, buf, off, pp)
        idx.save(path, pp)
    finally:
        idx.close()
    idx2 = sd.TargetIndex.load(gpu_ctx, path, pp)
    try:
        assert idx2.n == len(off) - 1 and idx2.entries > 0
        hits2, st2 = sd.prefilter(gpu_ctx, idx2, buf, off, pp)
    finally:
        idx2.close()
    assert len(hits) == len(hits2) > 0
    for f in ("query", "target", "score", "diag"):
        assert np.array_equal(hits[f], hits2[f]), f
    for k in ("probes", "matches", "diag_residues", "hits"):
        assert st[k] == st2[k], k
    with pytest.raises(sd.SdError):
        sd.TargetIndex.load(gpu_ctx, path, sd.default_prefilter_params(same_db=1, sensitivity=7.5))
    # front-end: createindex, then clustersearch with and without the persisted index
    d = workdir
    db = os.path.join(d, "setDB")
    _run(["createindex", db, os.path.join(d, "tmp_ix")])
    assert os.path.exists(db + ".sdidx")
    out_a, out_b = os.path.join(d, "ix_a.tsv"), os.path.join(d, "ix_b.tsv")
    _run(["clustersearch", db, db, out_a, os.path.join(d, "tmp_ixa"), "--filter-self-match", "1"])
    _run(["clustersearch", db, db, out_b, os.path.join(d, "tmp_ixb"), "--filter-self-match", "1", "--ignore-index", "1"])
    assert open(out_a).read() == open(out_b).read() and open(out_a).read().count("#") > 5
    os.remove(db + ".sdidx")


def test_clusterhits_module_and_intermediate_dbs(tmp_path):
    """Module-level drop-in of the last steps (data/clustersearch.sh:119-146): `sdb200 clusterhits` on the REFERENCE's matches DB
    gives the reference's clusters / clusters_h records; `sdb200 clustersearch --write-dbs 1` leaves result_prefixed, aggregate,
    aggregate_merged and matches(_h) equal to the reference's, and the reference's own clusterhits + summarizeresults resumed from
    OUR matches DB give the same TSV."""
    if not helpers.have_ref():
        pytest.skip("oracle/_ref/spacedust not built")
    from spacedust_b200 import synth
    d = str(tmp_path)
    prot = synth.generate(3, 120, seed=11)
    paths = synth.write_fasta(prot, os.path.join(d, "faa"))
    db = os.path.join(d, "setDB")
    helpers.run_ref(["createsetdb"] + paths + [db, os.path.join(d, "tmp"), "-v", "1"])
    helpers.run_ref(["clustersearch", db, db, os.path.join(d, "ref.tsv"), os.path.join(d, "tmpcs"), "--filter-self-match", "1", "--threads", "2", "-v", "1"])
    L = os.path.join(d, "tmpcs", "latest")

    def pairs(p):
        h, e = helpers.read_db_entries(p + "_h"), helpers.read_db_entries(p)
        assert len(h) == len(e)
        return sorted((x[1], y[1]) for x, y in zip(h, e))
    _run(["clusterhits", db, db, os.path.join(L, "matches"), os.path.join(d, "m_clusters")])
    ref_clusters = pairs(os.path.join(L, "clusters"))
    assert pairs(os.path.join(d, "m_clusters")) == ref_clusters and len(ref_clusters) > 5
    # non-default thresholds reach the module
    _run(["clusterhits", db, db, os.path.join(L, "matches"), os.path.join(d, "m_clusters5"), "--cluster-size", "5", "--max-gene-gap", "10", "--db-output", "1"])
    helpers.run_ref(["clusterhits", db, db, os.path.join(L, "matches"), os.path.join(d, "r_clusters5"), "--cluster-size", "5", "--max-gene-gap", "10",
                     "--db-output", "1", "--threads", "2", "-v", "1"])
    assert pairs(os.path.join(d, "m_clusters5")) == pairs(os.path.join(d, "r_clusters5"))

    out = os.path.join(d, "mine.tsv")
    mt = os.path.join(d, "tmpmine")
    _run(["clustersearch", db, db, out, mt, "--filter-self-match", "1", "--write-dbs", "1", "-v", "1"])
    for name in ("result_prefixed", "aggregate", "aggregate_merged"):
        assert helpers.read_db(os.path.join(mt, name)) == helpers.read_db(os.path.join(L, name)), name
    assert pairs(os.path.join(mt, "matches")) == pairs(os.path.join(L, "matches"))
    # the reference resumes from our matches DB
    helpers.run_ref(["clusterhits", db, db, os.path.join(mt, "matches"), os.path.join(d, "r_from_mine"), "--db-output", "1", "--threads", "2", "-v", "1"])
    assert pairs(os.path.join(d, "r_from_mine")) == ref_clusters


def test_reference_workflow_script_drives_sdb200(tmp_path):
    """Script-level drop-in: the reference's OWN data/clustersearch.sh (the copy its binary leaves in the tmp dir), with the parameter
    strings of a real reference run, executed with MMSEQS pointing at sdb200 -- search (prefilter + align), prefixid, besthitbyset,
    mergeresultsbyset, combinehits, clusterhits, summarizeresults, filterdb, swapdb as separate processes -- gives the reference's TSV."""
    if not helpers.have_ref():
        pytest.skip("oracle/_ref/spacedust not built")
    from spacedust_b200 import synth
    d = str(tmp_path)
    prot = synth.generate(3, 120, seed=11)
    synth.write_fasta(prot, os.path.join(d, "faa"))
    faa = sorted(os.path.join("faa", f) for f in os.listdir(os.path.join(d, "faa")))
    helpers.run_ref(["createsetdb"] + faa + ["setDB", "tmp", "-v", "1"], cwd=d)
    log = helpers.run_ref(["clustersearch", "setDB", "setDB", "ref.tsv", "tmpcs", "--filter-self-match", "1", "--threads", "2", "-v", "3"], cwd=d)

    def par(module, n_pos):
        line = [l for l in log.splitlines() if l.startswith(module + " ")][0]
        return " ".join(line.split(" ")[1 + n_pos:]).strip()      # verbatim, as the workflow puts it into the environment
    env = dict(os.environ)
    env.update({"MMSEQS": SDB200, "SEARCH_PAR": par("search", 4), "BESTHITBYSET_PAR": par("besthitbyset", 4), "COMBINEHITS_PAR": par("combinehits", 5),
                "CLUSTERHITS_PAR": par("clusterhits", 4), "THREADS_PAR": par("prefixid", 2), "VERBOSITY": "-v 3"})
    script = os.path.join(d, "tmpcs", "latest", "clustersearch.sh")
    os.makedirs(os.path.join(d, "tmpmine"))          # the workflow binary creates the tmp directory before it runs the script
    r = subprocess.run(["sh", script, "setDB", "setDB", "mine.tsv", "tmpmine"], cwd=d, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    assert r.returncode == 0, r.stdout.decode()[-3000:]

    def blocks(path):
        out, cur = [], None
        for line in open(path):
            if line.startswith("#"):
                cur = [line.split("\t", 1)[1], []]
                out.append(cur)
            else:
                cur[1].append(line)
        return sorted((a, tuple(b)) for a, b in out)
    ref = blocks(os.path.join(d, "ref.tsv"))
    assert blocks(os.path.join(d, "mine.tsv")) == ref and len(ref) > 5
    for name in ("result", "result_prefixed", "aggregate", "aggregate_merged", "matches", "clusters", "clu_to_seq"):
        assert os.path.exists(os.path.join(d, "tmpmine", name + ".index")), name
