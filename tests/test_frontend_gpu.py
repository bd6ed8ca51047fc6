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
