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


def test_oracle_align_vs_reference(ref_setdb, synth_small):
    """oracle align on the reference's own prefilter DB == reference result DB (text identical)."""
    d = ref_setdb
    db = os.path.join(d, "setDB")
    if not os.path.exists(os.path.join(d, "pref.index")):
        helpers.run_ref(["prefilter", db, db, os.path.join(d, "pref"), "-s", "5.7", "--max-seqs", "300", "-c", "0.8",
                         "--cov-mode", "2", "--min-ungapped-score", "15", "--threads", "2", "-v", "1"])
    helpers.run_ref(["align", db, db, os.path.join(d, "pref"), os.path.join(d, "result"), "-a", "1", "--alignment-mode", "2",
                     "-e", "10", "--min-aln-len", "30", "-c", "0.8", "--cov-mode", "2", "--threads", "2", "-v", "1"])
    keys, buf, off = helpers.load_seqdb(db)
    pq, pt = helpers.parse_pref_db(os.path.join(d, "pref"), keys, keys)
    alns, bt, st = helpers.orc_align(buf, off, keys, buf, off, keys, pq, pt, helpers.default_align_params())
    mine = helpers.format_alignments(alns, bt, keys)
    ref = helpers.read_db(os.path.join(d, "result"))
    for k in keys:
        assert [l for l in ref[int(k)].split(b"\n") if l] == mine[int(k)], "query %d" % k
    assert st["passed"] > 0 and st["wordMode"] > 0


def test_oracle_clustersearch_vs_reference(ref_setdb, synth_small):
    """Whole path: reference `clustersearch --filter-self-match 1` final TSV == oracle pipeline (canonical order)."""
    d = ref_setdb
    db = os.path.join(d, "setDB")
    out = os.path.join(d, "res.tsv")
    helpers.run_ref(["clustersearch", db, db, out, os.path.join(d, "tmpcs"), "--filter-self-match", "1", "--threads", "2", "-v", "1"])
    o = helpers.oracle_clustersearch(synth_small, filter_self_match=True)
    # the reference's lookup must be what synth.set_metadata predicts
    lookup = {}
    for line in open(db + ".lookup"):
        a = line.rstrip("\n").split("\t")
        lookup[int(a[0])] = a[1]
    assert lookup == o["lookup"]
    src = {}
    for line in open(db + ".source"):
        a = line.rstrip("\n").split("\t")
        src[a[1]] = int(a[0])
    name_to_key = {v: k for k, v in lookup.items()}
    ref = helpers.parse_result_tsv(out, src, name_to_key)
    assert ref == helpers.canonical_clusters(o["clusters"])
    assert len(ref) > 5


def test_oracle_clustersearch_c5_flags_vs_reference(ref_setdb, synth_small):
    """BASELINE config 5's clustering flags (--max-gene-gap 10 --cluster-size 5): reference TSV == oracle pipeline."""
    d = ref_setdb
    db = os.path.join(d, "setDB")
    out = os.path.join(d, "res_c5.tsv")
    helpers.run_ref(["clustersearch", db, db, out, os.path.join(d, "tmpcs_c5"), "--filter-self-match", "1", "--max-gene-gap", "10",
                     "--cluster-size", "5", "--threads", "2", "-v", "1"])
    o = helpers.oracle_clustersearch(synth_small, filter_self_match=True, max_gene_gap=10, cluster_size=5)
    src = {}
    for line in open(db + ".source"):
        a = line.rstrip("\n").split("\t")
        src[a[1]] = int(a[0])
    name_to_key = {v: k for k, v in o["lookup"].items()}
    ref = helpers.parse_result_tsv(out, src, name_to_key)
    assert ref == helpers.canonical_clusters(o["clusters"])
    assert len(ref) > 2 and all(len(c[-1]) >= 5 for c in ref)


def _saturated_family(n_copies=60, n_unrelated=20, length=300, seed=9):
    """One ancestral protein in many close copies (88-97 % identity: every pair's diagonal score is far above the 8-bit
    cap) plus unrelated proteins, as ONE proteome in Prodigal form."""
    rng = np.random.default_rng(seed)
    anc = rng.choice(20, size=length, p=synth.BACKGROUND).astype(np.uint8)
    seqs = [synth.LETTERS[synth._mutate(rng, anc, rng.uniform(0.88, 0.97))].tobytes() + b"*" for _ in range(n_copies)]
    seqs += [synth.LETTERS[rng.choice(20, size=int(rng.integers(80, 400)), p=synth.BACKGROUND)].tobytes() + b"*" for _ in range(n_unrelated)]
    names, pos = [], 1
    for i, sq in enumerate(seqs):
        end = pos + 3 * len(sq) - 1
        names.append("SAT00001.1_%d # %d # %d # 1 # ID=1_%d;partial=00" % (i + 1, pos, end, i + 1))
        pos = end + 50
    return [{"acc": "SAT00001.1", "names": names, "seqs": seqs}]


def test_oracle_prefilter_truncated_scores_vs_reference():
    """More than --max-seqs targets at the 8-bit cap: the reference rescales the uncapped scores (rescoreHits,
    QueryMatcher.cpp:157-171,525-544); the oracle must follow, hit for hit."""
    if not helpers.have_ref():
        pytest.skip("oracle/_ref/spacedust not built")
    prot = _saturated_family()
    d = tempfile.mkdtemp(prefix="sdref_")
    try:
        paths = synth.write_fasta(prot, os.path.join(d, "faa"))
        db = os.path.join(d, "setDB")
        helpers.run_ref(["createsetdb"] + paths + [db, os.path.join(d, "tmp"), "-v", "1"])
        helpers.run_ref(["prefilter", db, db, os.path.join(d, "pref"), "-s", "5.7", "--max-seqs", "16", "-c", "0.8", "--cov-mode", "2",
                         "--min-ungapped-score", "15", "--threads", "2", "-v", "1"])
        keys, buf, off = helpers.load_seqdb(db)
        hits, st = helpers.orc_prefilter(buf, off, buf, off, helpers.default_pref_params(maxSeqs=16))
        mine = helpers.hits_to_lines(hits, keys, keys)
        ref = helpers.read_db(os.path.join(d, "pref"))
        n_rescaled = 0
        for k in keys:
            r = [l for l in ref[int(k)].split(b"\n") if l]
            assert r == mine[int(k)], "query %d" % k
            n_rescaled += sum(1 for l in r[1:] if int(l.split(b"\t")[1]) > 255)
        assert n_rescaled > 200          # the family members' lists really went through the rescaling
    finally:
        shutil.rmtree(d, ignore_errors=True)


PREF_VARIANTS = [
    ("s4", ["-s", "4.0", "--max-seqs", "300", "-c", "0.8", "--cov-mode", "2", "--min-ungapped-score", "15"], dict(sensitivity=4.0)),
    ("s7.5", ["-s", "7.5", "--max-seqs", "300", "-c", "0.8", "--cov-mode", "2", "--min-ungapped-score", "15"], dict(sensitivity=7.5)),
    ("cov0", ["-s", "5.7", "--max-seqs", "300", "-c", "0.5", "--cov-mode", "0", "--min-ungapped-score", "15"], dict(covThr=0.5, covMode=0)),
    ("cov1", ["-s", "5.7", "--max-seqs", "300", "-c", "0.8", "--cov-mode", "1", "--min-ungapped-score", "15"], dict(covMode=1)),
    ("nomask", ["-s", "5.7", "--mask", "0", "--max-seqs", "300", "-c", "0.8", "--cov-mode", "2", "--min-ungapped-score", "15"], dict(mask=0)),
    ("nocomp", ["-s", "5.7", "--comp-bias-corr", "0", "--max-seqs", "300", "-c", "0.8", "--cov-mode", "2", "--min-ungapped-score", "15"], dict(compBias=0)),
    ("tight", ["-s", "5.7", "--max-seqs", "20", "-c", "0.8", "--cov-mode", "2", "--min-ungapped-score", "30"], dict(maxSeqs=20, minDiagScore=30)),
]


@pytest.mark.parametrize("tag,flags,kw", PREF_VARIANTS, ids=[v[0] for v in PREF_VARIANTS])
def test_oracle_prefilter_parameter_sweep(ref_setdb, tag, flags, kw):
    """The oracle follows the reference's `prefilter` flags (sensitivity, coverage modes, masking, composition bias, tight cuts)."""
    d = ref_setdb
    db = os.path.join(d, "setDB")
    out = os.path.join(d, "pref_" + tag)
    helpers.run_ref(["prefilter", db, db, out] + flags + ["--threads", "2", "-v", "1"])
    keys, buf, off = helpers.load_seqdb(db)
    hits, _ = helpers.orc_prefilter(buf, off, buf, off, helpers.default_pref_params(**kw))
    mine = helpers.hits_to_lines(hits, keys, keys)
    ref = helpers.read_db(out)
    for k in keys:
        assert [l for l in ref[int(k)].split(b"\n") if l] == mine[int(k)], "query %d" % k


ALIGN_VARIANTS = [
    ("e1e-5", ["-e", "1e-5", "--min-aln-len", "30", "-c", "0.8", "--cov-mode", "2"], dict(evalThr=1e-5)),
    ("cov0", ["-e", "10", "--min-aln-len", "0", "-c", "0.5", "--cov-mode", "0"], dict(covThr=0.5, covMode=0, minAlnLen=0)),
    ("cov1", ["-e", "10", "--min-aln-len", "30", "-c", "0.7", "--cov-mode", "1"], dict(covThr=0.7, covMode=1)),
    ("open", ["-e", "100", "--min-aln-len", "0", "-c", "0.0"], dict(covThr=0.0, evalThr=100.0, minAlnLen=0)),
    ("minid", ["-e", "10", "--min-aln-len", "30", "-c", "0.8", "--cov-mode", "2", "--min-seq-id", "0.5"], dict(seqIdThr=0.5)),
    ("nocomp", ["-e", "10", "--min-aln-len", "30", "-c", "0.8", "--cov-mode", "2", "--comp-bias-corr", "0"], dict(compBias=0)),
]


@pytest.mark.parametrize("tag,flags,kw", ALIGN_VARIANTS, ids=[v[0] for v in ALIGN_VARIANTS])
def test_oracle_align_parameter_sweep(ref_setdb, tag, flags, kw):
    """... and the `align` flags (E-value, coverage modes, identity, composition bias).  Gap costs other than 11/1 are NOT covered:
    the reference estimates their Gumbel parameters by simulation (EvalueComputation.h:56-110); the product refuses them."""
    d = ref_setdb
    db = os.path.join(d, "setDB")
    pref = os.path.join(d, "pref_all")
    if not os.path.exists(pref + ".index"):
        helpers.run_ref(["prefilter", db, db, pref, "-s", "5.7", "--max-seqs", "300", "-c", "0.0", "--min-ungapped-score", "15",
                         "--threads", "2", "-v", "1"])
    out = os.path.join(d, "res_" + tag)
    helpers.run_ref(["align", db, db, pref, out, "-a", "1", "--alignment-mode", "2"] + flags + ["--threads", "2", "-v", "1"])
    keys, buf, off = helpers.load_seqdb(db)
    pq, pt = helpers.parse_pref_db(pref, keys, keys)
    alns, bt, _ = helpers.orc_align(buf, off, keys, buf, off, keys, pq, pt, helpers.default_align_params(**kw))
    mine = helpers.format_alignments(alns, bt, keys)
    ref = helpers.read_db(out)
    for k in keys:
        assert [l for l in ref[int(k)].split(b"\n") if l] == mine[int(k)], "query %d" % k


@pytest.fixture(scope="module")
def ref_profile_db(ref_setdb):
    """Profile DB made by the reference itself (prefilter -> align -e 0.001 -> result2profile), as iteration 1 of
    `search --num-iterations` does (lib/mmseqs/data/workflow/blastpgp.sh:73-133)."""
    d = ref_setdb
    db = os.path.join(d, "setDB")
    if not os.path.exists(os.path.join(d, "prof0.index")):
        helpers.run_ref(["prefilter", db, db, os.path.join(d, "it0_pref"), "-s", "5.7", "--max-seqs", "300", "--threads", "2", "-v", "1"])
        helpers.run_ref(["align", db, db, os.path.join(d, "it0_pref"), os.path.join(d, "it0_aln"), "-a", "1", "-e", "0.001", "--threads", "2", "-v", "1"])
        helpers.run_ref(["result2profile", db, db, os.path.join(d, "it0_aln"), os.path.join(d, "prof0"), "--threads", "2", "-v", "1"])
    return os.path.join(d, "prof0")


def test_oracle_profile_prefilter_vs_reference(ref_setdb, ref_profile_db):
    """Row a22 groundwork: `prefilter <profileDB> <targetDB>` -- per-column sorted score lists combined into similar k-mers,
    BLOSUM62-based masking, index threshold 0, profile/4 diagonal scores -- restated by the oracle, hit for hit."""
    d = ref_setdb
    db = os.path.join(d, "setDB")
    assert open(ref_profile_db + ".dbtype", "rb").read() == b"\x02\x00\x00\x00"      # profile, no context pseudo counts
    helpers.run_ref(["prefilter", ref_profile_db, db, os.path.join(d, "prof_pref"), "-s", "5.7", "--max-seqs", "300", "--threads", "2", "-v", "1"])
    keys, buf, off = helpers.load_seqdb(db)
    pkeys, pbuf, poff = helpers.load_profiledb(ref_profile_db)
    hits, st = helpers.orc_prefilter_profile(buf, off, pbuf, poff, helpers.default_pref_params(sameDB=0, covThr=0.0))
    mine = helpers.hits_to_lines(hits, pkeys, keys)
    ref = helpers.read_db(os.path.join(d, "prof_pref"))
    for k in pkeys:
        assert [l for l in ref[int(k)].split(b"\n") if l] == mine[int(k)], "profile %d" % k
    assert len(hits) > len(pkeys) and st["probes"] > 10 ** 6


def test_oracle_profile_align_vs_reference(ref_setdb, ref_profile_db):
    """Row a22 groundwork: `align <profileDB> <targetDB>` (profile/4 score rows, X = 0, no composition bias, striped-SW semantics,
    traceback, identities against the profile's query letter) restated by the oracle: result lines identical."""
    d = ref_setdb
    db = os.path.join(d, "setDB")
    pref = os.path.join(d, "prof_pref")
    if not os.path.exists(pref + ".index"):
        helpers.run_ref(["prefilter", ref_profile_db, db, pref, "-s", "5.7", "--max-seqs", "300", "--threads", "2", "-v", "1"])
    helpers.run_ref(["align", ref_profile_db, db, pref, os.path.join(d, "prof_aln"), "-a", "1", "-e", "0.001", "--threads", "2", "-v", "1"])
    keys, buf, off = helpers.load_seqdb(db)
    pkeys, pbuf, poff = helpers.load_profiledb(ref_profile_db)
    pq, pt = helpers.parse_pref_db(pref, pkeys, keys)
    alns, bt, st = helpers.orc_align(buf, off, keys, pbuf, poff, pkeys, pq, pt,
                                     helpers.default_align_params(evalThr=0.001, covThr=0.0, minAlnLen=0, sameDB=0), profile=True)
    mine = helpers.format_alignments(alns, bt, pkeys)
    ref = helpers.read_db(os.path.join(d, "prof_aln"))
    for k in pkeys:
        assert [l for l in ref[int(k)].split(b"\n") if l] == mine[int(k)], "profile %d" % k
    assert st["passed"] > len(pkeys) and st["wordMode"] > 0


def _overflow_family(seed=11, n_a=6000, len_a=560, n_b=6000, len_b=1100, n_unrel=200):
    """Two huge families (92-99 % identity) + unrelated proteins, shuffled: a query from family A returns > 2*10^6 index entries
    (one overflow of the reference's hit buffer, QueryMatcher.cpp:43,280), one from family B > 4*10^6 (two overflows)."""
    rng = np.random.default_rng(seed)
    anc_a = rng.choice(20, size=len_a, p=synth.BACKGROUND).astype(np.uint8)
    anc_b = rng.choice(20, size=len_b, p=synth.BACKGROUND).astype(np.uint8)
    seqs = [synth.LETTERS[synth._mutate(rng, anc_a, rng.uniform(0.92, 0.99))].tobytes() for _ in range(n_a)]
    seqs += [synth.LETTERS[synth._mutate(rng, anc_b, rng.uniform(0.92, 0.99))].tobytes() for _ in range(n_b)]
    seqs += [synth.LETTERS[rng.choice(20, size=int(rng.integers(80, 400)), p=synth.BACKGROUND)].tobytes() for _ in range(n_unrel)]
    perm = rng.permutation(len(seqs))
    inv = np.argsort(perm)
    targets = [seqs[i] for i in perm]
    q_idx = [int(inv[i]) for i in (0, 1, 2, n_a, n_a + 1, n_a + 2, n_a + n_b, n_a + n_b + 1)]
    return targets, [targets[i] for i in q_idx]


def _write_plain_fasta(path, seqs, prefix):
    with open(path, "wb") as f:
        for i, s in enumerate(seqs):
            f.write(b">%s%d\n%s\n" % (prefix.encode(), i, s))


def test_oracle_prefilter_hit_buffer_overflow_vs_reference():
    """Queries whose index hits overflow the reference's hit buffer (maxDbMatches = 2*max(10^6, dbSize), QueryMatcher.cpp:43): the
    buffered part is reduced with a fresh double-hit table and merged later (:280-320, :322-333); from the second overflow on
    the reversed merge + intermediate keepMax (:296-303) applies.  The oracle must follow the reference hit for hit."""
    if not helpers.have_ref():
        pytest.skip("oracle/_ref/spacedust not built")
    targets, queries = _overflow_family()
    d = tempfile.mkdtemp(prefix="sdref_")
    try:
        _write_plain_fasta(os.path.join(d, "t.faa"), targets, "t")
        _write_plain_fasta(os.path.join(d, "q.faa"), queries, "q")
        helpers.run_ref(["createdb", os.path.join(d, "t.faa"), os.path.join(d, "T"), "--shuffle", "0", "-v", "1"])
        helpers.run_ref(["createdb", os.path.join(d, "q.faa"), os.path.join(d, "Q"), "--shuffle", "0", "-v", "1"])
        helpers.run_ref(["prefilter", os.path.join(d, "Q"), os.path.join(d, "T"), os.path.join(d, "pref"), "-s", "5.7", "--max-seqs", "300",
                         "-c", "0.8", "--cov-mode", "2", "--min-ungapped-score", "15", "--threads", "4", "-v", "1"])
        tkeys, tbuf, toff = helpers.load_seqdb(os.path.join(d, "T"))
        qkeys, qbuf, qoff = helpers.load_seqdb(os.path.join(d, "Q"))
        hits, st = helpers.orc_prefilter(tbuf, toff, qbuf, qoff, helpers.default_pref_params(sameDB=0))
        assert st["overflowQueries"] >= 2 and st["multiOverflowQueries"] >= 1 and st["overflowQueries"] > st["multiOverflowQueries"], st
        mine = helpers.hits_to_lines(hits, qkeys, tkeys)
        ref = helpers.read_db(os.path.join(d, "pref"))
        for k in qkeys:
            r = [l for l in ref[int(k)].split(b"\n") if l]
            assert r == mine[int(k)], "query %d" % k
    finally:
        shutil.rmtree(d, ignore_errors=True)
