"""CPU tests of the drop-in boundary: the C-ABI library exports every symbol include/sdgpu.h declares, the product fails
loudly without a CUDA device (no CPU fallback), and the native `sdb200 createsetdb` writes the reference's set DB byte for byte."""
import ctypes
import hashlib
import json
import os
import re
import shutil
import subprocess
import tempfile

import pytest

import helpers
import spacedust_b200 as sd
from spacedust_b200 import synth

SDB200 = os.path.join(helpers.ROOT, "spacedust_b200", "sdb200")


def _declared_symbols():
    text = open(os.path.join(helpers.ROOT, "include", "sdgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sd_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = _declared_symbols()
    assert len(names) >= 24 and "sd_clustersearch_run" in names and "sd_prefilter_run" in names
    L = ctypes.CDLL(sd.LIB_PATH)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert sd.lib().sd_version() >= 100


def _no_gpu():
    import torch
    return not torch.cuda.is_available()


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful on a box without a CUDA device")
def test_no_cpu_fallback():
    """Without a device the context cannot be created (SD_ENODEV) and the front-end exits non-zero with the message."""
    with pytest.raises(sd.SdError) as e:
        sd.Context(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)
    d = tempfile.mkdtemp(prefix="sdfe_")
    try:
        paths = synth.write_fasta(synth.generate(2, 5, seed=3), os.path.join(d, "faa"))
        subprocess.run([SDB200, "createsetdb"] + paths + [os.path.join(d, "setDB"), os.path.join(d, "tmp")], check=True)
        r = subprocess.run([SDB200, "clustersearch", os.path.join(d, "setDB"), os.path.join(d, "setDB"), os.path.join(d, "out.tsv"),
                            os.path.join(d, "tmp")], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        assert r.returncode != 0 and b"Error" in r.stderr
        assert not os.path.exists(os.path.join(d, "out.tsv"))
    finally:
        shutil.rmtree(d, ignore_errors=True)


def test_frontend_rejects_unimplemented_modes():
    d = tempfile.mkdtemp(prefix="sdfe_")
    try:
        paths = synth.write_fasta(synth.generate(2, 5, seed=3), os.path.join(d, "faa"))
        db = os.path.join(d, "setDB")
        subprocess.run([SDB200, "createsetdb"] + paths + [db, os.path.join(d, "tmp")], check=True)
        for flags in (["--search-mode", "1"], ["--num-iterations", "3"], ["--realign", "1"]):
            r = subprocess.run([SDB200, "clustersearch", db, db, os.path.join(d, "o.tsv"), os.path.join(d, "tmp")] + flags,
                               stdout=subprocess.PIPE, stderr=subprocess.PIPE)
            assert r.returncode != 0 and b"not implemented" in r.stderr, flags
    finally:
        shutil.rmtree(d, ignore_errors=True)


def _sha(dirname):
    return {f: hashlib.sha256(open(os.path.join(dirname, f), "rb").read()).hexdigest()
            for f in sorted(os.listdir(dirname)) if f.startswith("setDB")}


def test_createsetdb_golden():
    """Every file of the set DB equals what the reference's createsetdb wrote (committed sha256 fixture)."""
    g = json.load(open(os.path.join(helpers.GOLDEN, "createsetdb_syn2x5.json")))
    d = tempfile.mkdtemp(prefix="sdfe_")
    try:
        paths = synth.write_fasta(synth.generate(g["n_proteomes"], g["genes"], seed=g["seed"]), os.path.join(d, "faa"))
        subprocess.run([SDB200, "createsetdb"] + paths + [os.path.join(d, "setDB"), os.path.join(d, "tmp"), "--threads", "2"], check=True)
        assert _sha(d) == g["sha256"]
    finally:
        shutil.rmtree(d, ignore_errors=True)


@pytest.mark.parametrize("case", ["synth", "examples"])
def test_createsetdb_vs_reference(case, synth_small):
    if not helpers.have_ref():
        pytest.skip("oracle/_ref/spacedust not built")
    d = tempfile.mkdtemp(prefix="sdfe_")
    try:
        if case == "synth":
            paths = synth.write_fasta(synth_small, os.path.join(d, "faa"))
        else:
            paths = [os.path.join("/root/reference/examples", f) for f in ("NC_000913.faa", "NC_000915.faa")]
            if not all(os.path.exists(p) for p in paths):
                pytest.skip("reference examples not mounted")
        os.makedirs(os.path.join(d, "a"))
        os.makedirs(os.path.join(d, "b"))
        helpers.run_ref(["createsetdb"] + paths + [os.path.join(d, "a", "setDB"), os.path.join(d, "a", "tmp"), "-v", "1"])
        subprocess.run([SDB200, "createsetdb"] + paths + [os.path.join(d, "b", "setDB"), os.path.join(d, "b", "tmp")], check=True)
        assert _sha(os.path.join(d, "a")) == _sha(os.path.join(d, "b"))
        assert len(_sha(os.path.join(d, "a"))) == 19
    finally:
        shutil.rmtree(d, ignore_errors=True)


def test_post_modules_vs_reference(synth_small):
    """Host-only modules after the clustering step (data/clustersearch.sh:151-165) on the REFERENCE's own clusters DB:
    summarizeresults TSV identical (cluster order is thread-scheduling dependent in the reference: compared as sorted blocks),
    filterdb --trim-to-one-column identical entries, swapdb the same keys, lengths and key sets per entry."""
    if not helpers.have_ref():
        pytest.skip("oracle/_ref/spacedust not built")
    d = tempfile.mkdtemp(prefix="sdfe_")
    try:
        paths = synth.write_fasta(synth_small, os.path.join(d, "faa"))
        db = os.path.join(d, "setDB")
        helpers.run_ref(["createsetdb"] + paths + [db, os.path.join(d, "tmp"), "-v", "1"])
        helpers.run_ref(["clustersearch", db, db, os.path.join(d, "ref.tsv"), os.path.join(d, "tmpcs"), "--filter-self-match", "1",
                         "--threads", "2", "-v", "1"])
        latest = os.path.join(d, "tmpcs", "latest")
        subprocess.run([SDB200, "summarizeresults", db, db, os.path.join(latest, "clusters"), os.path.join(d, "mine.tsv")], check=True)

        def blocks(path):
            out, cur = [], None
            for line in open(path):
                if line.startswith("#"):
                    cur = [line, []]
                    out.append(cur)
                else:
                    cur[1].append(line)
            return sorted((a, tuple(b)) for a, b in out)
        ref_blocks = blocks(os.path.join(d, "ref.tsv"))
        assert blocks(os.path.join(d, "mine.tsv")) == ref_blocks and len(ref_blocks) > 5

        subprocess.run([SDB200, "filterdb", os.path.join(latest, "clusters"), os.path.join(d, "c2s"), "--trim-to-one-column"], check=True)
        # (the reference writes the entries in thread order; content per key, lengths and dbtype must agree)
        assert helpers.read_db(os.path.join(d, "c2s")) == helpers.read_db(os.path.join(latest, "clu_to_seq"))
        assert sorted((k, l) for k, _, l in helpers.read_index(os.path.join(d, "c2s"))) == \
            sorted((k, l) for k, _, l in helpers.read_index(os.path.join(latest, "clu_to_seq")))
        assert open(os.path.join(d, "c2s.dbtype"), "rb").read() == open(os.path.join(latest, "clu_to_seq.dbtype"), "rb").read()
        subprocess.run([SDB200, "swapdb", os.path.join(d, "c2s"), os.path.join(d, "s2c")], check=True)
        ref_s2c = os.path.join(d, "ref.tsv_seq_to_clu")
        assert sorted((k, l) for k, _, l in helpers.read_index(os.path.join(d, "s2c"))) == sorted((k, l) for k, _, l in helpers.read_index(ref_s2c))
        assert open(os.path.join(d, "s2c.dbtype"), "rb").read() == open(ref_s2c + ".dbtype", "rb").read()
        mine = {k: sorted(v.split()) for k, v in helpers.read_db(os.path.join(d, "s2c")).items()}
        ref = {k: sorted(v.split()) for k, v in helpers.read_db(ref_s2c).items()}
        assert mine == ref
    finally:
        shutil.rmtree(d, ignore_errors=True)


def test_db_reader_accepts_split_data_files():
    """Per-thread split files <db>.0 .. <db>.N-1 with one merged index (DBWriter.cpp:531-623) read like a merged DB."""
    d = tempfile.mkdtemp(prefix="sdfe_")
    try:
        paths = synth.write_fasta(synth.generate(2, 5, seed=3), os.path.join(d, "faa"))
        db = os.path.join(d, "setDB")
        subprocess.run([SDB200, "createsetdb"] + paths + [db, os.path.join(d, "tmp")], check=True)
        # a tiny "clusters" DB in two split files: entries "0\t1\tx\n" style lines
        with open(os.path.join(d, "clu.0"), "wb") as f:
            f.write(b"3\t4\tx\n5\t6\ty\n\0")
        with open(os.path.join(d, "clu.1"), "wb") as f:
            f.write(b"7\t8\tz\n\0")
        with open(os.path.join(d, "clu.index"), "w") as f:
            f.write("0\t0\t13\n1\t13\t7\n")
        with open(os.path.join(d, "clu.dbtype"), "wb") as f:
            f.write((5).to_bytes(4, "little"))
        subprocess.run([SDB200, "filterdb", os.path.join(d, "clu"), os.path.join(d, "one"), "--trim-to-one-column"], check=True)
        got = helpers.read_db(os.path.join(d, "one"))
        assert got == {0: b"3\n5\n", 1: b"7\n"}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def test_iteration_glue_modules_vs_reference(synth_small):
    """subtractdbs / mergedbs (text glue of the iterative search, blastpgp.sh:83-124) on the reference's own DBs."""
    if not helpers.have_ref():
        pytest.skip("oracle/_ref/spacedust not built")
    d = tempfile.mkdtemp(prefix="sdfe_")
    try:
        paths = synth.write_fasta(synth_small, os.path.join(d, "faa"))
        db = os.path.join(d, "setDB")
        helpers.run_ref(["createsetdb"] + paths + [db, os.path.join(d, "tmp"), "-v", "1"])
        helpers.run_ref(["prefilter", db, db, os.path.join(d, "pref"), "-s", "5.7", "--threads", "2", "-v", "1"])
        helpers.run_ref(["align", db, db, os.path.join(d, "pref"), os.path.join(d, "aln"), "-a", "1", "-e", "0.001", "--threads", "2", "-v", "1"])
        helpers.run_ref(["subtractdbs", os.path.join(d, "pref"), os.path.join(d, "aln"), os.path.join(d, "sub_ref"), "--e-profile", "0.001",
                         "--threads", "2", "-v", "1"])
        subprocess.run([SDB200, "subtractdbs", os.path.join(d, "pref"), os.path.join(d, "aln"), os.path.join(d, "sub_mine"), "--e-profile", "0.001"],
                       check=True)
        ref = helpers.read_db(os.path.join(d, "sub_ref"))
        assert helpers.read_db(os.path.join(d, "sub_mine")) == ref
        assert 0 < sum(len(v) for v in ref.values()) < sum(len(v) for v in helpers.read_db(os.path.join(d, "pref")).values())
        # -e and --e-profile differ: the threshold is the smaller one (subtractdbs.cpp:17), in either order
        helpers.run_ref(["align", db, db, os.path.join(d, "pref"), os.path.join(d, "aln10"), "-a", "1", "-e", "10", "--threads", "2", "-v", "1"])
        for tag, e, ep in (("a", "1e-8", "0.1"), ("b", "1", "1e-8")):
            helpers.run_ref(["subtractdbs", os.path.join(d, "aln10"), os.path.join(d, "aln"), os.path.join(d, "sub2_ref" + tag), "--e-profile", ep,
                             "-e", e, "--threads", "2", "-v", "1"])
            subprocess.run([SDB200, "subtractdbs", os.path.join(d, "aln10"), os.path.join(d, "aln"), os.path.join(d, "sub2_mine" + tag),
                            "--e-profile", ep, "-e", e], check=True)
            assert helpers.read_db(os.path.join(d, "sub2_mine" + tag)) == helpers.read_db(os.path.join(d, "sub2_ref" + tag)), tag
        helpers.run_ref(["mergedbs", db, os.path.join(d, "mrg_ref"), os.path.join(d, "aln"), os.path.join(d, "sub_ref"), "-v", "1"])
        subprocess.run([SDB200, "mergedbs", db, os.path.join(d, "mrg_mine"), os.path.join(d, "aln"), os.path.join(d, "sub_mine")], check=True)
        assert helpers.read_db(os.path.join(d, "mrg_mine")) == helpers.read_db(os.path.join(d, "mrg_ref"))
        assert open(os.path.join(d, "mrg_mine.dbtype"), "rb").read() == open(os.path.join(d, "mrg_ref.dbtype"), "rb").read()
    finally:
        shutil.rmtree(d, ignore_errors=True)


def _write_db(path, entries, dbtype_bytes):
    off = 0
    with open(path, "wb") as f, open(path + ".index", "w") as ix:
        for k, data in entries:
            f.write(data + b"\0")
            ix.write("%d\t%d\t%d\n" % (k, off, len(data) + 1))
            off += len(data) + 1
    open(path + ".dbtype", "wb").write(dbtype_bytes)


def test_aggregation_modules_vs_reference(synth_small):
    """The stand-alone text modules between align and clusterhits (data/clustersearch.sh:119-141, csrc/aggmods.cpp) on the
    REFERENCE's own intermediates: prefixid, besthitbyset, mergeresultsbyset give the reference's DBs entry for entry (and the
    same dbtype files); combinehits gives the reference's (header, lines) records -- compared as sorted pairs against the
    2-thread reference run (its entry keys are per-thread counters), key for key against a 1-thread run on weakened hits
    (non-zero combined E-values in the headers)."""
    if not helpers.have_ref():
        pytest.skip("oracle/_ref/spacedust not built")
    import random
    d = tempfile.mkdtemp(prefix="sdfe_")
    try:
        paths = synth.write_fasta(synth_small, os.path.join(d, "faa"))
        db = os.path.join(d, "setDB")
        helpers.run_ref(["createsetdb"] + paths + [db, os.path.join(d, "tmp"), "-v", "1"])
        helpers.run_ref(["clustersearch", db, db, os.path.join(d, "ref.tsv"), os.path.join(d, "tmpcs"), "--filter-self-match", "1",
                         "--threads", "2", "-v", "1"])
        L = os.path.join(d, "tmpcs", "latest")

        def mine(*args):
            subprocess.run([SDB200] + list(args), check=True)
        mine("prefixid", os.path.join(L, "result"), os.path.join(d, "m_prefixed"))
        mine("besthitbyset", db, db, os.path.join(L, "result_prefixed"), os.path.join(d, "m_aggregate"))
        mine("mergeresultsbyset", db + "_set_to_member", os.path.join(L, "aggregate"), os.path.join(d, "m_merged"))
        mine("combinehits", db, db, os.path.join(L, "aggregate_merged"), os.path.join(d, "m_matches"), os.path.join(d, "tmpm"),
             "--filter-self-match", "1", "--alpha", "1", "--threads", "2")
        for m, r in (("m_prefixed", "result_prefixed"), ("m_aggregate", "aggregate"), ("m_merged", "aggregate_merged")):
            a, b = helpers.read_db(os.path.join(d, m)), helpers.read_db(os.path.join(L, r))
            assert a == b and len(a) > 0, m
            assert open(os.path.join(d, m + ".dbtype"), "rb").read() == open(os.path.join(L, r + ".dbtype"), "rb").read(), m

        def pairs(p):
            h, e = helpers.read_db_entries(p + "_h"), helpers.read_db_entries(p)
            assert len(h) == len(e)
            return list(zip(h, e))
        a, b = pairs(os.path.join(d, "m_matches")), pairs(os.path.join(L, "matches"))
        assert sorted((x[0][1], x[1][1]) for x in a) == sorted((x[0][1], x[1][1]) for x in b) and len(a) > 3
        for suffix in ("", "_h"):
            assert open(os.path.join(d, "m_matches" + suffix + ".dbtype"), "rb").read() == open(os.path.join(L, "matches" + suffix + ".dbtype"), "rb").read()

        # few, weak hits per genome pair: the truncated-Fisher sum of the header is no longer 0
        rnd = random.Random(5)
        weak = []
        for k, data in helpers.read_db_entries(os.path.join(L, "aggregate_merged")):
            lines = []
            for ln in data.decode().split("\n"):
                if ln and rnd.random() < 0.07:
                    c = ln.split("\t")
                    c[2] = "%.3E" % (-rnd.uniform(10, 22))
                    lines.append("\t".join(c))
            weak.append((k, ("\n".join(lines) + "\n").encode()))
        _write_db(os.path.join(d, "weak"), weak, open(os.path.join(L, "aggregate_merged.dbtype"), "rb").read())
        helpers.run_ref(["combinehits", db, db, os.path.join(d, "weak"), os.path.join(d, "r_weak"), os.path.join(d, "tmpr"), "--filter-self-match", "1",
                         "--threads", "1", "-v", "1"])
        mine("combinehits", db, db, os.path.join(d, "weak"), os.path.join(d, "m_weak"), os.path.join(d, "tmpw"), "--filter-self-match", "1")
        a, b = pairs(os.path.join(d, "m_weak")), pairs(os.path.join(d, "r_weak"))
        assert a == b and len(a) > 2
        assert any(not h[1].rstrip().endswith(b"0.000E+00") for h, _ in a)
        # flags that select code these modules do not have stop them
        r = subprocess.run([SDB200, "besthitbyset", db, db, os.path.join(L, "result_prefixed"), os.path.join(d, "x"), "--simple-best-hit", "0"],
                           stderr=subprocess.PIPE)
        assert r.returncode != 0 and b"not implemented" in r.stderr
    finally:
        shutil.rmtree(d, ignore_errors=True)


def test_aggregation_modules_edge_cases_vs_reference():
    """Empty entries, an entry whose last line lacks its newline, genes without hits: the four text modules give the reference's
    entries (1-thread reference run: its running keys are then the same as ours)."""
    if not helpers.have_ref():
        pytest.skip("oracle/_ref/spacedust not built")
    d = tempfile.mkdtemp(prefix="sdfe_")
    try:
        prot = synth.generate(2, 8, seed=3)
        paths = synth.write_fasta(prot, os.path.join(d, "faa"))
        db = os.path.join(d, "setDB")
        helpers.run_ref(["createsetdb"] + paths + [db, os.path.join(d, "tmp"), "-v", "1"])
        keys = [k for k, _, _ in helpers.read_index(db)]
        rec = "%d\t%d\t0.9\t%s\t0\t9\t10\t0\t9\t10\t10M"
        ents = []
        for k in keys:
            if k == keys[1]:
                ents.append((k, (rec % (keys[9], 50, "1E-20") + "\n").encode()))
            elif k == keys[2]:
                ents.append((k, (rec % (keys[10], 50, "1E-12") + "\n" + rec % (keys[11], 40, "1E-30")).encode()))
            else:
                ents.append((k, b""))
        _write_db(os.path.join(d, "result"), ents, b"\x05\0\0\0")
        for tool, pre in (("ref", "r_"), ("mine", "m_")):
            def run(args):
                if tool == "ref":
                    helpers.run_ref(args + ["--threads", "1", "-v", "1"], cwd=d)
                else:
                    subprocess.run([SDB200] + args, check=True, cwd=d)
            run(["prefixid", "result", pre + "prefixed"])
            run(["besthitbyset", db, db, pre + "prefixed", pre + "aggregate"])
            run(["mergeresultsbyset", db + "_set_to_member", pre + "aggregate", pre + "merged"])
            run(["combinehits", db, db, pre + "merged", pre + "matches", "tmpc_" + tool, "--filter-self-match", "1"])
        for n in ("prefixed", "aggregate", "merged", "matches", "matches_h"):
            a, b = helpers.read_db_entries(os.path.join(d, "m_" + n)), helpers.read_db_entries(os.path.join(d, "r_" + n))
            assert a == b and len(a) > 0, n
    finally:
        shutil.rmtree(d, ignore_errors=True)


def test_compressed_dbs_are_read():
    """DBs written by the reference with --compressed 1 (zstd frames per entry, DBWriter.cpp:326-395) are expanded on open: a text
    module gives the same output on the compressed and on the plain set DB, and the compressed _set_size / _member_to_set side DBs
    feed besthitbyset like the plain ones."""
    if not helpers.have_ref():
        pytest.skip("oracle/_ref/spacedust not built")
    d = tempfile.mkdtemp(prefix="sdfe_")
    try:
        prot = synth.generate(2, 30, seed=4)
        paths = synth.write_fasta(prot, os.path.join(d, "faa"))
        for c in ("0", "1"):
            helpers.run_ref(["createsetdb"] + paths + [os.path.join(d, "set" + c), os.path.join(d, "tmp" + c), "--compressed", c, "-v", "1"])
        assert open(os.path.join(d, "set1.dbtype"), "rb").read()[3] & 0x80 and not open(os.path.join(d, "set0.dbtype"), "rb").read()[3] & 0x80
        for c in ("0", "1"):
            subprocess.run([SDB200, "prefixid", os.path.join(d, "set" + c), os.path.join(d, "p" + c)], check=True)
        a, b = helpers.read_db(os.path.join(d, "p0")), helpers.read_db(os.path.join(d, "p1"))
        assert a == b and len(a) == 60
        keys = sorted(a)
        rec = "%d\t%d\t0.9\t%s\t0\t9\t10\t0\t9\t10\t10M\n"
        ents = [(k, ((str(k) + "\t" + rec % (keys[40], 50, "1E-20")) if k == keys[1] else "").encode()) for k in keys]
        _write_db(os.path.join(d, "prefixed"), ents, b"\x05\0\0\0")
        for c in ("0", "1"):
            subprocess.run([SDB200, "besthitbyset", os.path.join(d, "set" + c), os.path.join(d, "set" + c), os.path.join(d, "prefixed"),
                            os.path.join(d, "agg" + c)], check=True)
        a, b = helpers.read_db(os.path.join(d, "agg0")), helpers.read_db(os.path.join(d, "agg1"))
        assert a == b and any(v for v in a.values())
    finally:
        shutil.rmtree(d, ignore_errors=True)


def test_compressed_output_is_read_by_the_reference():
    """--compressed 1 on our side: the reference reads the zstd DBs we write -- an aggregation module's compressed output expands to
    the plain one, and the reference's clustersearch on OUR compressed set DB gives the TSV it gives on its own plain set DB."""
    if not helpers.have_ref():
        pytest.skip("oracle/_ref/spacedust not built")
    d = tempfile.mkdtemp(prefix="sdfe_")
    try:
        prot = synth.generate(2, 30, seed=4)
        paths = synth.write_fasta(prot, os.path.join(d, "faa"))
        helpers.run_ref(["createsetdb"] + paths + [os.path.join(d, "rset"), os.path.join(d, "tmpr"), "-v", "1"])
        subprocess.run([SDB200, "createsetdb"] + paths + [os.path.join(d, "mset"), os.path.join(d, "tmpm"), "--compressed", "1"], check=True,
                       stdout=subprocess.DEVNULL)
        assert open(os.path.join(d, "mset.dbtype"), "rb").read()[3] & 0x80
        outs = []
        for db in ("rset", "mset"):
            out = os.path.join(d, db + ".tsv")
            helpers.run_ref(["clustersearch", os.path.join(d, db), os.path.join(d, db), out, os.path.join(d, "tmpcs_" + db), "--filter-self-match", "1",
                             "--threads", "2", "-v", "1"])
            outs.append(sorted(l for l in open(out) if not l.startswith("#")))
        assert outs[0] == outs[1] and len(outs[0]) > 3
        L = os.path.join(d, "tmpcs_rset", "latest")
        subprocess.run([SDB200, "prefixid", os.path.join(L, "result"), os.path.join(d, "c_prefixed"), "--compressed", "1"], check=True)
        assert open(os.path.join(d, "c_prefixed.dbtype"), "rb").read()[3] & 0x80
        assert os.path.getsize(os.path.join(d, "c_prefixed")) < os.path.getsize(os.path.join(L, "result_prefixed"))
        helpers.run_ref(["besthitbyset", os.path.join(d, "rset"), os.path.join(d, "rset"), os.path.join(d, "c_prefixed"), os.path.join(d, "r_agg"),
                         "--threads", "1", "-v", "1"])
        assert helpers.read_db(os.path.join(d, "r_agg")) == helpers.read_db(os.path.join(L, "aggregate"))
    finally:
        shutil.rmtree(d, ignore_errors=True)


def test_createsetdb_gzip_input_vs_reference():
    """gzip-compressed protein FASTA (createdb reads it through zlib): the same 19 files as the reference's createsetdb, byte for byte."""
    if not helpers.have_ref():
        pytest.skip("oracle/_ref/spacedust not built")
    import gzip
    d = tempfile.mkdtemp(prefix="sdfe_")
    try:
        paths = synth.write_fasta(synth.generate(2, 20, seed=6), os.path.join(d, "faa"))
        gz = []
        for p in paths:
            with open(p, "rb") as f, gzip.open(p + ".gz", "wb") as g:
                g.write(f.read())
            gz.append(p + ".gz")
        os.makedirs(os.path.join(d, "a"))
        os.makedirs(os.path.join(d, "b"))
        helpers.run_ref(["createsetdb"] + gz + [os.path.join(d, "a", "setDB"), os.path.join(d, "a", "tmp"), "-v", "1"])
        subprocess.run([SDB200, "createsetdb"] + gz + [os.path.join(d, "b", "setDB"), os.path.join(d, "b", "tmp")], check=True, stdout=subprocess.DEVNULL)
        assert _sha(os.path.join(d, "a")) == _sha(os.path.join(d, "b"))
        assert len(_sha(os.path.join(d, "a"))) == 19
    finally:
        shutil.rmtree(d, ignore_errors=True)


def test_reference_child_command_lines_are_accepted(synth_small):
    """the exact child command lines of the reference's clustersearch workflow (printed with -v 3) replayed through
    sdb200, flag for flag.  The host-only modules must reproduce the reference's DBs; the modules that need the GPU must get past
    their flag checks (without a device they stop at sd_context_create, never at an unknown flag)."""
    if not helpers.have_ref():
        pytest.skip("oracle/_ref/spacedust not built")
    import shlex
    d = tempfile.mkdtemp(prefix="sdfe_")
    try:
        paths = synth.write_fasta(synth_small, os.path.join(d, "faa"))
        helpers.run_ref(["createsetdb"] + paths + ["setDB", "tmp", "-v", "1"], cwd=d)
        log = helpers.run_ref(["clustersearch", "setDB", "setDB", "ref.tsv", "tmpcs", "--filter-self-match", "1", "--threads", "2", "-v", "3"], cwd=d)
        mods = ("prefilter", "align", "prefixid", "besthitbyset", "mergeresultsbyset", "combinehits", "clusterhits", "summarizeresults", "filterdb", "swapdb")
        lines = [l for l in log.splitlines() if l.split(" ")[0] in mods and " setDB" in l or l.startswith(("prefixid ", "filterdb ", "swapdb "))]
        seen = set()
        for line in lines:
            args = shlex.split(line)
            mod = args[0]
            if mod in seen:
                continue
            seen.add(mod)
            L = os.path.dirname([a for a in args[1:] if "/" in a][0])
            # outputs go next to the reference's (suffix _m); inputs are the reference's own files
            out_pos = {"prefilter": 3, "align": 4, "prefixid": 2, "besthitbyset": 4, "mergeresultsbyset": 3, "combinehits": 4, "clusterhits": 4,
                       "summarizeresults": 4, "filterdb": 2, "swapdb": 2}[mod]
            ref_out = args[out_pos]
            args[out_pos] = ref_out + "_m"
            r = subprocess.run([SDB200] + args, cwd=d, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
            if mod in ("prefilter", "align", "clusterhits"):
                assert r.returncode == 0 or b"no CUDA device" in r.stderr, (mod, r.stderr[-300:])
                continue
            assert r.returncode == 0, (mod, r.stderr[-300:])
            if mod == "summarizeresults":
                a = sorted(l for l in open(os.path.join(d, ref_out)) if not l.startswith("#"))
                b = sorted(l for l in open(os.path.join(d, ref_out + "_m")) if not l.startswith("#"))
                assert a == b and len(a) > 5
            elif mod == "combinehits":
                def pairs(p):
                    return sorted(zip([x[1] for x in helpers.read_db_entries(p + "_h")], [x[1] for x in helpers.read_db_entries(p)]))
                assert pairs(os.path.join(d, ref_out)) == pairs(os.path.join(d, ref_out + "_m"))
            elif mod == "swapdb":
                a = {k: sorted(v.split()) for k, v in helpers.read_db(os.path.join(d, ref_out)).items()}
                b = {k: sorted(v.split()) for k, v in helpers.read_db(os.path.join(d, ref_out + "_m")).items()}
                assert a == b
            else:
                assert helpers.read_db(os.path.join(d, ref_out)) == helpers.read_db(os.path.join(d, ref_out + "_m")), mod
            assert L
        assert seen == set(mods), seen
    finally:
        shutil.rmtree(d, ignore_errors=True)
