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


def align_fixture(n_proteomes=2, genes=60, seed=3):
    """Reference `align` result DB (text lines per query key) on the reference's own prefilter DB."""
    import gzip
    import json
    d = tempfile.mkdtemp(prefix="sdgold_")
    prot = synth.generate(n_proteomes, genes, seed=seed)
    paths = synth.write_fasta(prot, os.path.join(d, "faa"))
    db = os.path.join(d, "setDB")
    helpers.run_ref(["createsetdb"] + paths + [db, os.path.join(d, "tmp"), "-v", "1"])
    helpers.run_ref(["prefilter", db, db, os.path.join(d, "pref"), "-s", "5.7", "--max-seqs", "300", "-c", "0.8",
                     "--cov-mode", "2", "--min-ungapped-score", "15", "--threads", "1", "-v", "1"])
    helpers.run_ref(["align", db, db, os.path.join(d, "pref"), os.path.join(d, "result"), "-a", "1", "--alignment-mode", "2",
                     "-e", "10", "--min-aln-len", "30", "-c", "0.8", "--cov-mode", "2", "--threads", "1", "-v", "1"])
    ref = helpers.read_db(os.path.join(d, "result"))
    out = {"n_proteomes": n_proteomes, "genes": genes, "seed": seed,
           "result": {str(k): [l.decode() for l in v.split(b"\n") if l] for k, v in sorted(ref.items())}}
    with gzip.open(os.path.join(HERE, "align_syn%dx%d.json.gz" % (n_proteomes, genes)), "wt") as f:
        json.dump(out, f)
    shutil.rmtree(d, ignore_errors=True)
    print("align fixture: %d alignments" % sum(len(v) for v in out["result"].values()))


def clustersearch_fixture(n_proteomes=3, genes=120, seed=7):
    """Final TSV of the reference `clustersearch --filter-self-match 1`, canonicalised (helpers.parse_result_tsv)."""
    import gzip
    import json
    d = tempfile.mkdtemp(prefix="sdgold_")
    prot = synth.generate(n_proteomes, genes, seed=seed)
    paths = synth.write_fasta(prot, os.path.join(d, "faa"))
    db = os.path.join(d, "setDB")
    helpers.run_ref(["createsetdb"] + paths + [db, os.path.join(d, "tmp"), "-v", "1"])
    out = os.path.join(d, "res.tsv")
    helpers.run_ref(["clustersearch", db, db, out, os.path.join(d, "tmpcs"), "--filter-self-match", "1", "--threads", "2", "-v", "1"])
    src, lookup = helpers.read_source(db), helpers.read_lookup(db)
    canon = helpers.parse_result_tsv(out, src, {v: k for k, v in lookup.items()})
    with gzip.open(os.path.join(HERE, "clusters_syn%dx%d.json.gz" % (n_proteomes, genes)), "wt") as f:
        json.dump({"n_proteomes": n_proteomes, "genes": genes, "seed": seed, "clusters": canon}, f)
    shutil.rmtree(d, ignore_errors=True)
    print("clustersearch fixture: %d clusters" % len(canon))


def createsetdb_fixture(n_proteomes=2, genes=5, seed=3):
    """sha256 of every file `createsetdb` leaves behind (the on-disk contract of the set DB)."""
    import hashlib
    import json
    d = tempfile.mkdtemp(prefix="sdgold_")
    prot = synth.generate(n_proteomes, genes, seed=seed)
    paths = synth.write_fasta(prot, os.path.join(d, "faa"))
    helpers.run_ref(["createsetdb"] + paths + [os.path.join(d, "setDB"), os.path.join(d, "tmp"), "-v", "1"])
    files = {f: hashlib.sha256(open(os.path.join(d, f), "rb").read()).hexdigest() for f in sorted(os.listdir(d)) if f.startswith("setDB")}
    with open(os.path.join(HERE, "createsetdb_syn%dx%d.json" % (n_proteomes, genes)), "w") as f:
        json.dump({"n_proteomes": n_proteomes, "genes": genes, "seed": seed, "sha256": files}, f, indent=1)
    shutil.rmtree(d, ignore_errors=True)
    print("createsetdb fixture: %d files" % len(files))


def examples_fixture():
    """The reference's own example (examples/NC_000913.faa + NC_000915.faa, BASELINE config[0]): gzipped copies of the two input
    proteomes (public NCBI data, not reference source) and the reference's final TSV in canonical form (cluster numbers dropped,
    lines sorted) -- tests/test_scale_gpu.py::test_c1_reference_example_through_frontend runs the same input through sdb200."""
    import gzip
    ex = os.path.join(HERE, "examples")
    os.makedirs(ex, exist_ok=True)
    d = tempfile.mkdtemp(prefix="sdgold_")
    paths = []
    for n in ("NC_000913.faa", "NC_000915.faa"):
        src = os.path.join("/root/reference/examples", n)
        with open(src, "rb") as f, gzip.GzipFile(os.path.join(ex, n + ".gz"), "wb", 9, mtime=0) as g:
            shutil.copyfileobj(f, g)
        paths.append(src)
    db = os.path.join(d, "setDB")
    helpers.run_ref(["createsetdb"] + paths + [db, os.path.join(d, "tmp"), "-v", "1"])
    out = os.path.join(d, "res.tsv")
    helpers.run_ref(["clustersearch", db, db, out, os.path.join(d, "tmpcs"), "--filter-self-match", "1", "--threads", "8", "-v", "1"])
    lines = [l for l in open(out).read().split("\n") if l]
    canon = sorted("\t".join(l.split("\t")[1:]) if l.startswith("#") else l for l in lines)
    open(os.path.join(ex, "result_canonical.tsv"), "w").write("\n".join(canon) + "\n")
    print("examples fixture: %d clusters, %d hit lines" % (sum(1 for l in lines if l.startswith("#")), sum(1 for l in lines if l.startswith(">"))))
    shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    if not helpers.have_ref():
        sys.exit("oracle/_ref/spacedust missing: run python oracle/build_ref.py")
    prefilter_fixture()
    align_fixture()
    clustersearch_fixture()
    examples_fixture()
    createsetdb_fixture()
