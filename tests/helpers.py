"""Shared test helpers: MMseqs2 flat-file DB access, the oracle (ctypes), the reference binary."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_BIN = os.path.join(ORACLE_DIR, "_ref", "spacedust")
GOLDEN = os.path.join(ROOT, "tests", "golden")


# ----------------------------------------------------------------------------- DB files
def read_index(path):
    """<db>.index -> list of (key, offset, length) in file order."""
    out = []
    with open(path + ".index") as f:
        for line in f:
            k, o, l = line.split("\t")
            out.append((int(k), int(o), int(l)))
    return out


def read_db(path):
    """Return {key: bytes-without-trailing-NUL}.  Accepts merged DBs and per-thread splits
    (<db>.0 .. <db>.N-1 concatenated in order, as DBReader does)."""
    if os.path.exists(path):
        data = open(path, "rb").read()
    else:
        parts = []
        i = 0
        while os.path.exists("%s.%d" % (path, i)):
            parts.append(open("%s.%d" % (path, i), "rb").read())
            i += 1
        data = b"".join(parts)
    out = {}
    for k, o, l in read_index(path):
        out[k] = data[o:o + l - 1]
    return out


def read_db_entries(path):
    """Entries in index order, keys may repeat (matches/clusters DBs)."""
    if os.path.exists(path):
        data = open(path, "rb").read()
    else:
        parts = []
        i = 0
        while os.path.exists("%s.%d" % (path, i)):
            parts.append(open("%s.%d" % (path, i), "rb").read())
            i += 1
        data = b"".join(parts)
    return [(k, data[o:o + l - 1]) for k, o, l in read_index(path)]


def load_seqdb(path):
    """Sequence DB -> (keys, concatenated ASCII residues (np.uint8), offsets (np.uint64, n+1)).
    Order = ascending data offset (DBReader LINEAR_ACCCESS), residues exclude the trailing newline."""
    idx = sorted(read_index(path), key=lambda e: e[1])
    data = open(path, "rb").read()
    keys = np.array([e[0] for e in idx], dtype=np.uint32)
    lens = np.array([e[2] - 2 for e in idx], dtype=np.uint64)
    off = np.zeros(len(idx) + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    buf = np.empty(int(off[-1]), dtype=np.uint8)
    raw = np.frombuffer(data, dtype=np.uint8)
    for i, (k, o, l) in enumerate(idx):
        buf[int(off[i]):int(off[i + 1])] = raw[o:o + l - 2]
    return keys, buf, off


def read_fasta(path):
    names, seqs = [], []
    cur = []
    for line in open(path):
        line = line.rstrip("\n")
        if line.startswith(">"):
            if names:
                seqs.append("".join(cur))
            names.append(line[1:])
            cur = []
        else:
            cur.append(line)
    if names:
        seqs.append("".join(cur))
    return names, seqs


def pack_seqs(seqs):
    lens = np.array([len(s) for s in seqs], dtype=np.uint64)
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    buf = np.frombuffer("".join(seqs).encode(), dtype=np.uint8).copy()
    return buf, off


# ----------------------------------------------------------------------------- oracle
class OrcPrefParams(ctypes.Structure):
    _fields_ = [("sensitivity", ctypes.c_float), ("maxSeqs", ctypes.c_int), ("minDiagScore", ctypes.c_int),
                ("covThr", ctypes.c_float), ("covMode", ctypes.c_int), ("mask", ctypes.c_int),
                ("maskProb", ctypes.c_float), ("compBias", ctypes.c_int), ("binSize", ctypes.c_int),
                ("sameDB", ctypes.c_int), ("maxDbMatches", ctypes.c_uint64), ("foundDiagonalsSize", ctypes.c_uint64)]


class OrcPrefStats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint64) for n in
                ("probes", "matches", "diagResidues", "queryResidues", "hits", "candidates", "truncated",
                 "overflowQueries", "multiOverflowQueries")]


ORC_HIT_DTYPE = np.dtype([("query", "<u4"), ("target", "<u4"), ("score", "<i4"), ("diag", "<i2")], align=True)

_orc = None


def build_oracle():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)
    return os.path.join(ORACLE_DIR, "_build", "liborc.so")


def oracle():
    global _orc
    if _orc is None:
        path = os.path.join(ORACLE_DIR, "_build", "liborc.so")
        if not os.path.exists(path):
            build_oracle()
        _orc = ctypes.CDLL(path)
        _orc.orc_prefilter.restype = ctypes.c_long
    return _orc


def default_pref_params(**kw):
    p = OrcPrefParams(sensitivity=5.7, maxSeqs=300, minDiagScore=15, covThr=0.8, covMode=2, mask=1,
                      maskProb=0.9, compBias=1, binSize=2, sameDB=1, maxDbMatches=0, foundDiagonalsSize=0)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def orc_prefilter(tbuf, toff, qbuf, qoff, params, want_masked=False):
    lib = oracle()
    nT, nQ = len(toff) - 1, len(qoff) - 1
    cap = max(1024, nQ * (params.maxSeqs + 1))
    out = np.zeros(cap, dtype=ORC_HIT_DTYPE)
    stats = OrcPrefStats()
    masked = np.zeros(len(tbuf), dtype=np.uint8) if want_masked else None
    n = lib.orc_prefilter(tbuf.ctypes.data_as(ctypes.c_char_p), toff.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(nT),
                          qbuf.ctypes.data_as(ctypes.c_char_p), qoff.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(nQ),
                          ctypes.byref(params), out.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(cap),
                          ctypes.byref(stats), masked.ctypes.data_as(ctypes.c_void_p) if want_masked else None)
    if n < 0:
        raise RuntimeError("orc_prefilter failed: %d" % n)
    st = {f[0]: getattr(stats, f[0]) for f in OrcPrefStats._fields_}
    return (out[:n], st, masked) if want_masked else (out[:n], st)


def load_profiledb(path):
    """Profile DB (dbtype 2) -> (keys, profile bytes (np.uint8, 25 per column, entries concatenated), offsets in columns).
    Entry = L * 25 bytes + terminator (Sequence.h:458-471); order = ascending data offset like load_seqdb."""
    idx = sorted(read_index(path), key=lambda e: e[1])
    raw = np.frombuffer(open(path, "rb").read(), dtype=np.uint8)
    keys = np.array([e[0] for e in idx], dtype=np.uint32)
    cols = np.array([(e[2] - 1) // 25 for e in idx], dtype=np.uint64)
    off = np.zeros(len(idx) + 1, dtype=np.uint64)
    np.cumsum(cols, out=off[1:])
    buf = np.empty(int(off[-1]) * 25, dtype=np.uint8)
    for i, (k, o, l) in enumerate(idx):
        n = int(cols[i]) * 25
        buf[int(off[i]) * 25:int(off[i]) * 25 + n] = raw[o:o + n]
    return keys, buf, off


def orc_prefilter_profile(tbuf, toff, pbuf, poff, params):
    lib = oracle()
    lib.orc_prefilter_profile.restype = ctypes.c_long
    nT, nQ = len(toff) - 1, len(poff) - 1
    cap = max(1024, nQ * (params.maxSeqs + 1))
    out = np.zeros(cap, dtype=ORC_HIT_DTYPE)
    stats = OrcPrefStats()
    n = lib.orc_prefilter_profile(tbuf.ctypes.data_as(ctypes.c_char_p), toff.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(nT),
                                  pbuf.ctypes.data_as(ctypes.c_void_p), poff.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(nQ),
                                  ctypes.byref(params), out.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(cap), ctypes.byref(stats))
    if n < 0:
        raise RuntimeError("orc_prefilter_profile failed: %d" % n)
    return out[:n], {f[0]: getattr(stats, f[0]) for f in OrcPrefStats._fields_}


def hits_to_lines(hits, qkeys, tkeys):
    """-> {queryKey: [b'tKey\\tscore\\tdiag', ...]} in output order (prefilterHitToBuffer QueryMatcher.h:118-130)."""
    d = {int(k): [] for k in qkeys}
    for h in hits:
        d[int(qkeys[h["query"]])].append(b"%d\t%d\t%d" % (int(tkeys[h["target"]]), int(h["score"]), int(h["diag"])))
    return d


# ----------------------------------------------------------------------------- reference binary
def have_ref():
    return os.path.exists(REF_BIN) and os.access(REF_BIN, os.X_OK)


def run_ref(args, cwd=None, threads=None):
    env = dict(os.environ)
    env["MMSEQS_FORCE_MERGE"] = "1"
    r = subprocess.run([REF_BIN] + [str(a) for a in args], cwd=cwd, env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT)
    if r.returncode != 0:
        raise RuntimeError("reference failed: %s\n%s" % (args, r.stdout.decode()[-2000:]))
    return r.stdout.decode()


def read_lookup(db):
    """<db>.lookup -> {key: name}."""
    out = {}
    for line in open(db + ".lookup"):
        a = line.rstrip("\n").split("\t")
        out[int(a[0])] = a[1]
    return out


def read_source(db):
    """<db>.source -> {filename: set key}."""
    out = {}
    for line in open(db + ".source"):
        a = line.rstrip("\n").split("\t")
        out[a[1]] = int(a[0])
    return out


# ----------------------------------------------------------------------------- oracle: align
class OrcAlignParams(ctypes.Structure):
    _fields_ = [("evalThr", ctypes.c_double), ("seqIdThr", ctypes.c_double), ("covThr", ctypes.c_float),
                ("covMode", ctypes.c_int), ("minAlnLen", ctypes.c_int), ("compBias", ctypes.c_int),
                ("gapOpen", ctypes.c_int), ("gapExtend", ctypes.c_int), ("scoreBias", ctypes.c_float),
                ("sameDB", ctypes.c_int)]


class OrcAlignStats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint64) for n in
                ("alignments", "passed", "wordMode", "reversePasses", "tracebacks", "tracebackFailed", "scoreMismatch",
                 "cellsForward", "cellsReverse", "cellsTraceback")]


ORC_ALN_DTYPE = np.dtype([("query", "<u4"), ("target", "<u4"), ("targetKey", "<u4"), ("score", "<i4"), ("bits", "<i4"),
                          ("seqId", "<f4"), ("qcov", "<f4"), ("tcov", "<f4"), ("evalue", "<f8"),
                          ("qStart", "<i4"), ("qEnd", "<i4"), ("qLen", "<i4"), ("tStart", "<i4"), ("tEnd", "<i4"),
                          ("tLen", "<i4"), ("alnLen", "<i4"), ("btOff", "<u8"), ("btLen", "<u4")], align=True)


def default_align_params(**kw):
    p = OrcAlignParams(evalThr=10.0, seqIdThr=0.0, covThr=0.8, covMode=2, minAlnLen=30, compBias=1, gapOpen=11,
                       gapExtend=1, scoreBias=0.0, sameDB=1)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def orc_align(tbuf, toff, tkeys, qbuf, qoff, qkeys, pair_q, pair_t, params, db_residues=None, profile=False):
    """profile=True: qbuf holds profile entries (25 bytes per column), qoff counts columns (orc_align_profile)."""
    lib = oracle()
    lib.orc_align.restype = ctypes.c_long
    lib.orc_align_profile.restype = ctypes.c_long
    nT, nQ = len(toff) - 1, len(qoff) - 1
    pair_q = np.ascontiguousarray(pair_q, dtype=np.uint32)
    pair_t = np.ascontiguousarray(pair_t, dtype=np.uint32)
    tkeys = np.ascontiguousarray(tkeys, dtype=np.uint32)
    qkeys = np.ascontiguousarray(qkeys, dtype=np.uint32)
    if db_residues is None:
        db_residues = int(toff[-1])
    cap = max(16, len(pair_q))
    out = np.zeros(cap, dtype=ORC_ALN_DTYPE)
    assert ORC_ALN_DTYPE.itemsize == 88, ORC_ALN_DTYPE.itemsize
    lens_t = np.diff(toff).astype(np.int64)
    lens_q = np.diff(qoff).astype(np.int64)
    bt_cap = int((lens_t[pair_t] + lens_q[pair_q]).sum()) + 64
    bt = np.zeros(bt_cap, dtype=np.uint8)
    stats = OrcAlignStats()
    fn = lib.orc_align_profile if profile else lib.orc_align
    n = fn(tbuf.ctypes.data_as(ctypes.c_char_p), toff.ctypes.data_as(ctypes.c_void_p), tkeys.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(nT),
                      qbuf.ctypes.data_as(ctypes.c_char_p), qoff.ctypes.data_as(ctypes.c_void_p), qkeys.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(nQ),
                      pair_q.ctypes.data_as(ctypes.c_void_p), pair_t.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint64(len(pair_q)),
                      ctypes.c_uint64(db_residues), ctypes.byref(params), out.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(cap),
                      bt.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint64(bt_cap), ctypes.byref(stats))
    if n < 0:
        raise RuntimeError("orc_align failed: %d" % n)
    st = {f[0]: getattr(stats, f[0]) for f in OrcAlignStats._fields_}
    return out[:n], bt, st


def format_alignments(alns, bt, qkeys):
    """-> {queryKey: [line bytes (no newline), ...]} exactly as Matcher::resultToBuffer prints them."""
    lib = oracle()
    buf = ctypes.create_string_buffer(1 << 20)
    d = {int(k): [] for k in qkeys}
    for i in range(len(alns)):
        rec = alns[i:i + 1]
        n = lib.orc_format_alignment(rec.ctypes.data_as(ctypes.c_void_p), bt.ctypes.data_as(ctypes.c_char_p), buf, ctypes.c_int(1 << 20))
        d[int(qkeys[alns[i]["query"]])].append(buf.raw[:n - 1])
    return d


def parse_pref_db(path, qkeys, tkeys):
    """prefilter DB -> (pair_q, pair_t) index arrays in DB order of queries = order of qkeys."""
    key2t = {int(k): i for i, k in enumerate(tkeys)}
    db = read_db(path)
    pq, pt = [], []
    for qi, k in enumerate(qkeys):
        for line in db.get(int(k), b"").split(b"\n"):
            if line:
                pq.append(qi)
                pt.append(key2t[int(line.split(b"\t")[0])])
    return np.array(pq, dtype=np.uint32), np.array(pt, dtype=np.uint32)


# ----------------------------------------------------------------------------- oracle: whole clustersearch path
def oracle_clustersearch(proteomes, filter_self_match=True, max_gene_gap=3, cluster_size=2, cluster_pval=0.01,
                         multihit_pval=0.01, alpha=1.0, alns_override=None, query_sets=None):
    """prefilter -> align -> prefixid -> besthitbyset -> mergeresultsbyset -> combinehits -> clusterhits with the
    oracle restatements; -> dict(alns, bt, result, matches, clusters) (text records as the reference exchanges them)."""
    import sys
    sys.path.insert(0, ORACLE_DIR)
    import orc_cluster as oc
    from spacedust_b200 import synth
    buf, off, _ = synth.pack(proteomes)
    meta = synth.set_metadata(proteomes)
    n = len(off) - 1
    keys = np.arange(n, dtype=np.uint32)
    if alns_override is None:
        hits, _ = orc_prefilter(buf, off, buf, off, default_pref_params())
        alns, bt, _ = orc_align(buf, off, keys, buf, off, keys, hits["query"], hits["target"], default_align_params())
    else:
        alns, bt = alns_override
    result = {k: [l.decode() for l in v] for k, v in format_alignments(alns, bt, keys).items()}
    if query_sets is not None:      # a rank's shard: only the query genes of its sets (targets stay whole)
        qs = set(int(x) for x in query_sets)
        result = {k: v for k, v in result.items() if int(meta["set_of"][k]) in qs}
    m2s = {i: int(meta["set_of"][i]) for i in range(n)}
    s2m = {}
    for i in range(n):
        if query_sets is None or int(meta["set_of"][i]) in qs:
            s2m.setdefault(int(meta["set_of"][i]), []).append(i)
    ssz = {i: int(s) for i, s in enumerate(meta["set_size"])}
    lookup = {i: meta["names"][i] for i in range(n)}
    agg = oc.besthitbyset(oc.prefixid(result), m2s)
    merged = oc.mergeresultsbyset(s2m, agg)
    matches = oc.combinehits(merged, m2s, ssz, ssz, alpha, filter_self_match)
    clusters = oc.clusterhits(matches, lookup, lookup, max(ssz.values()), d=max_gene_gap, cls=cluster_size,
                              p_clu_thr=cluster_pval, p_mh_thr=multihit_pval, alpha=alpha)
    return {"alns": alns, "bt": bt, "result": result, "matches": matches, "clusters": clusters, "meta": meta, "lookup": lookup}


def gpu_clusters(res):
    """Result of sd.clustersearch / sd.clustersearch_staged -> [(header, [lines])] in the clusterhits convention (query and target
    by key = index, P-values as printed), comparable through canonical_clusters()."""
    import spacedust_b200 as sd
    out = []
    for c in res["clusters"]:
        lines = []
        for m in res["members"][int(c["first"]):int(c["first"]) + int(c["n"])]:
            al = res["alns"][int(m["aln"])]
            line = sd.format_alignment(al, al["target"], res["pool"]).decode().rstrip("\n").split("\t")
            lines.append("\t".join([str(int(al["query"])), line[0], "%.3E" % m["pval"]] + line[2:]))
        out.append(("%d\t%d\t%.3E\t%.3E\t%d" % (c["q_set"], c["t_set"], c["p_co"], c["p_mh"], c["n"]), lines))
    return out


def canonical_clusters(clusters):
    """[(header, [lines])] -> sorted list of (qSet, tSet, pCO, pMH, n, tuple(lines)) (cluster keys/order are not deterministic)."""
    out = []
    for hdr, lines in clusters:
        h = hdr.split("\t")
        out.append((int(h[0]), int(h[1]), h[2], h[3], int(h[4]), tuple(lines)))
    return sorted(out)


def parse_result_tsv(path, source_to_set, name_to_key):
    """Final clustersearch TSV -> canonical clusters in the (header, lines) convention of clusterhits, with
    names mapped back to keys so that it can be compared with canonical_clusters()."""
    out, cur = [], None
    for line in open(path):
        line = line.rstrip("\n")
        if line.startswith("#"):
            f = line.split("\t")
            cur = [source_to_set[f[1]], source_to_set[f[2]], f[3], f[4], int(f[5]), []]
            out.append(cur)
        elif line.startswith(">"):
            f = line[1:].split("\t")
            cur[5].append("\t".join([str(name_to_key[f[0]]), str(name_to_key[f[1]])] + f[2:]))
    return sorted((a, b, c, d, e, tuple(l)) for a, b, c, d, e, l in out)
