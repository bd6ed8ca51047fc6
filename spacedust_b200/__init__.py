"""spacedust_b200 -- B200 (sm_100a) implementation of the Spacedust `clustersearch` hot path.

This package is a thin ctypes view of ``libsdgpu.so`` (CUDA kernels + C ABI, ``include/sdgpu.h``) used
by the tests and by ``bench.py``; the product itself is the shared library and the ``sdb200`` front-end
binary built from ``spacedust_b200/csrc``.  There is no CPU fallback: every compute call raises
``SdError`` when the library or a CUDA device is missing.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# SD_LIB_PATH: another build of the same library (variant builds for A/B measurements, csrc/Makefile)
LIB_PATH = os.environ.get("SD_LIB_PATH") or os.path.join(_HERE, "libsdgpu.so")
CSRC = os.path.join(_HERE, "csrc")


class SdError(RuntimeError):
    pass


class PrefilterParams(ctypes.Structure):
    """Mirrors sd_prefilter_params (flags of the reference's `prefilter` module)."""
    _fields_ = [("sensitivity", ctypes.c_float), ("max_seqs", ctypes.c_int), ("min_ungapped_score", ctypes.c_int),
                ("cov_thr", ctypes.c_float), ("cov_mode", ctypes.c_int), ("mask", ctypes.c_int),
                ("mask_prob", ctypes.c_float), ("comp_bias_corr", ctypes.c_int), ("bin_size", ctypes.c_int),
                ("same_db", ctypes.c_int)]


class PrefilterStats(ctypes.Structure):
    _fields_ = [("probes", ctypes.c_uint64), ("matches", ctypes.c_uint64), ("diag_residues", ctypes.c_uint64),
                ("query_residues", ctypes.c_uint64), ("hits", ctypes.c_uint64), ("candidates", ctypes.c_uint64),
                ("truncated", ctypes.c_uint64), ("kernel_ms", ctypes.c_double), ("total_ms", ctypes.c_double)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


HIT_DTYPE = np.dtype([("query", "<u4"), ("target", "<u4"), ("score", "<i4"), ("diag", "<i2"), ("pad", "<i2")])

_lib = None


def build(verbose=False):
    """Compile libsdgpu.so in-tree (nvcc, sm_100a)."""
    r = subprocess.run(["make", "-C", CSRC], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if verbose or r.returncode != 0:
        print(r.stdout.decode())
    if r.returncode != 0:
        raise SdError("building libsdgpu.so failed")
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SdError("libsdgpu.so is missing (%s); run `python -c 'import __graft_entry__ as g; g.build()'` -- "
                          "there is no CPU fallback" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        L.sd_last_error.restype = ctypes.c_char_p
        L.sd_context_launches.restype = ctypes.c_uint64
        L.sd_target_index_entries.restype = ctypes.c_uint64
        L.sd_target_index_targets.restype = ctypes.c_uint32
        L.sd_search_result_sizes.restype = None
        L.sd_search_result_destroy.restype = None
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise SdError("libsdgpu error %d: %s" % (rc, lib().sd_last_error().decode()))


def default_prefilter_params(**kw):
    p = PrefilterParams()
    lib().sd_prefilter_params_default(ctypes.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


class Context:
    def __init__(self, device=-1):
        self._h = ctypes.c_void_p()
        _check(lib().sd_context_create(ctypes.c_int(device), ctypes.byref(self._h)))

    def close(self):
        if self._h:
            lib().sd_context_destroy(self._h)
            self._h = ctypes.c_void_p()

    @property
    def launches(self):
        return int(lib().sd_context_launches(self._h))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _seq_args(buf, off):
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.uint64)
    return buf, off


class TargetIndex:
    """Device-resident target DB + k-mer index (IndexBuilder::fillDatabase counterpart)."""

    def __init__(self, ctx, seqs, off, params):
        self.ctx = ctx
        self.seqs, self.off = _seq_args(seqs, off)
        self.n = len(self.off) - 1
        self._h = ctypes.c_void_p()
        _check(lib().sd_target_index_build(ctx._h, self.seqs.ctypes.data_as(ctypes.c_char_p),
                                           self.off.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(self.n),
                                           ctypes.byref(params), ctypes.byref(self._h)))

    @classmethod
    def load(cls, ctx, path, params):
        """Persisted index (sd_target_index_load): read + upload, nothing is recomputed."""
        self = cls.__new__(cls)
        self.ctx = ctx
        self.seqs = self.off = None
        self._h = ctypes.c_void_p()
        _check(lib().sd_target_index_load(ctx._h, path.encode(), ctypes.byref(params), ctypes.byref(self._h)))
        self.n = int(lib().sd_target_index_targets(self._h))
        return self

    def save(self, path, params):
        _check(lib().sd_target_index_save(self._h, ctypes.byref(params), path.encode()))

    @property
    def entries(self):
        return int(lib().sd_target_index_entries(self._h))

    def masked_residues(self):
        out = np.zeros(int(self.off[-1]), dtype=np.uint8)
        _check(lib().sd_target_index_masked(self._h, out.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint64(len(out))))
        return out

    def close(self):
        if self._h:
            lib().sd_target_index_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def prefilter(ctx, index, qseqs, qoff, params, query_first=0):
    """Body of the reference's prefilter query loop for a batch of queries -> (hits, stats)."""
    qseqs, qoff = _seq_args(qseqs, qoff)
    nq = len(qoff) - 1
    cap = max(1, nq * min(max(params.max_seqs, 1), index.n))
    hits = np.zeros(cap, dtype=HIT_DTYPE)
    n = ctypes.c_uint64(0)
    st = PrefilterStats()
    _check(lib().sd_prefilter_run(ctx._h, index._h, qseqs.ctypes.data_as(ctypes.c_char_p),
                                  qoff.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(nq), ctypes.c_uint32(query_first),
                                  ctypes.byref(params), hits.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint64(cap),
                                  ctypes.byref(n), ctypes.byref(st)))
    return hits[:n.value], st.as_dict()


# ----------------------------------------------------------------------------- align
class AlignParams(ctypes.Structure):
    """Mirrors sd_align_params (flags of the reference's `align` module)."""
    _fields_ = [("evalue_thr", ctypes.c_double), ("min_seq_id", ctypes.c_double), ("cov_thr", ctypes.c_float),
                ("cov_mode", ctypes.c_int), ("min_aln_len", ctypes.c_int), ("comp_bias_corr", ctypes.c_int),
                ("gap_open", ctypes.c_int), ("gap_extend", ctypes.c_int), ("score_bias", ctypes.c_float),
                ("same_db", ctypes.c_int)]


class AlignStats(ctypes.Structure):
    _fields_ = [("alignments", ctypes.c_uint64), ("passed", ctypes.c_uint64), ("word_mode", ctypes.c_uint64),
                ("reverse_passes", ctypes.c_uint64), ("tracebacks", ctypes.c_uint64), ("cells_forward", ctypes.c_uint64),
                ("cells_reverse", ctypes.c_uint64), ("cells_traceback", ctypes.c_uint64), ("sw_ms", ctypes.c_double),
                ("traceback_ms", ctypes.c_double), ("total_ms", ctypes.c_double), ("target_residues", ctypes.c_uint64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


ALN_DTYPE = np.dtype([("query", "<u4"), ("target", "<u4"), ("score", "<i4"), ("bits", "<i4"), ("seq_id", "<f4"),
                      ("qcov", "<f4"), ("tcov", "<f4"), ("evalue", "<f8"), ("q_start", "<i4"), ("q_end", "<i4"),
                      ("q_len", "<i4"), ("t_start", "<i4"), ("t_end", "<i4"), ("t_len", "<i4"), ("aln_len", "<i4"),
                      ("bt_off", "<u8"), ("bt_len", "<u4"), ("bt_cigar", "<u4")], align=True)


def default_align_params(**kw):
    p = AlignParams()
    lib().sd_align_params_default(ctypes.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def align(ctx, index, qseqs, qoff, pair_q, pair_t, params, query_first=0, db_residues=0):
    """Body of the reference's align query loop for prefilter pairs -> (alignments, backtrace pool, stats)."""
    assert ALN_DTYPE.itemsize == 88, ALN_DTYPE.itemsize
    qseqs, qoff = _seq_args(qseqs, qoff)
    nq = len(qoff) - 1
    pair_q = np.ascontiguousarray(pair_q, dtype=np.uint32)
    pair_t = np.ascontiguousarray(pair_t, dtype=np.uint32)
    n_pairs = len(pair_q)
    cap = max(1, n_pairs)
    lens_t = np.diff(index.off).astype(np.int64)
    lens_q = np.diff(qoff).astype(np.int64)
    # (indices are clipped only for this size estimate: the library itself rejects out-of-range pairs)
    bt_cap = int((lens_t[np.minimum(pair_t, len(lens_t) - 1)] + lens_q[np.minimum(pair_q, len(lens_q) - 1)]).sum()) + 64 if n_pairs else 64
    out = np.zeros(cap, dtype=ALN_DTYPE)
    pool = np.zeros(bt_cap, dtype=np.uint8)
    n = ctypes.c_uint64(0)
    used = ctypes.c_uint64(0)
    st = AlignStats()
    _check(lib().sd_align_run(ctx._h, index._h, qseqs.ctypes.data_as(ctypes.c_char_p), qoff.ctypes.data_as(ctypes.c_void_p),
                              ctypes.c_uint32(nq), ctypes.c_uint32(query_first), pair_q.ctypes.data_as(ctypes.c_void_p),
                              pair_t.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint64(n_pairs), ctypes.c_uint64(db_residues),
                              ctypes.byref(params), out.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint64(cap), ctypes.byref(n),
                              pool.ctypes.data_as(ctypes.c_char_p), ctypes.c_uint64(bt_cap), ctypes.byref(used), ctypes.byref(st)))
    return out[:n.value], pool[:used.value], st.as_dict()


def format_alignment(aln, target_key, pool):
    buf = ctypes.create_string_buffer(1 << 18)
    rec = np.ascontiguousarray(aln.reshape(1))
    n = lib().sd_format_alignment(rec.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(int(target_key)),
                                  pool.ctypes.data_as(ctypes.c_char_p), buf, ctypes.c_int(1 << 18))
    if n <= 0:
        raise SdError("sd_format_alignment failed")
    return buf.raw[:n]


# ----------------------------------------------------------------------------- besthitbyset / combinehits / clusterhits
class ClusterParams(ctypes.Structure):
    """Mirrors sd_cluster_params (flags of combinehits / clusterhits)."""
    _fields_ = [("alpha", ctypes.c_float), ("filter_self_match", ctypes.c_int), ("max_gene_gap", ctypes.c_int),
                ("cluster_size", ctypes.c_int), ("cluster_pval", ctypes.c_double), ("multihit_pval", ctypes.c_double)]


class ClusterStats(ctypes.Structure):
    _fields_ = [("best_hits", ctypes.c_uint64), ("genome_pairs", ctypes.c_uint64), ("pairs_clustered", ctypes.c_uint64),
                ("max_hits_per_pair", ctypes.c_uint64), ("merges", ctypes.c_uint64), ("inexact_log", ctypes.c_uint64),
                ("kernel_ms", ctypes.c_double), ("total_ms", ctypes.c_double)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


CLUSTER_DTYPE = np.dtype([("q_set", "<u4"), ("t_set", "<u4"), ("p_co", "<f8"), ("p_mh", "<f8"), ("n", "<u4"), ("first", "<u4")],
                         align=True)
MEMBER_DTYPE = np.dtype([("aln", "<u4"), ("pad", "<u4"), ("pval", "<f8")], align=True)


def default_cluster_params(**kw):
    p = ClusterParams()
    lib().sd_cluster_params_default(ctypes.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def cluster(ctx, alns, q_meta, t_meta, params):
    """besthitbyset + combinehits + clusterhits on accepted alignments.
    q_meta / t_meta: dict(set_of=uint32[n], pos=uint32[n], strand=uint8[n], set_size=uint32[nSets])."""
    alns = np.ascontiguousarray(alns, dtype=ALN_DTYPE)

    def arr(x, dt):
        return np.ascontiguousarray(x, dtype=dt)
    qs, qp, qst, qsz = arr(q_meta["set_of"], np.uint32), arr(q_meta["pos"], np.uint32), arr(q_meta["strand"], np.uint8), arr(q_meta["set_size"], np.uint32)
    ts, tp, tst, tsz = arr(t_meta["set_of"], np.uint32), arr(t_meta["pos"], np.uint32), arr(t_meta["strand"], np.uint8), arr(t_meta["set_size"], np.uint32)
    cap = max(1, len(alns))
    clusters = np.zeros(cap, dtype=CLUSTER_DTYPE)
    members = np.zeros(cap, dtype=MEMBER_DTYPE)
    nc, nm = ctypes.c_uint64(0), ctypes.c_uint64(0)
    st = ClusterStats()
    p = ctypes.c_void_p
    _check(lib().sd_cluster_run(ctx._h, alns.ctypes.data_as(p), ctypes.c_uint64(len(alns)),
                                qs.ctypes.data_as(p), qp.ctypes.data_as(p), qst.ctypes.data_as(p), ctypes.c_uint32(len(qs)),
                                ts.ctypes.data_as(p), tp.ctypes.data_as(p), tst.ctypes.data_as(p), ctypes.c_uint32(len(ts)),
                                qsz.ctypes.data_as(p), ctypes.c_uint32(len(qsz)), tsz.ctypes.data_as(p), ctypes.c_uint32(len(tsz)),
                                ctypes.byref(params), clusters.ctypes.data_as(p), ctypes.c_uint64(cap), ctypes.byref(nc),
                                members.ctypes.data_as(p), ctypes.c_uint64(cap), ctypes.byref(nm), ctypes.byref(st)))
    return clusters[:nc.value], members[:nm.value], st.as_dict()


# ----------------------------------------------------------------------------- device-resident query batches
class QueryBatch:
    """Query residues uploaded and mapped once (sd_query_batch); reusable by prefilter_batch / align_batch."""

    def __init__(self, ctx, seqs, off):
        self.ctx = ctx
        self.seqs, self.off = _seq_args(seqs, off)
        self.n = len(self.off) - 1
        self._h = ctypes.c_void_p()
        _check(lib().sd_query_batch_create(ctx._h, self.seqs.ctypes.data_as(ctypes.c_char_p),
                                           self.off.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(self.n), ctypes.byref(self._h)))

    def close(self):
        if self._h:
            lib().sd_query_batch_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def prefilter_batch(ctx, index, batch, params, query_first=0, out=None):
    cap = max(1, batch.n * min(max(params.max_seqs, 1), index.n))
    hits = out if out is not None and len(out) >= cap else np.zeros(cap, dtype=HIT_DTYPE)
    n = ctypes.c_uint64(0)
    st = PrefilterStats()
    _check(lib().sd_prefilter_run_batch(ctx._h, index._h, batch._h, ctypes.c_uint32(query_first), ctypes.byref(params),
                                        hits.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint64(len(hits)), ctypes.byref(n), ctypes.byref(st)))
    return hits[:n.value], st.as_dict()


def align_batch(ctx, index, batch, pair_q, pair_t, params, query_first=0, db_residues=0):
    pair_q = np.ascontiguousarray(pair_q, dtype=np.uint32)
    pair_t = np.ascontiguousarray(pair_t, dtype=np.uint32)
    n_pairs = len(pair_q)
    # accepted alignments are a small fraction of the pairs: start small, grow on SD_ESMALL
    cap = max(1024, n_pairs // 8)
    bt_cap = max(1 << 16, cap * 512)
    n = ctypes.c_uint64(0)
    used = ctypes.c_uint64(0)
    st = AlignStats()
    for _ in range(2):
        out = np.zeros(cap, dtype=ALN_DTYPE)
        pool = np.zeros(bt_cap, dtype=np.uint8)
        rc = lib().sd_align_run_batch(ctx._h, index._h, batch._h, ctypes.c_uint32(query_first), pair_q.ctypes.data_as(ctypes.c_void_p),
                                      pair_t.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint64(n_pairs), ctypes.c_uint64(db_residues),
                                      ctypes.byref(params), out.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint64(cap), ctypes.byref(n),
                                      pool.ctypes.data_as(ctypes.c_char_p), ctypes.c_uint64(bt_cap), ctypes.byref(used), ctypes.byref(st))
        if rc == -5:     # SD_ESMALL: sizes are reported back
            cap, bt_cap = int(n.value) + 16, int(used.value) + 64
            continue
        _check(rc)
        return out[:n.value], pool[:used.value], st.as_dict()
    raise SdError("sd_align_run_batch: output buffers too small twice")


def clustersearch_staged(ctx, index, batch, q_meta, t_meta, pref_params=None, aln_params=None, clu_params=None, query_first=0):
    """The hot path stage by stage through the three C-ABI calls (hit lists travel through the host)."""
    pp = pref_params or default_prefilter_params(same_db=1)
    ap = aln_params or default_align_params(same_db=pp.same_db)
    cp = clu_params or default_cluster_params(filter_self_match=1)
    hits, pst = prefilter_batch(ctx, index, batch, pp, query_first)
    alns, pool, ast = align_batch(ctx, index, batch, hits["query"], hits["target"], ap, query_first)
    clusters, members, cst = cluster(ctx, alns, q_meta, t_meta, cp)
    return {"hits": hits, "alns": alns, "pool": pool, "clusters": clusters, "members": members,
            "prefilter_stats": pst, "align_stats": ast, "cluster_stats": cst}


NCCL_ID_BYTES = 128


def shard_unique_id():
    """ncclGetUniqueId through the C ABI (rank 0; the caller sends the 128 bytes to the other ranks)."""
    buf = (ctypes.c_uint8 * NCCL_ID_BYTES)()
    _check(lib().sd_shard_unique_id(buf))
    return bytes(buf)


def shard_init(ctx, rank, world, unique_id):
    """One process per GPU: NCCL communicator of this rank inside the context (sd_shard_init)."""
    assert len(unique_id) == NCCL_ID_BYTES
    buf = (ctypes.c_uint8 * NCCL_ID_BYTES).from_buffer_copy(unique_id)
    _check(lib().sd_shard_init(ctx._h, ctypes.c_int(rank), ctypes.c_int(world), buf))


def shard_partition(set_residues, world):
    """sd_shard_partition: contiguous residue-balanced blocks of query sets -> [(lo, hi)] * world."""
    w = np.ascontiguousarray(set_residues, dtype=np.uint64)
    first = np.zeros(world + 1, dtype=np.uint32)
    _check(lib().sd_shard_partition(w.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(len(w)), ctypes.c_int(world),
                                    first.ctypes.data_as(ctypes.c_void_p)))
    return [(int(first[i]), int(first[i + 1])) for i in range(world)]


def shard_partition_genes(gene_off, world):
    """sd_shard_partition_genes: contiguous residue-balanced blocks of query genes -> [(lo, hi)] * world."""
    o = np.ascontiguousarray(gene_off, dtype=np.uint64)
    first = np.zeros(world + 1, dtype=np.uint32)
    _check(lib().sd_shard_partition_genes(o.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(len(o) - 1), ctypes.c_int(world),
                                          first.ctypes.data_as(ctypes.c_void_p)))
    return [(int(first[i]), int(first[i + 1])) for i in range(world)]


def clustersearch(ctx, index, batch, q_meta, t_meta, pref_params=None, aln_params=None, clu_params=None, query_first=0,
                  db_residues=0, out=None, gather=False, shard_genes=False):
    """The hot path of `spacedust clustersearch` after createsetdb for one query batch, fused (sd_clustersearch_run):
    prefilter -> align -> besthitbyset/combinehits -> clusterhits with the prefilter hit list kept in HBM.
    out: optional object with carve(n_alns, pool_bytes, n_clusters, n_members) -> (alns, pool, clusters, members) arrays to
    receive the result (e.g. shard.ResultExchange: persistent pinned memory, no per-call allocation).
    gather=True (after shard_init): this rank's block of query sets is searched, then sd_shard_allgather_result (ONE ncclAllGather
    over NVLink, inside the library) returns every rank's lists with global query indices, in the order of a single-GPU run."""
    pp = pref_params or default_prefilter_params(same_db=1)
    ap = aln_params or default_align_params(same_db=pp.same_db)
    cp = clu_params or default_cluster_params(filter_self_match=1)

    def arr(x, dt):
        return np.ascontiguousarray(x, dtype=dt)
    qs, qp, qst, qsz = arr(q_meta["set_of"], np.uint32), arr(q_meta["pos"], np.uint32), arr(q_meta["strand"], np.uint8), arr(q_meta["set_size"], np.uint32)
    ts, tp, tst, tsz = arr(t_meta["set_of"], np.uint32), arr(t_meta["pos"], np.uint32), arr(t_meta["strand"], np.uint8), arr(t_meta["set_size"], np.uint32)
    assert (shard_genes or len(qs) == batch.n) and len(ts) == index.n
    res = ctypes.c_void_p()
    pst, ast, cst = PrefilterStats(), AlignStats(), ClusterStats()
    p = ctypes.c_void_p
    L = lib()
    gathered = ctypes.c_uint64(0)
    if shard_genes:
        # gene-sharded multi-GPU step (after shard_init): batch = this rank's genes query_first.., q_meta describes ALL query genes;
        # both exchanges happen inside the library (sd_shard_clustersearch_run)
        assert query_first + batch.n <= len(qs)
        _check(L.sd_shard_clustersearch_run(ctx._h, index._h, batch._h, ctypes.c_uint32(query_first), ctypes.byref(pp), ctypes.byref(ap),
                                            ctypes.byref(cp), qs.ctypes.data_as(p), qp.ctypes.data_as(p), qst.ctypes.data_as(p),
                                            ctypes.c_uint32(len(qs)), ts.ctypes.data_as(p), tp.ctypes.data_as(p), tst.ctypes.data_as(p),
                                            ctypes.c_uint32(len(ts)), qsz.ctypes.data_as(p), ctypes.c_uint32(len(qsz)),
                                            tsz.ctypes.data_as(p), ctypes.c_uint32(len(tsz)), ctypes.c_uint64(db_residues),
                                            ctypes.byref(res), ctypes.byref(pst), ctypes.byref(ast), ctypes.byref(cst), ctypes.byref(gathered)))
        gather = False
    else:
        _check(L.sd_clustersearch_run(ctx._h, index._h, batch._h, ctypes.c_uint32(query_first), ctypes.byref(pp), ctypes.byref(ap),
                                      ctypes.byref(cp), qs.ctypes.data_as(p), qp.ctypes.data_as(p), qst.ctypes.data_as(p),
                                      ts.ctypes.data_as(p), tp.ctypes.data_as(p), tst.ctypes.data_as(p), ctypes.c_uint32(len(ts)),
                                      qsz.ctypes.data_as(p), ctypes.c_uint32(len(qsz)), tsz.ctypes.data_as(p), ctypes.c_uint32(len(tsz)),
                                      ctypes.c_uint64(db_residues), ctypes.byref(res), ctypes.byref(pst), ctypes.byref(ast), ctypes.byref(cst)))
    if gather:
        merged = ctypes.c_void_p()
        try:
            _check(L.sd_shard_allgather_result(ctx._h, res, ctypes.c_uint32(query_first), ctypes.byref(merged), ctypes.byref(gathered)))
        finally:
            L.sd_search_result_destroy(res)
        res = merged
    try:
        na, nb, nc, nm = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
        L.sd_search_result_sizes(res, ctypes.byref(na), ctypes.byref(nb), ctypes.byref(nc), ctypes.byref(nm))
        if out is not None:
            alns, pool, clusters, members = out.carve(na.value, nb.value, nc.value, nm.value)
        else:
            alns = np.empty(max(1, na.value), dtype=ALN_DTYPE)
            pool = np.empty(max(1, nb.value), dtype=np.uint8)
            clusters = np.empty(max(1, nc.value), dtype=CLUSTER_DTYPE)
            members = np.empty(max(1, nm.value), dtype=MEMBER_DTYPE)
        _check(L.sd_search_result_copy(res, alns.ctypes.data_as(p), pool.ctypes.data_as(ctypes.c_char_p), clusters.ctypes.data_as(p),
                                       members.ctypes.data_as(p)))
    finally:
        L.sd_search_result_destroy(res)
    pst_d = pst.as_dict()
    return {"alns": alns[:na.value], "pool": pool[:nb.value], "clusters": clusters[:nc.value], "members": members[:nm.value],
            "n_prefilter_hits": None, "prefilter_stats": pst_d, "align_stats": ast.as_dict(), "cluster_stats": cst.as_dict(),
            "bytes_gathered": int(gathered.value)}
