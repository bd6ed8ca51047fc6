"""CPU tests of the N>1 layout (spacedust_b200/shard.py): set partitioning, and -- with world_size-2 gloo processes --
that sharding the path by query set and exchanging the per-rank result lists with one all_gather reproduces the
unsharded result exactly.  The per-rank "path" here is the ORACLE (checker) applied to the rank's query sets; the GPU
counterpart is tests/test_shard_gpu.py / bench.py --gpus N."""
import os
import socket
import sys

import numpy as np
import pytest

import helpers
import spacedust_b200 as sd
from spacedust_b200 import shard, synth


def test_partition_sets_covers_and_balances():
    rng = np.random.default_rng(1)
    for n_sets, world in ((20, 1), (20, 2), (20, 8), (200, 8), (3, 8), (1, 2), (0, 2)):
        w = rng.integers(500_000, 1_500_000, size=n_sets)
        blocks = shard.partition_sets(w, world)
        assert len(blocks) == world
        assert blocks[0][0] == 0 and blocks[-1][1] == n_sets
        for (a, b), (c, d) in zip(blocks, blocks[1:]):
            assert a <= b == c <= d
        if n_sets >= 4 * world:
            loads = [w[a:b].sum() for a, b in blocks]
            assert max(loads) <= 1.35 * (w.sum() / world), (n_sets, world, loads)
    assert shard.partition_sets([5, 5, 5, 5], 2) == [(0, 2), (2, 4)]
    # fewer sets than ranks: the first ranks get one set each, rank 0 is never empty
    assert shard.partition_sets([100], 2) == [(0, 1), (1, 1)]
    assert shard.partition_sets([1000, 1, 1], 3) == [(0, 1), (1, 2), (2, 3)]
    assert shard.partition_sets([7, 7, 7], 8)[:4] == [(0, 1), (1, 2), (2, 3), (3, 3)]
    for n_sets, world in ((20, 8), (9, 8), (8, 8)):
        assert all(b > a for a, b in shard.partition_sets(np.ones(n_sets), world))


def test_query_block():
    set_of = np.array([0, 0, 1, 1, 1, 2, 3, 3])
    assert shard.query_block(set_of, 1, 3) == (2, 6)
    assert shard.query_block(set_of, 0, 4) == (0, 8)
    assert shard.query_block(set_of, 4, 4) == (0, 0)
    with pytest.raises(ValueError):
        shard.query_block(np.array([0, 1, 0]), 0, 1)


def _records(o, q_lo, q_hi):
    """Oracle output restricted to query genes [q_lo, q_hi) in the product's record layout (query indices block-relative)."""
    alns_o, bt = o["alns"], o["bt"]
    sel = np.nonzero((alns_o["query"] >= q_lo) & (alns_o["query"] < q_hi))[0]
    a = np.zeros(len(sel), dtype=sd.ALN_DTYPE)
    for fo, fg in (("query", "query"), ("target", "target"), ("score", "score"), ("bits", "bits"), ("seqId", "seq_id"),
                   ("qcov", "qcov"), ("tcov", "tcov"), ("evalue", "evalue"), ("qStart", "q_start"), ("qEnd", "q_end"),
                   ("qLen", "q_len"), ("tStart", "t_start"), ("tEnd", "t_end"), ("tLen", "t_len"), ("alnLen", "aln_len"),
                   ("btLen", "bt_len")):
        a[fg] = alns_o[fo][sel]
    a["query"] -= np.uint32(q_lo)
    pool = []
    p = 0
    for i, j in enumerate(sel):
        b = bt[int(alns_o["btOff"][j]):int(alns_o["btOff"][j]) + int(alns_o["btLen"][j])]
        a["bt_off"][i] = p
        pool.append(b)
        p += len(b)
    pool = np.concatenate(pool) if pool else np.zeros(0, dtype=np.uint8)
    where = {(int(alns_o["query"][j]), int(alns_o["target"][j])): i for i, j in enumerate(sel)}
    clusters, members = [], []
    for hdr, lines in o["clusters"]:
        h = hdr.split("\t")
        ms = []
        for l in lines:
            f = l.split("\t")
            if (int(f[0]), int(f[1])) not in where:
                ms = None
                break
            ms.append((where[(int(f[0]), int(f[1]))], 0, float(f[2])))
        if ms is None:
            continue
        clusters.append((int(h[0]), int(h[1]), float(h[2]), float(h[3]), len(ms), len(members)))
        members.extend(ms)
    return {"alns": a, "pool": pool, "clusters": np.array(clusters, dtype=sd.CLUSTER_DTYPE).reshape(-1),
            "members": np.array(members, dtype=sd.MEMBER_DTYPE).reshape(-1)}


def _worker(rank, world, port, tmpdir):
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(helpers.ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        prot = synth.generate(3, 120, seed=7)
        z = np.load(os.path.join(tmpdir, "alns.npz"))
        alns, bt = z["alns"], z["bt"]
        buf, off, _ = synth.pack(prot)
        meta = synth.set_metadata(prot)
        set_res = np.array([sum(len(s) for s in p["seqs"]) for p in prot])
        lo, hi = shard.partition_sets(set_res, world)[rank]
        q_lo, q_hi = shard.query_block(meta["set_of"], lo, hi)
        # the rank's share of the path: all target sets, only its query sets
        o = helpers.oracle_clustersearch(prot, filter_self_match=True, alns_override=(alns, bt), query_sets=range(lo, hi))
        local = _records(o, q_lo, q_hi)
        merged = shard.allgather_results(shard.rebase_result(local, q_lo))
        np.savez(os.path.join(tmpdir, "merged_%d.npz" % rank), **merged)
        # the persistent-buffer exchange bench.py uses: result written into the send block, parts come back as views
        ex = shard.ResultExchange()
        for _ in range(2):          # second round reuses the buffers
            views = ex.carve(len(local["alns"]), len(local["pool"]), len(local["clusters"]), len(local["members"]))
            for v, k in zip(views, ("alns", "pool", "clusters", "members")):
                v[...] = local[k]
            merged2 = shard.merge_results(ex.exchange(q_lo))
        np.savez(os.path.join(tmpdir, "merged2_%d.npz" % rank), **merged2)
    finally:
        dist.destroy_process_group()


def test_sharded_equals_unsharded_gloo(tmp_path, synth_small):
    import torch.multiprocessing as mp
    o = helpers.oracle_clustersearch(synth_small, filter_self_match=True)
    np.savez(os.path.join(tmp_path, "alns.npz"), alns=o["alns"], bt=o["bt"])
    whole = _records(o, 0, len(synth.set_metadata(synth_small)["set_of"]))
    assert len(whole["clusters"]) == len(o["clusters"]) > 5
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    world = 2
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for rank, name in [(r, n) for r in range(world) for n in ("merged", "merged2")]:
        m = np.load(os.path.join(tmp_path, "%s_%d.npz" % (name, rank)))
        for k in ("alns", "clusters", "members"):
            assert m[k].dtype == whole[k].dtype and len(m[k]) == len(whole[k]), k
            for f in whole[k].dtype.names:
                assert np.array_equal(m[k][f], whole[k][f]), (k, f)
        assert np.array_equal(m["pool"], whole["pool"])


def _worker_empty(rank, world, port, tmpdir):
    """Rank 1 owns no query set at all (fewer sets than ranks): its block is empty, the exchange must still work."""
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ex = shard.ResultExchange()
        n = 5 if rank == 0 else 0
        alns, pool, clusters, members = ex.carve(n, 3 * n, n // 2, n)
        alns["query"] = np.arange(n)
        alns["bt_off"] = 3 * np.arange(n)
        alns["bt_len"] = 3
        pool[...] = 77
        clusters["first"] = 2 * np.arange(n // 2)
        clusters["n"] = 2
        members["aln"] = np.arange(n)
        parts = ex.exchange(10 if rank == 0 else 0)
        merged = shard.merge_results(parts)
        assert [len(p["alns"]) for p in parts] == [5, 0]
        assert list(merged["alns"]["query"]) == [10, 11, 12, 13, 14] and len(merged["pool"]) == 15
        assert list(merged["clusters"]["first"]) == [0, 2] and list(merged["members"]["aln"]) == [0, 1, 2, 3, 4]
        np.save(os.path.join(tmpdir, "ok_%d.npy" % rank), np.array([1]))
    finally:
        dist.destroy_process_group()


def test_exchange_with_an_empty_rank(tmp_path):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker_empty, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert all(os.path.exists(os.path.join(tmp_path, "ok_%d.npy" % r)) for r in range(2))


def test_c_abi_partition_equals_python():
    """sd_shard_partition (csrc/shard.cu, host code) == shard.partition_sets, including fewer sets than ranks and skewed weights."""
    import spacedust_b200 as sd
    from spacedust_b200 import shard
    rng = np.random.default_rng(4)
    for n, world in [(1, 1), (1, 4), (3, 8), (8, 8), (20, 8), (20, 3), (200, 8), (7, 2), (500, 8)]:
        for _ in range(5):
            w = rng.integers(1, 10 ** 6, size=n).astype(np.uint64)
            if rng.random() < 0.3:
                w[rng.integers(0, n)] *= 50
            assert sd.shard_partition(w, world) == shard.partition_sets(w, world), (n, world, w[:8])


def test_c_abi_partition_genes():
    """sd_shard_partition_genes (csrc/shard.cu, host code): contiguous blocks that cover all genes, in order, balanced by residues to
    within the longest gene; more ranks than genes leaves empty blocks."""
    import spacedust_b200 as sd
    rng = np.random.default_rng(5)
    for n, world in [(1, 1), (1, 4), (3, 8), (60, 8), (1000, 3), (60000, 8)]:
        lens = rng.integers(30, 2500, size=n).astype(np.uint64)
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64) + 7      # offsets need not start at 0
        blocks = sd.shard_partition_genes(off, world)
        assert len(blocks) == world and blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1)) and all(lo <= hi for lo, hi in blocks)
        share = float(lens.sum()) / world
        for lo, hi in blocks:
            assert float(lens[lo:hi].sum()) <= share + 2 * float(lens.max())
