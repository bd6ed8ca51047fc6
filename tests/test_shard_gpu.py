"""GPU, 2 ranks (skipped on a 1-GPU box): one process per GPU over NCCL; query sets sharded with one all-gather of the result
lists, and query genes sharded with the two exchanges of sd_shard_clustersearch_run -> both identical to the single-GPU result."""
import os
import socket
import sys

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, tmpdir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(helpers.ROOT, "tests"))
    import spacedust_b200 as sd
    from spacedust_b200 import shard, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        prot = synth.generate(4, 150, seed=31)
        buf, off, _ = synth.pack(prot)
        meta = synth.set_metadata(prot)
        set_res = np.array([sum(len(s) for s in p["seqs"]) for p in prot])
        lo, hi = shard.partition_sets(set_res, world)[rank]
        q_lo, q_hi = shard.query_block(meta["set_of"], lo, hi)
        qbuf = buf[int(off[q_lo]):int(off[q_hi])]
        qoff = (off[q_lo:q_hi + 1] - off[q_lo]).astype(np.uint64)
        q_meta = {k: (meta[k][q_lo:q_hi] if k != "set_size" else meta[k]) for k in ("set_of", "pos", "strand", "set_size")}
        ctx = sd.Context(rank)
        pp = sd.default_prefilter_params(same_db=1, bin_size=2)
        idx = sd.TargetIndex(ctx, buf, off, pp)
        batch = sd.QueryBatch(ctx, qbuf, qoff)
        r = sd.clustersearch(ctx, idx, batch, q_meta, meta, pp, query_first=q_lo)
        merged = shard.allgather_results(shard.rebase_result(r, q_lo), device=torch.device("cuda", rank))
        ex = shard.ResultExchange(device=torch.device("cuda", rank))
        for _ in range(2):
            sd.clustersearch(ctx, idx, batch, q_meta, meta, pp, query_first=q_lo, out=ex)
            merged2 = shard.merge_results(ex.exchange(q_lo))
        for k in ("alns", "clusters", "members"):
            for f in merged[k].dtype.names:
                if True:
                    assert np.array_equal(merged[k][f], merged2[k][f]), (k, f)
        assert np.array_equal(merged["pool"], merged2["pool"])
        # the product path: communicator and all-gather inside libsdgpu.so (sd_shard_init / sd_shard_allgather_result)
        box = [sd.shard_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        sd.shard_init(ctx, rank, world, box[0])
        assert sd.shard_partition(set_res, world)[rank] == (lo, hi)
        for _ in range(2):
            r3 = sd.clustersearch(ctx, idx, batch, q_meta, meta, pp, query_first=q_lo, gather=True)
        assert r3["bytes_gathered"] > 0
        for k in ("alns", "clusters", "members"):
            assert len(r3[k]) == len(merged[k]), k
            for f in merged[k].dtype.names:
                assert np.array_equal(r3[k][f], merged[k][f]), (k, f)
        assert np.array_equal(r3["pool"], merged["pool"])
        np.savez(os.path.join(tmpdir, "merged_%d.npz" % rank), **merged)
        # gene-granular sharding (what bench.py and `sdb200 --gpus N` use): this rank's genes for prefilter + align, both exchanges
        # and the set-wise cluster stages inside the library (sd_shard_clustersearch_run)
        g_lo, g_hi = sd.shard_partition_genes(off, world)[rank]
        gbatch = sd.QueryBatch(ctx, buf[int(off[g_lo]):int(off[g_hi])], (off[g_lo:g_hi + 1] - off[g_lo]).astype(np.uint64))
        for _ in range(2):
            r4 = sd.clustersearch(ctx, idx, gbatch, meta, meta, pp, query_first=g_lo, shard_genes=True)
        assert r4["bytes_gathered"] > 0
        np.savez(os.path.join(tmpdir, "genes_%d.npz" % rank), **{k: r4[k] for k in ("alns", "pool", "clusters", "members")})
        gbatch.close()
        if rank == 0:      # the unsharded run on one GPU
            whole = sd.clustersearch(ctx, idx, sd.QueryBatch(ctx, buf, off), meta, meta, pp)
            np.savez(os.path.join(tmpdir, "whole.npz"), **{k: whole[k] for k in ("alns", "pool", "clusters", "members")})
        batch.close(); idx.close(); ctx.close()
    finally:
        dist.destroy_process_group()


def test_two_ranks_equal_one(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    whole = np.load(os.path.join(tmp_path, "whole.npz"))
    assert len(whole["clusters"]) > 10
    for name in ("merged_0", "merged_1", "genes_0", "genes_1"):
        m = np.load(os.path.join(tmp_path, name + ".npz"))
        for k in ("alns", "clusters", "members"):
            assert len(m[k]) == len(whole[k]), k
            for f in whole[k].dtype.names:
                if True:
                    assert np.array_equal(m[k][f], whole[k][f]), (k, f)
        assert np.array_equal(m["pool"], whole["pool"])


def test_frontend_two_gpus_equal_one(tmp_path):
    """`sdb200 clustersearch --gpus 2` (process per GPU started by the front-end, NCCL id through the tmp dir, one all-gather in
    libsdgpu.so, rank 0 writes) gives the TSV of the single-GPU run."""
    import subprocess

    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import spacedust_b200 as sd
    from spacedust_b200 import synth
    sdb = os.path.join(os.path.dirname(sd.LIB_PATH), "sdb200")
    prot = synth.generate(5, 120, seed=17)
    paths = synth.write_fasta(prot, os.path.join(tmp_path, "faa"))
    db = os.path.join(tmp_path, "setDB")
    subprocess.run([sdb, "createsetdb"] + paths + [db, os.path.join(tmp_path, "tmp")], check=True)
    outs = []
    for n in (1, 2):
        out = os.path.join(tmp_path, "res%d.tsv" % n)
        subprocess.run([sdb, "clustersearch", db, db, out, os.path.join(tmp_path, "tmp%d" % n), "--filter-self-match", "1", "--gpus", str(n), "-v", "1"],
                       check=True)
        outs.append(open(out).read())
    assert outs[0].count("#") > 5
    assert outs[0] == outs[1]
