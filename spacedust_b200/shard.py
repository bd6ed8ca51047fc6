"""Multi-GPU layout of the clustersearch hot path: one process per GPU, the target index replicated, QUERY SETS sharded.

Every stage of the path is independent per query set once the whole target side is visible: a query's prefilter hit
list depends on the query and the full target index (the `--max-seqs` cut and its tie order are global over targets,
lib/mmseqs/src/prefiltering/QueryMatcher.cpp:108-211), E-values use the total target residues (Alignment.cpp:263),
combinehits multiplies by the total number of target sets (src/util/combinehits.cpp:77,153), and besthitbyset /
combinehits / clusterhits work per (query set, target set).  So a rank owns a contiguous block of query sets balanced
by residues (the analogue of DBReader::decomposeDomainByAminoAcid), runs the path for it with no data-path collective,
and the per-rank result lists are exchanged with ONE all_gather (sizes, then padded payload) -- NCCL over NVLink on
the GPU box, gloo in the CPU tests.
"""
import numpy as np


def partition_sets(set_residues, world):
    """Contiguous blocks of sets with ~equal residues: -> [(lo, hi)] * world (hi exclusive; blocks may be empty
    when there are fewer sets than ranks).  Greedy split at the prefix-sum quantiles, every set in exactly one block."""
    w = np.asarray(set_residues, dtype=np.float64)
    n = len(w)
    if world <= 0:
        raise ValueError("world must be positive")
    pre = np.concatenate([[0.0], np.cumsum(w)])
    total = pre[-1]
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        # first boundary whose prefix is >= target, but never before the previous cut
        c = int(np.searchsorted(pre, target, side="left"))
        # choose the closer of c-1 / c
        if c > 0 and c <= n and abs(pre[c - 1] - target) <= abs(pre[min(c, n)] - target):
            c -= 1
        c = max(cuts[-1], min(c, n))
        cuts.append(c)
    cuts.append(n)
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


def query_block(set_of, lo, hi):
    """Genes (file order, sets contiguous as createsetdb writes them) of sets [lo, hi) -> (first, last+1)."""
    set_of = np.asarray(set_of)
    sel = np.nonzero((set_of >= lo) & (set_of < hi))[0]
    if len(sel) == 0:
        return 0, 0
    first, last = int(sel[0]), int(sel[-1]) + 1
    if last - first != len(sel):
        raise ValueError("sets are not contiguous in the DB order")
    return first, last


def rebase_result(res, query_first):
    """Local result of a rank (query indices relative to its block) -> global query indices."""
    out = dict(res)
    alns = res["alns"].copy()
    alns["query"] += np.uint32(query_first)
    out["alns"] = alns
    return out


def merge_results(parts):
    """Concatenate per-rank results in rank order, rebasing the cross references (cluster.first -> member array,
    member.aln -> alignment array, alignment.bt_off -> backtrace pool).  Ranks own ascending query-set blocks, so
    the merged alignment list is grouped by ascending query and clusters by ascending (query set, target set), i.e. the
    order a single-GPU run returns."""
    alns, pools, clusters, members = [], [], [], []
    a0 = p0 = m0 = 0
    for r in parts:
        a = r["alns"].copy()
        a["bt_off"] += np.uint64(p0)
        m = r["members"].copy()
        m["aln"] += np.uint32(a0)
        c = r["clusters"].copy()
        c["first"] += np.uint32(m0)
        alns.append(a); pools.append(r["pool"]); clusters.append(c); members.append(m)
        a0 += len(a); p0 += len(r["pool"]); m0 += len(m)
    cat = lambda xs, dt: np.concatenate(xs) if xs else np.zeros(0, dtype=dt)      # noqa: E731
    return {"alns": cat(alns, None), "pool": cat(pools, np.uint8), "clusters": cat(clusters, None), "members": cat(members, None)}


def allgather_bytes(payload, group=None, device=None):
    """ONE variable-size all_gather of a byte string per rank -> list of np.uint8 arrays in rank order.
    `device`: torch device of the exchange buffers (cuda for NCCL, None/cpu for gloo)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    buf = np.ascontiguousarray(payload, dtype=np.uint8)
    n = torch.tensor([buf.size], dtype=torch.int64, device=device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    cap = max(max(sizes), 1)
    send = torch.zeros(cap, dtype=torch.uint8, device=device)
    if buf.size:
        send[:buf.size].copy_(torch.from_numpy(buf), non_blocking=False)
    recv = [torch.empty(cap, dtype=torch.uint8, device=device) for _ in range(world)]
    dist.all_gather(recv, send, group=group)
    return [r[:s].cpu().numpy() for r, s in zip(recv, sizes)]


_FIELDS = ("alns", "pool", "clusters", "members")


def allgather_results(res, group=None, device=None):
    """Exchange the per-rank result lists (alignments, backtrace pool, clusters, members) with one sized all_gather and
    merge them (merge_results).  Every rank ends with the whole job's result, as the reference's single process does."""
    dtypes = {k: res[k].dtype for k in _FIELDS}
    head = np.array([res[k].nbytes for k in _FIELDS], dtype=np.int64)
    blob = np.concatenate([head.view(np.uint8)] + [np.ascontiguousarray(res[k]).view(np.uint8).reshape(-1) for k in _FIELDS])
    parts = []
    for raw in allgather_bytes(blob, group, device):
        h = raw[:8 * len(_FIELDS)].view(np.int64)
        o = 8 * len(_FIELDS)
        d = {}
        for k, nb in zip(_FIELDS, h):
            d[k] = raw[o:o + int(nb)].view(dtypes[k]).copy()
            o += int(nb)
        parts.append(d)
    return merge_results(parts)
