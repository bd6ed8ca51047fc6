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
    """Contiguous blocks of sets with ~equal residues: -> [(lo, hi)] * world (hi exclusive).  Split at the prefix-sum
    quantiles; every set in exactly one block; no rank is left empty while another holds two sets (with fewer sets than
    ranks the LAST ranks are empty, so rank 0 always has work)."""
    w = np.asarray(set_residues, dtype=np.float64)
    n = len(w)
    if world <= 0:
        raise ValueError("world must be positive")
    pre = np.concatenate([[0.0], np.cumsum(w)])
    total = pre[-1]
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        c = int(np.searchsorted(pre, target, side="left"))
        # the closer of the two neighbouring boundaries
        if 0 < c <= n and abs(pre[c - 1] - target) < abs(pre[c] - target):
            c -= 1
        c = min(c, n)
        lo_bound = min(cuts[-1] + 1, n)               # the previous rank gets at least one set while sets remain
        hi_bound = max(n - (world - r), lo_bound)     # ... and so does every later rank, when there are enough sets
        c = max(lo_bound, min(c, hi_bound))
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


class ResultExchange:
    """Persistent buffers for the result lists of one rank and for the one all_gather of the N>1 path.

    carve() hands sd.clustersearch four array views into ONE pinned host block laid out as
        [int64 n_alns, pool_bytes, n_clusters, n_members | alignments | backtrace pool | clusters | members]
    (fields 16-byte aligned), so the C ABI writes the result straight into the send buffer.  exchange() rebases the
    rank's query indices to global ones in place, all_gathers the block sizes, then moves every rank's block with one
    all_gather_into_tensor (padded to the largest block) -- pinned host -> HBM -> NCCL over NVLink -> pinned host on
    the GPU box, plain host tensors with gloo -- and returns the per-rank parts as views (merge_results() makes them
    contiguous when a consumer needs that)."""

    HEAD = 64

    def __init__(self, device=None, group=None):
        import torch
        self.torch = torch
        self.device = device
        self.group = group
        self.cuda = device is not None and torch.device(device).type == "cuda"
        self.send_h = None        # pinned (cuda) or plain host block of this rank
        self.recv_h = None
        self.send_d = self.recv_d = None
        self.used = 0
        self.sizes = (0, 0, 0, 0)

    def _host(self, nbytes):
        t = self.torch.empty(nbytes, dtype=self.torch.uint8)
        return t.pin_memory() if self.cuda else t

    @staticmethod
    def _layout(na, nb, nc, nm):
        import spacedust_b200 as sd
        al = lambda x: (x + 15) & ~15      # noqa: E731
        o_a = ResultExchange.HEAD
        o_p = al(o_a + na * sd.ALN_DTYPE.itemsize)
        o_c = al(o_p + nb)
        o_m = al(o_c + nc * sd.CLUSTER_DTYPE.itemsize)
        end = al(o_m + nm * sd.MEMBER_DTYPE.itemsize)
        return o_a, o_p, o_c, o_m, end

    @staticmethod
    def _views(block, na, nb, nc, nm):
        import spacedust_b200 as sd
        o_a, o_p, o_c, o_m, _ = ResultExchange._layout(na, nb, nc, nm)
        return (block[o_a:o_a + na * sd.ALN_DTYPE.itemsize].view(sd.ALN_DTYPE), block[o_p:o_p + nb],
                block[o_c:o_c + nc * sd.CLUSTER_DTYPE.itemsize].view(sd.CLUSTER_DTYPE),
                block[o_m:o_m + nm * sd.MEMBER_DTYPE.itemsize].view(sd.MEMBER_DTYPE))

    def carve(self, na, nb, nc, nm):
        na, nb, nc, nm = int(na), int(nb), int(nc), int(nm)
        end = self._layout(na, nb, nc, nm)[4]
        if self.send_h is None or self.send_h.numel() < end:
            self.send_h = self._host(end + end // 4 + 4096)
        block = self.send_h.numpy()
        block[:self.HEAD].view(np.int64)[:4] = (na, nb, nc, nm)
        self.used = end
        self.sizes = (na, nb, nc, nm)
        return self._views(block, na, nb, nc, nm)

    def local(self):
        a, p, c, m = self._views(self.send_h.numpy(), *self.sizes)
        return {"alns": a, "pool": p, "clusters": c, "members": m}

    def exchange(self, query_first):
        """-> list (rank order) of {"alns","pool","clusters","members"} views with global query indices."""
        torch = self.torch
        import torch.distributed as dist
        world = dist.get_world_size(self.group)
        self.local()["alns"]["query"] += np.uint32(query_first)
        n = torch.tensor([self.used], dtype=torch.int64, device=self.device)
        sizes = torch.empty(world, dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(sizes, n, group=self.group)
        cap = (int(sizes.max().item()) + 4095) & ~4095
        if self.recv_h is None or self.recv_h.numel() < cap * world:
            self.recv_h = self._host(cap * world + cap // 4 * world)
            if self.cuda:
                self.recv_d = torch.empty(self.recv_h.numel(), dtype=torch.uint8, device=self.device)
        if self.send_h.numel() < cap:          # the padded send must be readable up to cap
            grown = self._host(cap + cap // 4)
            grown[:self.used].copy_(self.send_h[:self.used])
            self.send_h = grown
        if self.cuda:
            if self.send_d is None or self.send_d.numel() < cap:
                self.send_d = torch.empty(cap + cap // 4, dtype=torch.uint8, device=self.device)
            self.send_d[:self.used].copy_(self.send_h[:self.used], non_blocking=True)
            dist.all_gather_into_tensor(self.recv_d[:cap * world], self.send_d[:cap], group=self.group)
            self.recv_h[:cap * world].copy_(self.recv_d[:cap * world], non_blocking=True)
            torch.cuda.current_stream().synchronize()
        else:
            dist.all_gather_into_tensor(self.recv_h[:cap * world], self.send_h[:cap], group=self.group)
        raw = self.recv_h.numpy()
        parts = []
        for r in range(world):
            block = raw[r * cap:(r + 1) * cap]
            na, nb, nc, nm = (int(x) for x in block[:self.HEAD].view(np.int64)[:4])
            a, p, c, m = self._views(block, na, nb, nc, nm)
            parts.append({"alns": a, "pool": p, "clusters": c, "members": m})
        return parts
