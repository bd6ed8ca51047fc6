#!/usr/bin/env python3
"""bench.py -- `spacedust clustersearch --search-mode 0` hot path on B200 vs the reference's AVX2 CPU path.

Metric (BASELINE.json): query-genes/s of the clustersearch hot path (prefilter -> align -> besthitbyset/combinehits ->
clusterhits) on synthetic proteomes.  Workload at N=1 (default, --workload c2) is BASELINE config[1]: 20 synthetic 3k-gene
prokaryote proteomes all-vs-all (60 000 query genes, 19.5 M residues), generator spacedust_b200/synth.py, seed 20261017.
--workload c3 = BASELINE config[2] (200 x 4000), --workload c5 = config[4] (500 x 4000, --max-gene-gap 10 --cluster-size 5);
--proteomes / --genes give any other size (labelled "custom").

One step = one pass of the hot path over the whole query batch.
  value  : target DB + k-mer index and the query batch already resident in HBM when the timed region starts
  e2e    : same step through the C ABI from pinned HOST buffers: target index build (upload, masking, index),
           query upload, all stages, results read back -- what `clustersearch` does after `createsetdb`
Multi-GPU (--gpus N via torchrun): the target index is replicated, query GENES are split across ranks in contiguous
residue-balanced blocks for prefilter + align, an ncclAllGather inside libsdgpu.so merges the alignment lists, every rank
clusters its block of query SETS, a second ncclAllGather merges the cluster lists, so that every rank ends with the whole job's
result; total work fixed => "strong".

`--impl reference` times the unmodified reference binary (oracle/_ref/spacedust, built by oracle/build_ref.py)
on the host cores on a bounded sample of the same workload: S query proteomes against the full target setDB, S chosen so
that the whole --steps/--warmup run stays within a few minutes (the line states S and an estimate for the full run, in which
the reference's fixed start-up -- index table, extended 3-mer matrix -- is amortised over all queries instead of S proteomes).
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_PROTEOMES = 20
GENES = 3000
SEED = 20261017
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "spacedust")
# BASELINE.json configs that bench.py can run ("cluster": flags of clusterhits that differ from the defaults)
WORKLOADS = {
    "c2": {"proteomes": N_PROTEOMES, "genes": GENES, "cluster": {}, "label": "BASELINE config[1]"},
    "c3": {"proteomes": 200, "genes": 4000, "cluster": {}, "label": "BASELINE config[2]"},
    "c5": {"proteomes": 500, "genes": 4000, "cluster": {"max_gene_gap": 10, "cluster_size": 5}, "label": "BASELINE config[4]"},
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS), help="BASELINE.json config (default c2 = config[1])")
    ap.add_argument("--proteomes", type=int, default=0, help="override the workload's proteome count")
    ap.add_argument("--genes", type=int, default=0, help="override the workload's genes per proteome")
    ap.add_argument("--ref-sample", type=int, default=0,
                    help="query proteomes in the reference / cpu_baseline sample (0: as many as fit the time budget)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=-1, help="timed passes of the end-to-end arm (default: --steps)")
    args = ap.parse_args()
    if args.e2e_steps < 0:
        args.e2e_steps = args.steps
    w = WORKLOADS[args.workload]
    args.custom = bool(args.proteomes or args.genes)
    args.proteomes = args.proteomes or w["proteomes"]
    args.genes = args.genes or w["genes"]
    args.cluster_flags = dict(w["cluster"])
    return args


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons every 200 ms while the timed region runs."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop_evt = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, timeout=5).stdout.decode().strip()
                f = [x.strip() for x in out.split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for n, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=5)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ reference arm
def parse_stage_times(log):
    """`Time for processing: 0h 0m 5s 798ms` after each module banner (lib/mmseqs/src/commons/Application.cpp:45-48)."""
    stages, cur = {}, None
    for line in log.split("\n"):
        w = line.split(" ")
        if w and w[0] in ("prefilter", "align", "prefixid", "besthitbyset", "mergeresultsbyset", "combinehits", "clusterhits",
                          "summarizeresults"):
            cur = w[0]
        if line.startswith("Time for processing:") and cur:
            f = line.replace("Time for processing:", "").split()
            sec = 0.0
            for tok in f:
                if tok.endswith("ms"):
                    sec += float(tok[:-2]) / 1e3
                elif tok.endswith("h"):
                    sec += float(tok[:-1]) * 3600
                elif tok.endswith("m"):
                    sec += float(tok[:-1]) * 60
                elif tok.endswith("s"):
                    sec += float(tok[:-1])
            stages[cur] = stages.get(cur, 0.0) + sec
            cur = None
        if line.startswith("Time for index table init:"):
            f = line.replace("Time for index table init:", "").split()
            sec = 0.0
            for tok in f:
                if tok.endswith("ms"):
                    sec += float(tok[:-2]) / 1e3
                elif tok.endswith("h"):
                    sec += float(tok[:-1]) * 3600
                elif tok.endswith("m"):
                    sec += float(tok[:-1]) * 60
                elif tok.endswith("s"):
                    sec += float(tok[:-1])
            stages["index_table_init"] = sec
    return stages


def pick_ref_sample(args, runs, budget_s):
    """Query proteomes per reference run such that `runs` runs stay within budget_s.  Cost model from the reference's own stage
    times on this pool's hosts (16 cores): ~1.5 s fixed start-up (index table + 3-mer matrix) and ~0.75 ms per query gene and
    per 60 k target genes; scaled by the core count."""
    if args.ref_sample > 0:
        return min(args.ref_sample, args.proteomes)
    cores = os.cpu_count() or 1
    per_proteome = 0.75e-3 * args.genes * (args.proteomes * args.genes / 60000.0) * (16.0 / cores)
    fixed = 1.5 * (args.proteomes * args.genes / 60000.0)
    s = int((budget_s / max(runs, 1) - fixed) / max(per_proteome, 1e-9))
    return max(1, min(args.proteomes, max(s, 2 if args.proteomes >= 2 else 1)))


class ReferenceRunner:
    """createsetdb once (excluded from the timing), then clustersearch per step."""

    def __init__(self, args, proteomes, n_query_sets):
        from spacedust_b200 import synth
        if not (os.path.exists(REF_BIN) and os.access(REF_BIN, os.X_OK)):
            raise RuntimeError("oracle/_ref/spacedust is missing: run `python oracle/build_ref.py` in the build container")
        self.dir = tempfile.mkdtemp(prefix="sdbench_ref_")
        self.threads = os.cpu_count() or 1
        self.env = dict(os.environ, MMSEQS_FORCE_MERGE="1")
        self.flags = []
        for k, v in args.cluster_flags.items():
            self.flags += ["--" + k.replace("_", "-"), str(v)]
        self.keep = False
        paths = synth.write_fasta(proteomes, os.path.join(self.dir, "faa"))
        self.n_queries = sum(len(p["seqs"]) for p in proteomes[:n_query_sets])
        n_targets = sum(len(p["seqs"]) for p in proteomes)
        for attempt in range(2):      # createsetdb runs awk/sort/join children: validate what they wrote, rebuild once if not
            for db in ("tDB", "qDB"):
                for f in os.listdir(self.dir):
                    if f.startswith(db):
                        os.remove(os.path.join(self.dir, f))
            self._run(["createsetdb"] + paths + [os.path.join(self.dir, "tDB"), os.path.join(self.dir, "tmp_t%d" % attempt), "-v", "1"])
            self._run(["createsetdb"] + paths[:n_query_sets] + [os.path.join(self.dir, "qDB"), os.path.join(self.dir, "tmp_q%d" % attempt), "-v", "1"])
            bad = self._check_lookup("qDB", self.n_queries) or self._check_lookup("tDB", n_targets)
            if not bad:
                break
            sys.stderr.write("[bench] reference createsetdb wrote an invalid lookup (%s), attempt %d\n" % (bad, attempt + 1))
            if attempt == 1:
                self.keep = True
                raise RuntimeError("reference createsetdb: %s (files kept in %s)" % (bad, self.dir))
        self.step_no = 0

    def _check_lookup(self, db, n_expected):
        """<db>.lookup: key \t accession_geneIdx_start_end \t set; clusterhits parses the last three '_' fields
        (src/util/ClusterHits.cpp:328-350) and dies with "Invalid query lookup record" otherwise."""
        n = 0
        try:
            for line in open(os.path.join(self.dir, db + ".lookup")):
                f = line.rstrip("\n").split("\t")
                parts = f[1].split("_") if len(f) == 3 else []
                if len(parts) < 4 or not (parts[-1].isdigit() and parts[-2].isdigit() and parts[-3].isdigit()):
                    return "%s.lookup line %d: %r" % (db, n + 1, line[:80])
                n += 1
        except OSError as e:
            return "%s.lookup: %s" % (db, e)
        return None if n == n_expected else "%s.lookup has %d records, expected %d" % (db, n, n_expected)

    def _run(self, args):
        r = subprocess.run([REF_BIN] + args, env=self.env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        if r.returncode != 0:
            raise RuntimeError("reference failed (rc %d): %s\n%s" % (r.returncode, " ".join(args[:3]), r.stdout.decode(errors="replace")[-1500:]))
        return r.stdout.decode(errors="replace")

    def step(self):
        """One clustersearch run; a failed run is repeated once (the reference's clusterhits has died sporadically on these hosts),
        the failing tmp dir and the lookup files are kept for inspection."""
        last = None
        for attempt in range(2):
            self.step_no += 1
            tmp = os.path.join(self.dir, "tmp_cs%d" % self.step_no)
            out = os.path.join(self.dir, "out%d.tsv" % self.step_no)
            t0 = time.perf_counter()
            try:
                log = self._run(["clustersearch", os.path.join(self.dir, "qDB"), os.path.join(self.dir, "tDB"), out, tmp,
                                 "--filter-self-match", "1", "--threads", str(self.threads), "-v", "3"] + self.flags)
            except RuntimeError as e:
                last = e
                self.keep = True
                keep_dir = os.path.join(ROOT, "gpurun_out") if os.path.isdir(os.path.join(ROOT, "gpurun_out")) else self.dir
                try:
                    with open(os.path.join(keep_dir, "ref_failure_step%d.log" % self.step_no), "w") as f:
                        f.write(str(e))
                    for db in ("qDB", "tDB"):
                        shutil.copy(os.path.join(self.dir, db + ".lookup"), os.path.join(keep_dir, "ref_failure_%s.lookup" % db))
                except OSError:
                    pass
                sys.stderr.write("[bench] reference clustersearch failed (attempt %d), evidence in %s and %s: %s\n"
                                 % (attempt + 1, keep_dir, tmp, str(e)[:300]))
                continue
            dt = time.perf_counter() - t0
            self.last_tsv = out
            n_clusters = sum(1 for l in open(out) if l.startswith("#"))
            shutil.rmtree(tmp, ignore_errors=True)
            return dt, parse_stage_times(log), n_clusters
        raise last

    def close(self):
        if not self.keep:
            shutil.rmtree(self.dir, ignore_errors=True)


def full_run_estimate(args, n_sample_queries, dt, stages):
    """The sample runs S query proteomes against the whole target setDB, so the reference's start-up that does not depend on the
    number of queries (index table init, and the extended 3-mer matrix inside the prefilter) is amortised over S instead of all
    proteomes.  Estimate for the full run: fixed part once + the per-query part scaled up."""
    fixed = stages.get("index_table_init", 0.0) + 0.6      # 0.6 s: 3-mer matrix + DB open, measured on these hosts
    fixed = min(fixed, 0.8 * dt)
    n_all = args.proteomes * args.genes
    t_full = fixed + (dt - fixed) * n_all / max(n_sample_queries, 1)
    return {"query_genes_per_s": n_all / t_full, "seconds": t_full, "fixed_seconds_in_sample": fixed,
            "note": "extrapolated from the sample; the sample's own value overstates the GPU/CPU ratio by value_full/value_sample"}


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def run_reference(args, proteomes, rank, world):
    if rank != 0:
        return
    n_sets = pick_ref_sample(args, args.steps + args.warmup, 200.0)
    runner = ReferenceRunner(args, proteomes, n_sets)
    try:
        for _ in range(args.warmup):
            runner.step()
        times, stages = [], {}
        for _ in range(args.steps):
            dt, stages, n_clu = runner.step()
            times.append(dt)
        total = sum(times)
        value = runner.n_queries * args.steps / total
        sample = ("%d of %d query proteomes (%d query genes) against the full %d-proteome target setDB, "
                  "clustersearch --filter-self-match 1 --threads %d%s; createsetdb excluded"
                  % (n_sets, args.proteomes, runner.n_queries, args.proteomes, runner.threads,
                     "".join(" " + f for f in runner.flags)))
        line = {"impl": "reference", "metric": "clustersearch query-genes/s", "value": value, "unit": "query-genes/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32/f64", "data": "synthetic",
                "config": workload_config(args),
                "cpu_baseline": {"value": value, "unit": "query-genes/s", "cores": runner.threads, "kind": "reference",
                                 "sample": sample, "cpu": cpu_model(), "stage_seconds": stages,
                                 "full_run_estimate": full_run_estimate(args, runner.n_queries, total / args.steps, stages)},
                "e2e": {"value": value, "unit": "query-genes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
    finally:
        runner.close()


def workload_config(args):
    label = "custom size" if args.custom else WORKLOADS[args.workload]["label"]
    extra = "".join(" --%s %s" % (k.replace("_", "-"), v) for k, v in args.cluster_flags.items())
    return {"workload": "%d synthetic %d-gene prokaryote proteomes all-vs-all (%s), seed %d, "
                        "clustersearch --search-mode 0 --filter-self-match 1%s, defaults -s 5.7 --max-seqs 300 -e 10 -c 0.8 --cov-mode 2"
                        % (args.proteomes, args.genes, label, SEED, extra),
            "query_genes": args.proteomes * args.genes, "proteomes": args.proteomes, "genes_per_proteome": args.genes,
            "l2": "inputs larger than L2 (k-mer index + neighbour tables + per-query scratch > 1 GB vs 126 MB L2)"}


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args, proteomes, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import spacedust_b200 as sd
    from spacedust_b200 import shard, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    buf, off, _ = synth.pack(proteomes)
    meta = synth.set_metadata(proteomes)
    n_total = len(off) - 1
    # this rank's query GENES: contiguous residue-balanced block (sd_shard_partition_genes, C ABI); the cluster stages, which need
    # whole query sets, are re-partitioned inside the library after the first exchange (sd_shard_clustersearch_run)
    g_lo, g_hi = sd.shard_partition_genes(off, world)[rank]
    q_first = g_lo
    qbuf = buf[int(off[g_lo]):int(off[g_hi])]
    qoff = (off[g_lo:g_hi + 1] - off[g_lo]).astype(np.uint64)
    q_meta = meta if world > 1 else {k: meta[k] for k in ("set_of", "pos", "strand", "set_size")}
    nq_local = g_hi - g_lo

    # pinned host staging for the end-to-end arm
    t_pin = torch.from_numpy(buf.copy()).pin_memory()
    q_pin = torch.from_numpy(qbuf.copy()).pin_memory()
    tbuf_h, qbuf_h = t_pin.numpy(), q_pin.numpy()

    ctx = sd.Context(local_rank)
    if world > 1:
        # the library's own NCCL communicator (csrc/shard.cu); torch.distributed only carries the 128-byte id and the timing reductions
        box = [sd.shard_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        sd.shard_init(ctx, rank, world, box[0])
    pp = sd.default_prefilter_params(same_db=1)
    ap = sd.default_align_params(same_db=1)
    cp = sd.default_cluster_params(filter_self_match=1, **args.cluster_flags)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    dev = torch.device("cuda", local_rank)

    exch = shard.ResultExchange(device=dev)      # persistent pinned result block (+ exchange buffers for N>1)

    def step(index, queries):
        """One pass of the hot path over this rank's query block, results into pinned host memory; N>1: + the one
        ncclAllGather of the result lists inside the library (sd_shard_allgather_result: every rank ends with the whole job's
        lists in host memory, global query indices, single-GPU order)."""
        return sd.clustersearch(ctx, index, queries, q_meta, meta, pp, ap, cp, q_first, out=exch, shard_genes=world > 1)

    # ---------------- resident arm
    idx = sd.TargetIndex(ctx, tbuf_h, off, pp)
    batch = sd.QueryBatch(ctx, qbuf_h, qoff)
    res = None
    for _ in range(args.warmup):
        res = step(idx, batch)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = ctx.launches
    t0 = time.perf_counter()
    stage = {"prefilter_ms": 0.0, "prefilter_kernel_ms": 0.0, "align_ms": 0.0, "sw_kernel_ms": 0.0, "traceback_ms": 0.0,
             "cluster_ms": 0.0, "cluster_kernel_ms": 0.0}
    for _ in range(args.steps):
        res = step(idx, batch)
        stage["prefilter_ms"] += res["prefilter_stats"]["total_ms"]
        stage["prefilter_kernel_ms"] += res["prefilter_stats"]["kernel_ms"]
        stage["align_ms"] += res["align_stats"]["total_ms"]
        stage["sw_kernel_ms"] += res["align_stats"]["sw_ms"]
        stage["traceback_ms"] += res["align_stats"]["traceback_ms"]
        stage["cluster_ms"] += res["cluster_stats"]["total_ms"]
        stage["cluster_kernel_ms"] += res["cluster_stats"]["kernel_ms"]
    barrier()
    dt = time.perf_counter() - t0
    launches = ctx.launches - launches0
    clocks = sampler.stop() if sampler else None
    # ---------------- end-to-end arm: host buffers in, results out, every step
    batch.close()
    idx.close()
    e2e_h2d = int(len(tbuf_h) + len(qbuf_h) + 8 * (len(off) + len(qoff)))
    e2e_d2h = 0
    e2e_parts = {"index_build_ms": 0.0, "query_upload_ms": 0.0, "search_ms": 0.0, "search_stage_sum_ms": 0.0, "teardown_ms": 0.0}
    sampler2 = ClockSampler(local_rank) if rank == 0 else None
    for it in range(1 + args.e2e_steps):      # one untimed pass, then K timed
        if it == 1:
            barrier()
            if sampler2:
                sampler2.start()
            t1 = time.perf_counter()
        ta = time.perf_counter()
        idx2 = sd.TargetIndex(ctx, tbuf_h, off, pp)
        tb = time.perf_counter()
        b2 = sd.QueryBatch(ctx, qbuf_h, qoff)
        tc = time.perf_counter()
        r2 = step(idx2, b2)
        td = time.perf_counter()
        e2e_d2h = int(r2["alns"].nbytes + r2["pool"].nbytes + r2["clusters"].nbytes + r2["members"].nbytes)
        b2.close()
        idx2.close()
        te = time.perf_counter()
        if it >= 1:     # host wall clock of the (synchronous) C-ABI calls
            e2e_parts["index_build_ms"] += 1e3 * (tb - ta)
            e2e_parts["query_upload_ms"] += 1e3 * (tc - tb)
            e2e_parts["search_ms"] += 1e3 * (td - tc)
            e2e_parts["search_stage_sum_ms"] += r2["prefilter_stats"]["total_ms"] + r2["align_stats"]["total_ms"] + r2["cluster_stats"]["total_ms"]
            e2e_parts["teardown_ms"] += 1e3 * (te - td)
    barrier()
    dt_e2e = time.perf_counter() - t1
    clocks_e2e = sampler2.stop() if sampler2 else None

    # max over ranks
    t_all = torch.tensor([dt, dt_e2e], dtype=torch.float64, device="cuda")
    counts = torch.tensor([int(res["align_stats"]["alignments"]), 0, 0, nq_local], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)          # reporting only (outside the timed region)
    counts[1], counts[2] = len(res["alns"]), len(res["clusters"])      # after the gather every rank holds the whole job's lists
    dt, dt_e2e = float(t_all[0]), float(t_all[1])
    if rank == 0:
        pst, ast, cst = res["prefilter_stats"], res["align_stats"], res["cluster_stats"]
        peak, peak_src = peaks()
        algo = 8 * pst["probes"] + 6 * pst["matches"] + pst["diag_residues"] + 21 * pst["query_residues"] + 10 * pst["hits"]
        k_ms = pst["kernel_ms"]
        achieved = algo / 1e9 / (k_ms / 1e3) if k_ms > 0 else 0.0
        # DRAM bytes of this launch from ncu --set full (dram__bytes_read.sum + dram__bytes_write.sum, profiles/prof_pf_stream_c2_r02.raw.csv);
        # known only for the default workload on one GPU
        traffic, traffic_src = None, None
        if world == 1 and args.proteomes == N_PROTEOMES and args.genes == GENES:
            traffic = 504474968000 + 70755082000
            traffic_src = ("ncu dram__bytes_read.sum + dram__bytes_write.sum of the C2 prefilterStreamKernel launch (154.8 ms under ncu): 575 GB = "
                           "6.76 x algorithmic, 3.71 TB/s = 0.57 of the measured HBM peak; profiles/prof_pf_stream_c2_r02.raw.csv")
        # which of the two prefilter kernels took this index (csrc/prefilter.cu: streaming kernel up to 262144 targets, else binned)
        pf_kernel = "prefilterStreamKernel" if n_total <= 262144 else "prefilterKernel"
        cells = ast["cells_forward"] + ast["cells_reverse"]
        line = {"metric": "clustersearch query-genes/s", "value": n_total * args.steps / dt, "unit": "query-genes/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32 (u8/i16 scores)/f64",
                "data": "synthetic", "config": workload_config(args),
                "e2e": {"value": n_total * args.e2e_steps / dt_e2e, "unit": "query-genes/s", "h2d_bytes_per_step": e2e_h2d,
                        "d2h_bytes_per_step": e2e_d2h, "ms_per_step": 1e3 * dt_e2e / args.e2e_steps, "steps": args.e2e_steps,
                        "parts_ms_per_step": {k: v / args.e2e_steps for k, v in e2e_parts.items()}, "clocks": clocks_e2e,
                        "includes": "target index build from host residues (upload, tantan masking and k-mer index on the device), query upload, "
                                    "prefilter, align, besthitbyset/combinehits, clusterhits, results to host"},
                "gpu_launches": int(launches),
                "roofline": {"kernel": pf_kernel, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                             "dram_frac_of_peak": (traffic / 1e9 / (k_ms / 1e3) / peak) if (traffic and k_ms > 0) else None,
                             "algorithmic_bytes_per_launch": int(algo),
                             "bytes_formula": "8*P + 6*M + R + 21*Lq + 10*H (SURVEY.md 8d), P/M/R/H counted by the kernel and "
                                              "equal to the oracle's counts (tests/test_prefilter_gpu.py)",
                             "kernel_ms": k_ms,
                             "counts": {"P_probes": int(pst["probes"]), "M_index_entries": int(pst["matches"]),
                                        "R_diag_residues": int(pst["diag_residues"]), "Lq_residues": int(pst["query_residues"]),
                                        "H_hits": int(pst["hits"]), "diag_candidates": int(pst["candidates"])}},
                "sw": {"gcups": cells / 1e9 / (ast["sw_ms"] / 1e3) if ast["sw_ms"] > 0 else 0.0, "cells": int(cells),
                       "kernel_ms": ast["sw_ms"], "pairs": int(ast["alignments"]),
                       "algorithmic_bytes": int(ast["target_residues"]) + 64 * int(ast["alignments"]),
                       "hbm_gbs_algorithmic": (int(ast["target_residues"]) + 64.0 * int(ast["alignments"])) / 1e9 / (ast["sw_ms"] / 1e3)
                       if ast["sw_ms"] > 0 else 0.0,
                       "bytes_formula": "sum over aligned pairs of (L_t + 64) (SURVEY.md 8d); the pass is integer-issue bound, not HBM bound"},
                "stage_ms_per_step": {k: v / args.steps for k, v in stage.items()},
                "results": {"aligned_pairs": int(counts[0]), "alignments": int(counts[1]), "clusters": int(counts[2]),
                            "query_genes": int(counts[3])},
                "clocks": clocks}
        if world > 1:
            q_all = res["alns"]["query"]
            assert len(q_all) == 0 or (int(q_all.max()) < n_total and bool(np.all(np.diff(q_all.astype(np.int64)) >= 0)))
            line["collective"] = {"op": "two ncclAllGather exchanges (sizes, then padded blocks) issued by libsdgpu.so (sd_shard_clustersearch_run, "
                                        "csrc/fused.cu + csrc/shard.cu): alignment/backtrace lists after the align stage, cluster lists after clusterhits",
                                  "bytes_gathered_per_step": int(res["bytes_gathered"]),
                                  "sharding": "query genes in contiguous residue-balanced blocks for prefilter + align (sd_shard_partition_genes), "
                                              "query sets balanced by alignments for besthitbyset/combinehits/clusterhits; target index replicated"}
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline(args, proteomes)
            except Exception as e:      # the baseline is reporting only; never lose the GPU line
                line["cpu_baseline"] = {"error": str(e)[:300]}
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(args, proteomes):
    """oracle/_ref (the unmodified reference, AVX2) on this box's host cores, one bounded sample of the workload (about 25 s)."""
    n_sets = pick_ref_sample(args, 1, 25.0)
    runner = ReferenceRunner(args, proteomes, n_sets)
    try:
        dt, stages, n_clu = runner.step()
        return {"value": runner.n_queries / dt, "unit": "query-genes/s", "cores": runner.threads, "kind": "reference",
                "sample": "%d of %d query proteomes (%d query genes) against the full %d-proteome target setDB, one clustersearch run, "
                          "%.1f s wall; createsetdb excluded" % (n_sets, args.proteomes, runner.n_queries, args.proteomes, dt),
                "cpu": cpu_model(), "stage_seconds": stages, "clusters_in_sample": n_clu,
                "full_run_estimate": full_run_estimate(args, runner.n_queries, dt, stages)}
    finally:
        runner.close()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference" and rank != 0:
        return 0
    from spacedust_b200 import synth
    proteomes = synth.generate(args.proteomes, args.genes, seed=SEED)
    if args.impl == "reference":
        run_reference(args, proteomes, rank, world)
    else:
        run_ours(args, proteomes, rank, world, local_rank)
    return 0


if __name__ == "__main__":
    sys.exit(main())
