"""GPU: empty and ragged inputs through the C ABI (what the reference's modules meet with empty DB entries, single-entry DBs and
sequences shorter than the k-mer span)."""
import numpy as np
import pytest

import helpers
import spacedust_b200 as sd
from spacedust_b200 import synth

pytestmark = pytest.mark.gpu


def _meta_for(n, n_sets=1):
    per = (n + n_sets - 1) // n_sets
    set_of = np.minimum(np.arange(n) // max(per, 1), n_sets - 1).astype(np.uint32)
    pos = np.zeros(n, dtype=np.uint32)
    for s in range(n_sets):
        idx = np.nonzero(set_of == s)[0]
        pos[idx] = np.arange(len(idx))
    return {"set_of": set_of, "pos": pos, "strand": np.ones(n, dtype=np.uint8),
            "set_size": np.bincount(set_of, minlength=n_sets).astype(np.uint32)}


def test_single_target_and_tiny_queries(gpu_ctx):
    """One target; queries shorter than the spaced seed (no k-mer at all), one exact copy, one unrelated."""
    target = "MKTAYIAKQRQISFVKSHFSRQLEERLGLIEVQAPILSRVGDGTQDNLSGAEKAVQVKVKALPDAQFEVVHSLAKWKRQTLGQHDFSAGEGLYTHMKALRPDEDRLSPLHSVYVDQWDW"
    tbuf, toff = helpers.pack_seqs([target])
    qbuf, qoff = helpers.pack_seqs(["M", "MKT", "ACDEFGHIK", target, "GGGGGGGGGGSSSSSSSSSSGGGGGGGGGG"])
    pp = sd.default_prefilter_params(same_db=0, bin_size=2)
    idx = sd.TargetIndex(gpu_ctx, tbuf, toff, pp)
    try:
        hits, st = sd.prefilter(gpu_ctx, idx, qbuf, qoff, pp)
        ohits, ost = helpers.orc_prefilter(tbuf, toff, qbuf, qoff, helpers.default_pref_params(sameDB=0))
        assert len(hits) == len(ohits) == 1 and int(hits["query"][0]) == 3 and int(hits["target"][0]) == 0
        assert int(hits["score"][0]) == int(ohits["score"][0])
        assert st["probes"] == ost["probes"] and st["matches"] == ost["matches"]
        # align: no pairs at all is a valid call
        ap = sd.default_align_params(same_db=0)
        alns, pool, ast = sd.align(gpu_ctx, idx, qbuf, qoff, np.zeros(0, dtype=np.uint32), np.zeros(0, dtype=np.uint32), ap)
        assert len(alns) == 0 and ast["alignments"] == 0
        alns, pool, ast = sd.align(gpu_ctx, idx, qbuf, qoff, hits["query"], hits["target"], ap)
        assert len(alns) == 1 and int(alns["q_start"][0]) == 0 and int(alns["t_end"][0]) == len(target) - 1
        assert sd.format_alignment(alns[0], 0, pool).decode().rstrip("\n").endswith("\t%dM" % len(target))
    finally:
        idx.close()


def test_no_hits_anywhere(gpu_ctx):
    """Unrelated low-complexity-free random proteins: the whole fused path must come back empty, not fail."""
    rng = np.random.default_rng(77)
    seqs = [synth.LETTERS[rng.choice(20, size=int(rng.integers(30, 90)), p=synth.BACKGROUND)].tobytes().decode() for _ in range(12)]
    buf, off = helpers.pack_seqs(seqs)
    meta = _meta_for(len(seqs), n_sets=3)
    pp = sd.default_prefilter_params(same_db=1, bin_size=2)
    idx = sd.TargetIndex(gpu_ctx, buf, off, pp)
    batch = sd.QueryBatch(gpu_ctx, buf, off)
    try:
        r = sd.clustersearch(gpu_ctx, idx, batch, meta, meta, pp, clu_params=sd.default_cluster_params(filter_self_match=1))
        # every gene still finds itself (identity hit), but self matches are filtered and nothing else clusters
        assert len(r["alns"]) >= len(seqs) and len(r["clusters"]) == 0 and len(r["members"]) == 0
        o = helpers.orc_prefilter(buf, off, buf, off, helpers.default_pref_params())[0]
        assert r["prefilter_stats"]["hits"] >= len(o)
    finally:
        batch.close()
        idx.close()


def test_bad_arguments_fail_loudly(gpu_ctx):
    buf, off = helpers.pack_seqs(["MKTAYIAKQRQISFVKSHFSRQ", "MKTAYIAKQRQISFVKSHFSRQ"])
    pp = sd.default_prefilter_params(same_db=1)
    idx = sd.TargetIndex(gpu_ctx, buf, off, pp)
    try:
        ap = sd.default_align_params(same_db=1)
        with pytest.raises(sd.SdError):      # pair indices out of range
            sd.align(gpu_ctx, idx, buf, off, np.array([5], dtype=np.uint32), np.array([0], dtype=np.uint32), ap)
        ap2 = sd.default_align_params(same_db=1, gap_open=1, gap_extend=3)
        with pytest.raises(sd.SdError):      # gap open < gap extend is rejected, not silently mis-scored
            sd.align(gpu_ctx, idx, buf, off, np.array([0], dtype=np.uint32), np.array([1], dtype=np.uint32), ap2)
        ap3 = sd.default_align_params(same_db=1, gap_open=9, gap_extend=2)
        with pytest.raises(sd.SdError) as e:  # the reference simulates Gumbel parameters for other gap costs: refused
            sd.align(gpu_ctx, idx, buf, off, np.array([0], dtype=np.uint32), np.array([1], dtype=np.uint32), ap3)
        assert "gap" in str(e.value)
    finally:
        idx.close()
