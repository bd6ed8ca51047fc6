"""GPU: the fused clustersearch call (hit list kept in HBM) gives exactly what the three staged C-ABI calls give,
and what the oracle pipeline gives."""
import numpy as np
import pytest

import helpers
import spacedust_b200 as sd
from spacedust_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("chunks", ["1", "3"])
def test_fused_equals_staged_and_oracle(gpu_ctx, synth_small, monkeypatch, chunks):
    """chunks=3: the fused call cuts the batch into query-set chunks processed by two host threads / contexts."""
    monkeypatch.setenv("SD_CHUNKS", chunks)
    buf, off, _ = synth.pack(synth_small)
    meta = synth.set_metadata(synth_small)
    pp = sd.default_prefilter_params(same_db=1, bin_size=2)
    idx = sd.TargetIndex(gpu_ctx, buf, off, pp)
    batch = sd.QueryBatch(gpu_ctx, buf, off)
    try:
        a = sd.clustersearch_staged(gpu_ctx, idx, batch, meta, meta, pp)
        b = sd.clustersearch(gpu_ctx, idx, batch, meta, meta, pp)
    finally:
        batch.close()
        idx.close()
    assert len(a["alns"]) == len(b["alns"]) and len(a["clusters"]) == len(b["clusters"]) > 0
    for f in a["alns"].dtype.names:
        assert np.array_equal(a["alns"][f], b["alns"][f]), f
    assert np.array_equal(a["pool"], b["pool"])
    for f in a["clusters"].dtype.names:
        assert np.array_equal(a["clusters"][f], b["clusters"][f]), f
    for f in ("aln", "pval"):
        assert np.array_equal(a["members"][f], b["members"][f]), f
    # and against the oracle's whole pipeline (cluster membership + printed P-values)
    o = helpers.oracle_clustersearch(synth_small, filter_self_match=True)
    mine = helpers.gpu_clusters(b)
    assert helpers.canonical_clusters(mine) == helpers.canonical_clusters(o["clusters"])
