"""GPU parity: besthitbyset/combinehits (host) + clusterhits (device kernel) through the C ABI vs the literal
Python restatement of the reference modules."""
import numpy as np
import pytest

import helpers
import spacedust_b200 as sd
from spacedust_b200 import synth

pytestmark = pytest.mark.gpu


def _gpu_clusters_as_text(gpu_ctx, alns_orc, bt, meta, **kw):
    """Feed the oracle's accepted alignments to sd_cluster_run and render clusters like clusterhits does."""
    g = np.zeros(len(alns_orc), dtype=sd.ALN_DTYPE)
    for fo, fg in (("query", "query"), ("target", "target"), ("score", "score"), ("bits", "bits"), ("seqId", "seq_id"),
                   ("qcov", "qcov"), ("tcov", "tcov"), ("evalue", "evalue"), ("qStart", "q_start"), ("qEnd", "q_end"),
                   ("qLen", "q_len"), ("tStart", "t_start"), ("tEnd", "t_end"), ("tLen", "t_len"), ("alnLen", "aln_len"),
                   ("btOff", "bt_off"), ("btLen", "bt_len")):
        g[fg] = alns_orc[fo]
    par = sd.default_cluster_params(**kw)
    clusters, members, st = sd.cluster(gpu_ctx, g, meta, meta, par)
    out = []
    for c in clusters:
        lines = []
        for m in members[int(c["first"]):int(c["first"]) + int(c["n"])]:
            a = g[int(m["aln"])]
            line = sd.format_alignment(a, a["target"], bt).decode().rstrip("\n").split("\t")
            # clusterhits records: qKey tKey pval seqId eval qS qE qL tS tE tL cigar
            lines.append("\t".join([str(int(a["query"])), line[0], "%.3E" % m["pval"]] + line[2:]))
        out.append(("%d\t%d\t%.3E\t%.3E\t%d" % (c["q_set"], c["t_set"], c["p_co"], c["p_mh"], c["n"]), lines))
    return out, st


@pytest.mark.parametrize("n_prot,genes,seed", [(3, 120, 7), (4, 150, 31)])
def test_cluster_matches_oracle(gpu_ctx, n_prot, genes, seed):
    prot = synth.generate(n_prot, genes, seed=seed)
    o = helpers.oracle_clustersearch(prot, filter_self_match=True)
    mine, st = _gpu_clusters_as_text(gpu_ctx, o["alns"], o["bt"], o["meta"], filter_self_match=1)
    assert st["pairs_clustered"] > 0 and st["merges"] > 0 and st["inexact_log"] == 0
    assert helpers.canonical_clusters(mine) == helpers.canonical_clusters(o["clusters"])
    assert len(mine) > 0


def test_cluster_self_pairs_and_loose_thresholds(gpu_ctx):
    """Without --filter-self-match the self pairs (every gene hits itself, K = genome size) are clustered too;
    loose P-value thresholds and a larger gene gap exercise merges below sMin and long merge chains."""
    prot = synth.generate(2, 90, seed=5)
    for kw_o, kw_g in ((dict(filter_self_match=False), dict(filter_self_match=0)),
                       (dict(filter_self_match=True, max_gene_gap=10, cluster_size=3, cluster_pval=1.0, multihit_pval=1.0),
                        dict(filter_self_match=1, max_gene_gap=10, cluster_size=3, cluster_pval=1.0, multihit_pval=1.0))):
        o = helpers.oracle_clustersearch(prot, **kw_o)
        mine, st = _gpu_clusters_as_text(gpu_ctx, o["alns"], o["bt"], o["meta"], **kw_g)
        assert helpers.canonical_clusters(mine) == helpers.canonical_clusters(o["clusters"])
