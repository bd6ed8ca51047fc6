"""GPU parity: CUDA align stage (Smith-Waterman passes + banded traceback, through the C ABI) vs the oracle."""
import itertools

import numpy as np
import pytest

import helpers
import spacedust_b200 as sd
from spacedust_b200 import synth

pytestmark = pytest.mark.gpu


def _pairs_all_vs_all(nq, nt):
    q = np.repeat(np.arange(nq, dtype=np.uint32), nt)
    t = np.tile(np.arange(nt, dtype=np.uint32), nq)
    return q, t


def _compare(gpu_ctx, tbuf, toff, qbuf, qoff, pq, pt, same_db, **kw):
    nT, nQ = len(toff) - 1, len(qoff) - 1
    tkeys = np.arange(nT, dtype=np.uint32)
    qkeys = np.arange(nQ, dtype=np.uint32)
    op = helpers.default_align_params(sameDB=1 if same_db else 0, **kw)
    oal, obt, ost = helpers.orc_align(tbuf, toff, tkeys, qbuf, qoff, qkeys, pq, pt, op)
    gp = sd.default_align_params(same_db=1 if same_db else 0, evalue_thr=op.evalThr, cov_thr=op.covThr, cov_mode=op.covMode,
                                 min_aln_len=op.minAlnLen, comp_bias_corr=op.compBias, min_seq_id=op.seqIdThr)
    idx = sd.TargetIndex(gpu_ctx, tbuf, toff, sd.default_prefilter_params())
    try:
        gal, gbt, gst = sd.align(gpu_ctx, idx, qbuf, qoff, pq, pt, gp)
    finally:
        idx.close()
    assert ost["alignments"] == gst["alignments"]
    assert ost["wordMode"] == gst["word_mode"], (ost["wordMode"], gst["word_mode"])
    assert ost["reversePasses"] == gst["reverse_passes"]
    assert ost["tracebacks"] == gst["tracebacks"]
    assert ost["cellsTraceback"] == gst["cells_traceback"]
    assert len(oal) == len(gal), (len(oal), len(gal))
    for fo, fg in (("query", "query"), ("target", "target"), ("score", "score"), ("bits", "bits"), ("qStart", "q_start"),
                   ("qEnd", "q_end"), ("tStart", "t_start"), ("tEnd", "t_end"), ("alnLen", "aln_len"), ("qLen", "q_len"),
                   ("tLen", "t_len"), ("seqId", "seq_id"), ("evalue", "evalue"), ("qcov", "qcov"), ("tcov", "tcov")):
        assert np.array_equal(oal[fo], gal[fg]), "field %s differs" % fo
    # the library returns the backtrace run-length encoded (CIGAR text, Matcher::compressAlignment done on the device)
    assert np.all(gal["bt_cigar"] == 1)
    for i in range(len(oal)):
        raw = bytes(obt[int(oal[i]["btOff"]):int(oal[i]["btOff"]) + int(oal[i]["btLen"])]).decode()
        cigar = "".join("%d%s" % (len(list(g)), k) for k, g in itertools.groupby(raw))
        b = bytes(gbt[int(gal[i]["bt_off"]):int(gal[i]["bt_off"]) + int(gal[i]["bt_len"])]).decode()
        assert cigar == b, "backtrace %d differs" % i
    # the text lines are what the reference writes to the result DB
    omine = helpers.format_alignments(oal, obt, qkeys)
    for i in range(min(len(gal), 50)):
        line = sd.format_alignment(gal[i], gal[i]["target"], gbt)
        assert line[:-1] in omine[int(gal[i]["query"])]
    return ost


def test_align_matches_oracle_on_prefilter_pairs(gpu_ctx, synth_small):
    buf, off, _ = synth.pack(synth_small)
    hits, _ = helpers.orc_prefilter(buf, off, buf, off, helpers.default_pref_params())
    st = _compare(gpu_ctx, buf, off, buf, off, hits["query"], hits["target"], True)
    assert st["passed"] > len(off) - 1 and st["wordMode"] > 0 and st["tracebacks"] > 0


def test_align_all_vs_all_small(gpu_ctx):
    """Every pair of a small set (unrelated pairs exercise the early outs, relaxed thresholds the traceback)."""
    prot = synth.generate(2, 40, seed=23)
    buf, off, _ = synth.pack(prot)
    n = len(off) - 1
    pq, pt = _pairs_all_vs_all(n, n)
    _compare(gpu_ctx, buf, off, buf, off, pq, pt, True)
    _compare(gpu_ctx, buf, off, buf, off, pq, pt, False, evalThr=1e6, covThr=0.0, minAlnLen=0)


def test_align_long_and_edge_sequences(gpu_ctx):
    """Queries longer than one 1024-row stripe, very short sequences, X runs."""
    rng = np.random.default_rng(3)
    anc = rng.choice(20, size=2600, p=synth.BACKGROUND).astype(np.uint8)
    seqs = [synth.LETTERS[anc].tobytes().decode(),
            synth.LETTERS[synth._mutate(rng, anc, 0.7)].tobytes().decode(),
            synth.LETTERS[synth._mutate(rng, anc[:1500], 0.5)].tobytes().decode(),
            synth.LETTERS[synth._mutate(rng, anc[900:2300], 0.6)].tobytes().decode(),
            "MKV", "ACDEFGHIKLMNPQRSTVWY", "XXXXXXXXXXXXXXXXXXXXXXXXXXXXXXXXXXXXXXXX",
            "MKTAYIAKQRQISFVKSHFSRQLEERLGLIEVQAPILSRVGDGTQDNLSGAEKAVQVKVKALPDAQFEVVHSLAKWKRQTLGQHDFSAGEGLYTHMKALRPDEDRLSPLHSVYVDQWDW",
            "MKTAYIAKQRQISFVKSHFSRQLEERLGLIEVQAPILSRVGDGTQDNLSGAEKAVQVKVKALPDAQXXXXFEVVHSLAKWKRQTLGQHDFSAGEGLYTHMKALRPDEDRLSPLHSVYVDQWDW"]
    buf, off = helpers.pack_seqs(seqs)
    n = len(seqs)
    pq, pt = _pairs_all_vs_all(n, n)
    _compare(gpu_ctx, buf, off, buf, off, pq, pt, True, covThr=0.0, evalThr=1e3, minAlnLen=0)
    _compare(gpu_ctx, buf, off, buf, off, pq, pt, False, covThr=0.3, covMode=0, evalThr=10.0)
