"""GPU parity: CUDA prefilter (through the C ABI) vs the scalar oracle on the same seeded inputs."""
import os
import subprocess

import numpy as np
import pytest

import helpers
import spacedust_b200 as sd
from spacedust_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["stream", "binned"], autouse=True)
def kernel_variant(request, monkeypatch):
    """Both prefilter kernels must give the reference's result: the streaming kernel (small target sets; it hands queries that
    reach the reference's hit-buffer limit to the binned kernel) and the binned kernel (SD_PF_STREAM_MAX=0 forces it)."""
    if request.param == "binned":
        monkeypatch.setenv("SD_PF_STREAM_MAX", "0")
    else:
        monkeypatch.delenv("SD_PF_STREAM_MAX", raising=False)
    return request.param


def _compare(gpu_ctx, tbuf, toff, qbuf, qoff, same_db, **kw):
    op = helpers.default_pref_params(sameDB=1 if same_db else 0, **kw)
    if op.maxDbMatches:
        os.environ["SD_PF_MAXDBMATCHES"] = str(op.maxDbMatches)       # test hook of the library, same knob as the oracle's
    else:
        os.environ.pop("SD_PF_MAXDBMATCHES", None)
    ohits, ost, omask = helpers.orc_prefilter(tbuf, toff, qbuf, qoff, op, want_masked=True)
    gp = sd.default_prefilter_params(same_db=1 if same_db else 0, bin_size=op.binSize, max_seqs=op.maxSeqs,
                                     min_ungapped_score=op.minDiagScore, cov_thr=op.covThr, cov_mode=op.covMode,
                                     mask=op.mask, comp_bias_corr=op.compBias, sensitivity=op.sensitivity)
    idx = sd.TargetIndex(gpu_ctx, tbuf, toff, gp)
    try:
        gmask = idx.masked_residues()
        assert np.array_equal(gmask, omask), "masked target residues differ"
        ghits, gst = sd.prefilter(gpu_ctx, idx, qbuf, qoff, gp)
    finally:
        idx.close()
        os.environ.pop("SD_PF_MAXDBMATCHES", None)
    for k_o, k_g in (("probes", "probes"), ("matches", "matches"), ("candidates", "candidates"),
                     ("diagResidues", "diag_residues"), ("hits", "hits"), ("truncated", "truncated")):
        assert ost[k_o] == gst[k_g], "counter %s: oracle %d gpu %d" % (k_o, ost[k_o], gst[k_g])
    assert len(ohits) == len(ghits)
    for f in ("query", "target", "score", "diag"):
        assert np.array_equal(ohits[f], ghits[f]), "hit field %s differs" % f
    return ost


def test_prefilter_matches_oracle_self(gpu_ctx, synth_small):
    buf, off, _ = synth.pack(synth_small)
    st = _compare(gpu_ctx, buf, off, buf, off, True)
    assert st["hits"] > len(off) - 1


def test_prefilter_matches_oracle_q_vs_t(gpu_ctx):
    prot = synth.generate(4, 90, seed=11)
    tbuf, toff, _ = synth.pack(prot[:3])
    qbuf, qoff, _ = synth.pack(prot[3:])
    _compare(gpu_ctx, tbuf, toff, qbuf, qoff, False)


def _weak_family_set(n_copies=120, n_unrelated=30, seed=5):
    """One ancestral protein in many diverged copies (50-60 % identity: diagonal scores well below the
    8-bit cap) plus unrelated proteins -> queries with many moderate hits, so --max-seqs truncates inside
    a score class."""
    rng = np.random.default_rng(seed)
    anc = rng.choice(20, size=100, p=synth.BACKGROUND).astype(np.uint8)
    seqs = [synth.LETTERS[synth._mutate(rng, anc, rng.uniform(0.5, 0.6))].tobytes().decode() for _ in range(n_copies)]
    seqs += [synth.LETTERS[rng.choice(20, size=int(rng.integers(80, 500)), p=synth.BACKGROUND)].tobytes().decode()
             for _ in range(n_unrelated)]
    return helpers.pack_seqs(seqs)


def test_prefilter_truncation_tie_order(gpu_ctx):
    """--max-seqs small enough that the threshold class is cut: survivors follow the reference's
    (id & binMask, stream order) rule."""
    buf, off = _weak_family_set()
    st = _compare(gpu_ctx, buf, off, buf, off, True, maxSeqs=16)
    assert st["truncated"] > 10
    _compare(gpu_ctx, buf, off, buf, off, True, maxSeqs=16, binSize=4)
    _compare(gpu_ctx, buf, off, buf, off, True, maxSeqs=12, minDiagScore=0, binSize=8)


def test_prefilter_hit_buffer_overflow(gpu_ctx, synth_small):
    """The reference buffers at most maxDbMatches index entries per query (QueryMatcher.cpp:43,280-333); beyond that the
    double-hit table restarts and the candidate lists are merged, from the second overflow on through the reversed merge +
    intermediate keepMax (:296-303).  The oracle restates that path (pinned on the reference binary by tests/test_oracle_cpu.py
    with queries of > 2*10^6 and > 4*10^6 entries); here the limit is lowered in both (test hook) and the GPU path must follow,
    hits and counters: once with single overflows only, once with up to a dozen overflows per query."""
    buf, off, _ = synth.pack(synth_small)
    single = multi = None
    for limit in (3000, 2000, 1500, 1000, 700, 500, 350, 250, 150, 100, 60):
        _, st = helpers.orc_prefilter(buf, off, buf, off, helpers.default_pref_params(maxDbMatches=limit))
        if single is None and st["overflowQueries"] >= 10 and st["multiOverflowQueries"] == 0:
            single = limit
        if st["multiOverflowQueries"] >= 30:
            multi = limit
            break
    assert single is not None and multi is not None, (single, multi)
    _compare(gpu_ctx, buf, off, buf, off, True, maxDbMatches=single)
    st = _compare(gpu_ctx, buf, off, buf, off, True, maxDbMatches=multi)
    assert st["multiOverflowQueries"] >= 30
    st = _compare(gpu_ctx, buf, off, buf, off, True, maxDbMatches=150)
    assert st["multiOverflowQueries"] >= 200
    # tie order of the --max-seqs cut after repeated overflows (reversed blocks), other radix
    fbuf, foff = _weak_family_set()
    _compare(gpu_ctx, fbuf, foff, fbuf, foff, True, maxSeqs=16, maxDbMatches=1500)
    _compare(gpu_ctx, fbuf, foff, fbuf, foff, True, maxSeqs=16, binSize=4, maxDbMatches=700)


def test_prefilter_candidate_overflow_second_pass(gpu_ctx, synth_small, monkeypatch):
    """Queries whose double-hit candidate list exceeds the first-pass capacity are re-run by a second launch with a
    larger list: hits AND the P/M/R/H counters must still equal the oracle's (SD_PF_CANDCAP forces the situation)."""
    monkeypatch.setenv("SD_PF_CANDCAP", "32")
    monkeypatch.setenv("SD_PF_REGIONCAP", "64")
    buf, off, _ = synth.pack(synth_small)
    st = _compare(gpu_ctx, buf, off, buf, off, True)
    assert st["candidates"] > 32 * 10
    fbuf, foff = _weak_family_set()
    _compare(gpu_ctx, fbuf, foff, fbuf, foff, True, maxSeqs=16)


def test_prefilter_truncated_scores(gpu_ctx):
    """More than --max-seqs targets at the 8-bit cap (close copies of one protein): the rescoreHits path
    (QueryMatcher.cpp:157-171,525-544) -- uncapped scores rescaled against the query's self score -- must give the oracle's
    hits (the oracle is pinned on the reference binary for this case by tests/test_oracle_cpu.py)."""
    rng = np.random.default_rng(9)
    anc = rng.choice(20, size=300, p=synth.BACKGROUND).astype(np.uint8)
    seqs = [synth.LETTERS[synth._mutate(rng, anc, rng.uniform(0.88, 0.97))].tobytes().decode() for _ in range(60)]
    seqs += [synth.LETTERS[rng.choice(20, size=int(rng.integers(80, 400)), p=synth.BACKGROUND)].tobytes().decode() for _ in range(20)]
    buf, off = helpers.pack_seqs(seqs)
    st = _compare(gpu_ctx, buf, off, buf, off, True, maxSeqs=16)
    assert st["truncated"] >= 60
    _compare(gpu_ctx, buf, off, buf, off, False, maxSeqs=24, binSize=4)


def test_prefilter_edge_sequences(gpu_ctx):
    """Short (< k-mer span), X-rich, low-complexity and duplicated sequences."""
    long_seq = ("MKTAYIAKQRQISFVKSHFSRQLEERLGLIEVQAPILSRVGDGTQDNLSGAEKAVQVKVKALPDAQFEVVHSLAKWKRQTLGQHDFSAGEGLYTHMKALRPDEDRLSPLHSVYV"
                "DQWDWERVMGDGERQFSTLKSTVEAIWAGIKATEAAVSEEFGLAPFLPDQIHFVHSQELLSRYPDLDAKGRERAIAKDLGAVFLVGIGGKLSDGHRHDVRAPDYDDWSTPSELGHAG"
                "LNGDILVWNPVLEDAFELSSMGIRVDADTLKHQLALTGDEDRLELEWHQALLRGEMPQTIGGGIGQSRLTMLLLQLPHIGQVQAGVWPAAVRESVPSLL*")
    seqs = ["MKV", "ACDEFGHIK", "ACDEFGHIKL", "XXXXXXXXXXXXXXXXXXXXXXXX", "A" * 200,
            "QQQQQQQQQQPPPPPPPPPPQQQQQQQQQQPPPPPPPPPP" * 3, long_seq, long_seq,
            long_seq[:150] + "GGSGGS" + long_seq[150:], long_seq[40:], "ubzjo" + long_seq[5:200].lower()]
    buf, off = helpers.pack_seqs(seqs)
    _compare(gpu_ctx, buf, off, buf, off, True)


def test_no_oracle_symbols_in_product():
    """The product library must not pull anything from oracle/."""
    out = subprocess.run(["nm", "-D", sd.LIB_PATH], stdout=subprocess.PIPE).stdout.decode()
    assert "orc_" not in out
