"""GPU parity on the BASELINE.json configurations that round 1 left to builder-run sweeps: the pair search at beam
width 100, the `row` traversal at T = 5000, long synthetic pairs (T = 50k and 100k, padding 150: bands hundreds of
timesteps wide, the window spill path) and the single-read search over a batch of reads -- each against the unmodified
reference core (oracle/_ref; the oracle port where that was not built), run on the host cores in a process pool."""
import multiprocessing as mp
import os

import numpy as np
import pytest

from poreover_b200 import _lib, batch, synth

pytestmark = pytest.mark.gpu
TOL = 1e-4
BAD = _lib.ST_POOL_OVERFLOW | _lib.ST_SHORT_BEAM_SKIP | _lib.ST_UNSET_BAND


def _pair(k, T):
    from oracle import oracle as O
    p1, p2, _ = synth.make_pair(k, T)
    return synth.bonito_log_prob(p1), np.ascontiguousarray(O.reverse_complement(synth.bonito_log_prob(p2), "bonito"))


def _ref_pair(job):
    """(k, T, W, padding, method) -> (consensus, score, envelope) from the reference core"""
    from oracle import oracle as O
    k, T, W, pad, method = job
    backend = "ref" if O.have_ref() else "port"
    lp1, lp2 = _pair(k, T)
    r = O.pair_decode(lp1, lp2, "bonito", W, padding=pad, method=method, backend=backend, with_score=True)
    return r["consensus"], r["score"], r["envelope"]


def _ref_read(job):
    from oracle import oracle as O
    i, T, W = job
    backend = "ref" if O.have_ref() else "port"
    return O.beam_search(synth.bonito_log_prob(synth.make_read(i, T)[0]), W, "ctc_merge_repeats", backend, True)


def _pool_map(fn, jobs):
    with mp.get_context("fork").Pool(min(len(jobs), os.cpu_count() or 1)) as pool:
        return pool.map(fn, jobs, chunksize=1)


def test_pair_search_width_100_and_row_traversal(oracle):
    """BASELINE configs[2] shape (T ~ 5000) at beam width 100 (row_col), and the `row` traversal (BeamSearch.h:110-172,
    what cpp_beam_search_2d runs by default) at beam width 25."""
    jobs = [(120, 5000, 100, 5, "row_col"), (121, 5000, 25, 5, "row"), (122, 5000, 25, 5, "row")]
    want = _pool_map(_ref_pair, jobs)
    for (k, T, W, pad, method), (cons, score, env) in zip(jobs, want):
        lp1, lp2 = _pair(k, T)
        seqs, sc, st = batch.beam_search_2d_batch([lp1], [lp2], [env], W, "ctc_merge_repeats", method)
        assert seqs[0] == cons, (k, W, method)
        assert abs(sc[0] - score) < TOL, (k, W, method, sc[0], score)
        assert not (st[0] & BAD)


def test_long_pairs_T50k_T100k_wide_envelopes(oracle):
    """BASELINE configs[3] shape: T = 50,000 and T = 100,000 synthetic pairs, padding 150 (every band ~300 timesteps
    wide), beam width 5 so that the CPU side finishes in about a minute; the whole fused pipeline."""
    jobs = [(7001, 50000, 5, 150, "row_col"), (7002, 100000, 5, 150, "row_col")]
    want = _pool_map(_ref_pair, jobs)
    for (k, T, W, pad, method), (cons, score, env) in zip(jobs, want):
        assert (env[:, 1] - env[:, 0]).max() >= 300
        p1, p2, _ = synth.make_pair(k, T)
        r = batch.pair_decode_batch([synth.bonito_log_prob(p1)], [synth.bonito_log_prob(p2)], "bonito", beam_width=W,
                                    padding=pad, rc2=True)[0]
        assert r["consensus"] == cons, (T, len(r["consensus"]), len(cons))
        assert abs(r["score"] - score) < TOL * max(1.0, T / 5000.0), (T, r["score"], score)
        assert not (r["status"] & BAD)


def test_single_read_search_batch_of_32(oracle):
    """BASELINE configs[1] shape: the single-read prefix search over a batch (32 reads, T = 3000, beam width 25)."""
    jobs = [(1200 + i, 3000, 25) for i in range(32)]
    want = _pool_map(_ref_read, jobs)
    arrays = [synth.bonito_log_prob(synth.make_read(i, T)[0]) for i, T, _ in jobs]
    seqs, sc, st = batch.beam_search_batch(arrays, 25, "ctc_merge_repeats")
    for g, gs, (w, ws) in zip(seqs, sc, want):
        assert g == w and abs(gs - ws) < TOL
    assert not (st & _lib.ST_POOL_OVERFLOW).any()


def test_small_alphabet_and_narrow_beams(oracle):
    """The reference interface takes any alphabet and beam width (decoding_cpp.pyx:88-139; its own tests use "AB"):
    3-state matrices and beam widths 1, 2, 3 against the oracle."""
    from poreover_b200.decoding import decoding_cpp
    rng = np.random.default_rng(5)
    for T in (6, 40):
        y = rng.dirichlet(np.ones(3) * 0.6, size=T)
        y2 = rng.dirichlet(np.ones(3) * 0.6, size=T + 3)
        ly, ly2 = np.log(y), np.log(y2)
        for model in ("ctc", "ctc_merge_repeats"):
            for W in (1, 2, 3, 25):
                want = oracle.beam_search(ly, W, model)
                got = decoding_cpp.cpp_beam_search(ly, W, "AB", model)
                assert got == want.translate(str.maketrans("AC", "AB")), (T, model, W)
                env = np.array([(max(0, i - 4), min(T + 3, i + 5)) for i in range(T)])
                for method in ("row", "row_col"):
                    want2 = oracle.beam_search_2d(ly, ly2, env, W, model, method)
                    got2 = decoding_cpp.cpp_beam_search_2d(ly, ly2, env.tolist(), W, "AB", model, method)
                    assert got2 == want2.translate(str.maketrans("AC", "AB")), (T, model, W, method)
    y5 = np.log(rng.dirichlet(np.ones(5) * 0.5, size=60))
    for W in (1, 2, 3):
        assert decoding_cpp.cpp_beam_search(y5, W, "ACGT", "ctc_merge_repeats") == oracle.beam_search(y5, W, "ctc_merge_repeats")
