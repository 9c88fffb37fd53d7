"""CPU: the batched file ingest of the command-line drivers (poreover_b200/ingest.py).
Every array that reaches the C ABI must be bit-identical to what the reference's loader path
(decode.model_from_trace -> transducer) produces; only the column order of plain bonito tables may differ, and the
batch says so."""
import os
import threading
import time

import numpy as np
import pytest

from poreover_b200 import _lib, batch, ingest, synth
from poreover_b200.decoding import decode as gdecode
from poreover_b200.decoding import transducer


def _reference_arrays(paths, basecaller):
    return [gdecode.model_from_trace(p, basecaller).device_array() for p in paths]


def _rows(b, i):
    return b.data[b.row_off[i]:b.row_off[i] + b.lens[i]]


def test_bonito_tables_keep_file_order_and_values(tmp_path):
    paths = []
    for k in range(6):
        f1, f2 = synth.save_pair(str(tmp_path), k, 200 + 17 * k)
        paths += [str(tmp_path / f1), str(tmp_path / f2)]
    ref = _reference_arrays(paths, "bonito")
    b = ingest.load_reads(paths, "bonito", rc=1)
    assert b.layout == _lib.BLANK_FIRST and b.dtype == _lib.F32 and b.n == len(paths)
    assert b.kinds == ["bonito"] * len(paths) and b.rc.tolist() == [1] * len(paths)
    assert all(o % batch.ALIGN_ROWS == 0 for o in b.row_off)
    for i, a in enumerate(ref):
        got = _rows(b, i)[:, [1, 2, 3, 4, 0]]  # decode.py:79 applied to the file-order rows
        assert got.dtype == a.dtype and np.array_equal(got, a)
    # padding rows stay zero, like ReadBatch
    same = batch.ReadBatch([a[:, [4, 0, 1, 2, 3]] for a in ref])
    assert np.array_equal(same.data, b.data) and np.array_equal(same.row_off, b.row_off)


def test_poreover_tables_and_exact_zero_probabilities(tmp_path):
    rng = np.random.default_rng(3)
    paths = []
    for k in range(4):
        p = rng.dirichlet(np.ones(5), size=50 + k).astype(np.float32)
        p[3, 2] = 0.0  # log(0) = -inf must come through without a warning turning into an error
        np.save(tmp_path / ("p%d.npy" % k), p)
        paths.append(str(tmp_path / ("p%d.npy" % k)))
    ref = _reference_arrays(paths, "poreover")
    b = ingest.load_reads(paths, "poreover")
    assert b.layout == _lib.BLANK_LAST and b.kinds == ["poreover"] * 4
    for i, a in enumerate(ref):
        assert np.array_equal(_rows(b, i), a)
        assert np.isneginf(_rows(b, i)[3, 2])


def test_unusual_files_take_the_reference_loader_path(tmp_path):
    rng = np.random.default_rng(4)
    plain = rng.dirichlet(np.ones(5), size=40).astype(np.float32)
    f64 = rng.dirichlet(np.ones(5), size=33)                               # float64 probabilities
    logits3d = rng.normal(size=(3, 20, 5)).astype(np.float32)              # windows x time x states, not normalised
    fortran = np.asfortranarray(rng.dirichlet(np.ones(5), size=25).astype(np.float32))
    names = {"a.npy": plain, "b.npy": f64, "c.npy": logits3d, "d.npy": fortran}
    paths = []
    for n, a in names.items():
        np.save(tmp_path / n, a)
        paths.append(str(tmp_path / n))
    for basecaller in ("bonito", "poreover"):
        ref = _reference_arrays(paths, basecaller)
        b = ingest.load_reads(paths, basecaller)
        assert b.layout == _lib.BLANK_LAST and b.dtype == _lib.F64  # a mixed batch is widened, exactly
        assert b.lens.tolist() == [40, 33, 60, 25]
        for i, a in enumerate(ref):
            assert np.array_equal(_rows(b, i), a.astype(np.float64)), (basecaller, i)
    # the same batch through ReadBatch gives the same bytes
    rb = batch.ReadBatch(_reference_arrays(paths, "bonito"))
    assert np.array_equal(rb.data, ingest.load_reads(paths, "bonito").data)


def test_csv_and_flipflop(tmp_path):
    rng = np.random.default_rng(6)
    p5 = rng.dirichlet(np.ones(5), size=12)
    np.savetxt(tmp_path / "x.csv", p5, delimiter=",", header="a,c,g,t,b")
    b = ingest.load_reads([str(tmp_path / "x.csv")], "poreover")
    assert b.kinds == ["poreover"] and np.array_equal(_rows(b, 0), np.log(p5))
    p8 = rng.dirichlet(np.ones(8), size=12)
    np.savetxt(tmp_path / "y.csv", p8, delimiter=",", header="A,C,G,T,a,c,g,t")
    with pytest.raises(NotImplementedError):
        ingest.load_reads([str(tmp_path / "y.csv")], "poreover")
    ms = ingest.load_models([str(tmp_path / "x.csv"), str(tmp_path / "y.csv")], "")
    assert [m.kind for m in ms] == ["poreover", "flipflop"]


def test_read_npy_equals_np_load(tmp_path):
    rng = np.random.default_rng(7)
    cases = [rng.random((7, 5)).astype(np.float32), rng.random((4, 3, 5)), rng.integers(0, 255, (9, 8)).astype(np.uint8),
             np.zeros((0, 5), np.float32), rng.random(6).astype(np.float32), np.float32(3.5),
             np.asfortranarray(rng.random((4, 5))), rng.random((3, 5)).astype(">f4"),
             np.zeros(3, dtype=[("a", "<f4"), ("b", "<i4")])]
    for i, a in enumerate(cases):
        f = tmp_path / ("c%d.npy" % i)
        np.save(f, a)
        got, want = ingest.read_npy(str(f)), np.load(f)
        assert got.dtype == want.dtype and got.shape == want.shape and np.array_equal(got, want), i
    # numpy's other header versions
    f = tmp_path / "v2.npy"
    with open(f, "wb") as fh:
        np.lib.format.write_array(fh, cases[0], version=(2, 0))
    assert np.array_equal(ingest.read_npy(str(f)), cases[0])
    # a truncated payload is left to np.load (which raises)
    raw = open(tmp_path / "c0.npy", "rb").read()
    open(tmp_path / "short.npy", "wb").write(raw[:-8])
    with pytest.raises(Exception):
        ingest.read_npy(str(tmp_path / "short.npy"))


def test_probability_row_rule_is_np_isclose():
    for s in (1.0, 1 + 4e-6, 1 - 9.9e-6, 1 + 1.0005e-5, 1 + 1.002e-5, 1 + 3e-5, 0.0, -7.25, np.nan, np.inf):
        row = np.array([s, 0, 0, 0, 0], dtype=np.float64)
        assert ingest._is_probability_row(row) == bool(np.isclose(np.sum(row), 1)), s
        row32 = row.astype(np.float32)
        assert ingest._is_probability_row(row32) == bool(np.isclose(np.sum(row32), 1)), s


def test_lookahead_overlaps_and_keeps_order():
    chunks = iter(range(6))
    log, main = [], threading.get_ident()
    src_threads = set()

    def source():
        src_threads.add(threading.get_ident())
        return next(chunks, None)

    def load(c):
        log.append(("load", c))
        assert threading.get_ident() != main
        time.sleep(0.01)
        return c * 10

    seen = []
    for c, payload in ingest.Lookahead(source, load):
        log.append(("use", c))
        seen.append((c, payload))
        time.sleep(0.02)
    assert seen == [(c, c * 10) for c in range(6)]
    assert src_threads == {main}  # the (possibly distributed) queue is only touched by the caller
    for c in range(5):
        assert log.index(("load", c + 1)) < log.index(("use", c + 1))
        assert log.index(("load", c + 1)) > log.index(("load", c))
    # chunk k+1 is being loaded while chunk k is in use
    assert log.index(("load", 2)) < log.index(("use", 2))

    def bad(c):
        raise RuntimeError("loader failed on %d" % c)

    chunks2 = iter(range(3))
    with pytest.raises(RuntimeError, match="loader failed on 0"):
        for _ in ingest.Lookahead(lambda: next(chunks2, None), bad):
            pass
    assert list(ingest.Lookahead(lambda: None, load)) == []


def test_three_stage_pipeline():
    """load (thread A) -> work (thread B, the GPU call) -> the caller, all three busy at once, results in order."""
    chunks = iter(range(8))
    busy, peak, lock = set(), [0], threading.Lock()
    threads = {"load": set(), "work": set()}

    def stage(name, dt):
        def f(x):
            threads[name].add(threading.get_ident())
            with lock:
                busy.add(name)
                peak[0] = max(peak[0], len(busy))
            time.sleep(dt)
            with lock:
                busy.discard(name)
            return x + (name,) if isinstance(x, tuple) else (x, name)
        return f

    out = []
    for c, r in ingest.Lookahead(lambda: next(chunks, None), stage("load", 0.01), stage("work", 0.02)):
        with lock:
            busy.add("use")
            peak[0] = max(peak[0], len(busy))
        time.sleep(0.02)
        with lock:
            busy.discard("use")
        out.append((c, r))
    assert out == [(c, (c, "load", "work")) for c in range(8)]
    assert peak[0] == 3
    assert len(threads["load"]) == 1 and 1 <= len(threads["work"]) <= 2 and not (threads["load"] & threads["work"])
    # two GPU workers, uneven durations: results still come back in chunk order
    chunks3 = iter(range(9))
    got = list(ingest.Lookahead(lambda: next(chunks3, None), lambda c: c,
                                lambda c: (time.sleep(0.02 if c % 2 == 0 else 0.001), c * c)[1], workers=2))
    assert got == [(c, c * c) for c in range(9)]
    chunks4 = iter(range(4))
    one = list(ingest.Lookahead(lambda: next(chunks4, None), lambda c: c, lambda c: -c, workers=1))
    assert one == [(c, -c) for c in range(4)]

    def bad(x):
        raise ValueError("gpu stage failed")

    chunks2 = iter(range(4))
    with pytest.raises(ValueError, match="gpu stage failed"):
        for _ in ingest.Lookahead(lambda: next(chunks2, None), stage("load", 0.0), bad):
            pass


def test_ramped_queue_and_pinned_fallback(monkeypatch):
    from poreover_b200 import multigpu
    q = multigpu.WorkQueue(1000, 256, ramp=2)
    got = list(iter(q.next, None))
    assert got == [(0, 64), (64, 128), (128, 384), (384, 640), (640, 896), (896, 1000)]
    assert list(iter(multigpu.WorkQueue(10, 4).next, None)) == [(0, 4), (4, 8), (8, 10)]
    assert list(iter(multigpu.WorkQueue(0, 4, ramp=1).next, None)) == []
    # chunks are also cut by weight (file bytes): long reads make small batches, a single heavy item still goes through
    q = multigpu.WorkQueue(10, 4, weights=[5, 5, 5, 1, 1, 1, 1, 1, 1, 1], weight_budget=6)
    assert list(iter(q.next, None)) == [(0, 1), (1, 2), (2, 4), (4, 8), (8, 10)]
    q = multigpu.WorkQueue(10, 4, ramp=1, weights=[50, 5, 5, 1, 1, 1, 1, 1, 1, 1], weight_budget=8)
    assert list(iter(q.next, None)) == [(0, 1), (1, 2), (2, 6), (6, 10)]
    assert multigpu._lanes_for([100, 3 << 20]) == 1 and multigpu._lanes_for([100, 200000]) is None
    out = multigpu.run_sharded(list(range(40)), [i % 5 for i in range(40)], lambda p: [x * 2 for x in p], chunk=8,
                               load_chunk=lambda sub: [x + 1 for x in sub], finish_chunk=lambda r: [x - 2 for x in r])
    assert out == [2 * i for i in range(40)]
    assert ingest.packed_alloc() is None
    monkeypatch.setenv("POREOVER_B200_PINNED", "1")
    alloc = ingest.packed_alloc()  # no GPU here: cudaMallocHost fails and numpy's memory is used
    a = alloc((16, 5), np.float32)
    assert a.shape == (16, 5) and a.dtype == np.float32
    assert alloc((0, 5), np.float32).shape == (0, 5)


def test_transducer_builds_float64_and_transition_on_demand():
    x = np.log(np.random.default_rng(8).random((10, 5)).astype(np.float32))
    m = transducer.bonito(x)
    assert m._lp64 is None and m._transition is None  # nothing allocated until somebody asks (transducer.py:16, :22)
    assert m.device_array() is m._f32
    assert m.log_prob.dtype == np.float64 and np.array_equal(m.log_prob, x.astype(np.float64))
    assert m.transition.shape == (10, 5) and (m.transition == 1).all()
    assert np.array_equal(m[3], m.log_prob[3]) and m.t_max == 10 and m.num_states == 5
    m.reverse_complement()
    assert np.array_equal(m.log_prob, x.astype(np.float64)[::-1, [3, 2, 1, 0, 4]])
    assert np.array_equal(m.device_array(), x[::-1, [3, 2, 1, 0, 4]])
    m.log_prob = np.zeros((3, 5))  # an assigned table replaces the float32 original
    assert m.device_array().dtype == np.float64 and m.device_array().shape == (3, 5)
    f = transducer.flipflop(np.zeros((4, 8)))
    assert f.transition.shape == (8, 8) and f.log_prob.dtype == np.float64
    with pytest.raises(AssertionError):
        transducer.bonito(np.zeros((4, 6), np.float32))


def test_load_pairs_host_stage(tmp_path):
    from argparse import Namespace
    from poreover_b200.decoding import pair_decode as gpd
    pairs = [list(synth.save_pair(str(tmp_path), k, 150 + 10 * k)) for k in range(3)]
    args = Namespace(dir=str(tmp_path), basecaller="bonito", reverse_complement=True, alignment="banded",
                     skip_matches=False, diagonal_envelope=False, single="viterbi")
    meta, b1, b2, kind, keep, total = gpd.load_pairs(args, pairs)
    assert keep == [0, 1, 2] and total == 3
    assert kind == "bonito" and b1.n == b2.n == 3 and b1.rc is None and b2.rc.tolist() == [1, 1, 1]
    assert b1.layout == b2.layout == _lib.BLANK_FIRST
    assert [m[0] for m in meta] == pairs and meta[0][3] == "bonito"
    args.skip_matches = True  # staged flags: per-read arrays in the reference's column order
    meta, m1, m2, kind, keep, total = gpd.load_pairs(args, pairs)
    ref = _reference_arrays([os.path.join(str(tmp_path), p[0]) for p in pairs], "bonito")
    assert all(np.array_equal(a, b) for a, b in zip(m1, ref)) and len(m2) == 3
    # a pair whose file is missing or corrupt is dropped alone (the reference's pool task dies alone, pair_decode.py:295)
    args.skip_matches = False
    (tmp_path / "broken_1.npy").write_bytes(b"not an npy file")
    bad = pairs[:1] + [["broken_1.npy", pairs[1][1]], ["missing_1.npy", "missing_2.npy"]] + pairs[1:]
    meta, b1, b2, kind, keep, total = gpd.load_pairs(args, bad)
    assert keep == [0, 3, 4] and total == 5 and b1.n == 3 and [m[0] for m in meta] == pairs
    assert gpd._scatter(["a", "b", "c"], keep, total) == ["a", None, None, "b", "c"]
    meta, b1, b2, kind, keep, total = gpd.load_pairs(args, [["missing_1.npy", "missing_2.npy"]])
    assert keep == [] and total == 1 and len(meta) == 0


def test_real_logits_fixture_through_the_batched_loader(tmp_path):
    """The reference's own real-data pair (PoreOverNet logits, windows x time x states, 62,000 and 75,600 timesteps):
    the batched loader must hand the kernels exactly the arrays decode.model_from_trace makes."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "real_pair.npz"))
    paths = []
    for name in ("read1", "read2"):
        f = tmp_path / (name + ".npy")
        np.save(f, g[name])
        paths.append(str(f))
    ref = _reference_arrays(paths, "poreover")
    b = ingest.load_reads(paths, "poreover", rc=[0, 1])
    assert b.layout == _lib.BLANK_LAST and b.dtype == _lib.F32 and b.lens.tolist() == [62000, 75600]
    for i, a in enumerate(ref):
        assert np.array_equal(_rows(b, i), a)


def test_work_queue_bounds_partition_the_items():
    """Whatever the chunk size, ramp and byte budget: the chunks are non-empty, contiguous and cover every item once."""
    from poreover_b200 import multigpu
    rng = np.random.default_rng(12)
    for _ in range(200):
        n = int(rng.integers(0, 300))
        chunk = int(rng.integers(1, 64))
        ramp = int(rng.integers(0, 4))
        weights = sorted((int(x) for x in rng.integers(0, 50, size=n)), reverse=True) if rng.random() < 0.7 else None
        budget = int(rng.integers(1, 200)) if weights is not None else None
        pullers = int(rng.integers(1, 17)) if rng.random() < 0.5 else 0  # tapering towards the end of the queue
        min_chunk = int(rng.integers(1, 32))
        q = multigpu.WorkQueue(n, chunk, ramp=ramp, weights=weights, weight_budget=budget, pullers=pullers,
                               min_chunk=min_chunk)
        got = list(iter(q.next, None))
        assert all(lo < hi for lo, hi in got)
        assert [lo for lo, _ in got] == [0] * bool(got) + [hi for _, hi in got[:-1]]
        assert (got[-1][1] if got else 0) == n
        assert all(hi - lo <= chunk for lo, hi in got[:-1])
        assert not got or got[-1][1] - got[-1][0] <= chunk + (min_chunk // 2 if pullers else 0)
        if pullers and weights is None and ramp == 0:  # guided: sizes never grow (but for the crumb merged at the end)
            sizes = [hi - lo for lo, hi in got[:-1]]
            assert sizes == sorted(sizes, reverse=True)
        if weights is not None and not pullers:
            for lo, hi in got:  # within the budget, except for a single item that alone exceeds it
                assert hi - lo == 1 or sum(weights[lo:hi]) <= budget


def test_fasta_format_is_the_reference_loop():
    """decode.py:20-27 literally, against the join-based version, for every length around the line width."""
    def ref(name, seq, width=60):
        fasta = '>' + name + '\n'
        window = 0
        while window + width < len(seq):
            fasta += (seq[window:window + width] + '\n')
            window += width
        fasta += (seq[window:] + '\n')
        return fasta
    for n in list(range(0, 200)) + [599, 600, 601, 6000]:
        s = ''.join("ACGT"[(7 * i) % 4] for i in range(n))
        for w in (60, 1, 7):
            assert gdecode.fasta_format("read", s, w) == ref("read", s, w), (n, w)
