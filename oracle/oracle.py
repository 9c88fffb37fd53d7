"""TEST INFRASTRUCTURE ONLY -- ctypes front end of the parity oracle.

Two back ends:
  * ``port``  : oracle/poreover_oracle.c, the plain-C restatement (always available, built by
                ``make -C oracle``).
  * ``ref``   : oracle/_ref/, the UNMODIFIED reference search core / Cython aligner compiled from
                /root/reference in the build container (``make -C oracle ref``); optional.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product (poreover_b200/) never does.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_PORT = os.path.join(HERE, "libporeover_oracle.so")
_REF = os.path.join(HERE, "_ref", "libporeover_ref.so")

MODELS = {"ctc": 0, "ctc_merge_repeats": 1}
KINDS = {"poreover": 0, "bonito": 1, "flipflop": 2}
BASES = np.frombuffer(b"ACGT", dtype=np.uint8)

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)
c_lp = C.POINTER(C.c_int64)
c_bp = C.POINTER(C.c_uint8)


def build(ref=True):
    """Compile the oracle (and oracle/_ref when the reference tree is present)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "all"])
    if ref and os.path.isdir("/root/reference/poreover"):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


_port = None
_ref = None


def port():
    global _port
    if _port is None:
        if not os.path.exists(_PORT):
            build(ref=False)
        _port = C.CDLL(_PORT)
        _port.orc_forward.restype = C.c_double
        _port.orc_logaddexp.restype = C.c_double
        _port.orc_logaddexp.argtypes = [C.c_double, C.c_double]
    return _port


def have_ref():
    return os.path.exists(_REF)


def ref():
    global _ref
    if _ref is None:
        if not have_ref():
            raise RuntimeError("oracle/_ref not built (needs /root/reference: make -C oracle ref)")
        _ref = C.CDLL(_REF)
        _ref.ref_forward.restype = C.c_double
        _ref.ref_trace_len.restype = C.c_long
    return _ref


def ref_align_module():
    """The reference's own Cython aligner (align.pyx) compiled into oracle/_ref/."""
    d = os.path.join(HERE, "_ref")
    if d not in sys.path:
        sys.path.insert(0, d)
    import align  # noqa

    return align


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _p(a, t):
    return a.ctypes.data_as(t)


def bases_to_str(idx):
    return BASES[np.asarray(idx, dtype=np.int64)].tobytes().decode()


# ---------------------------------------------------------------- viterbi / mapping
def viterbi(log_prob, kind):
    """(sequence, path) as transducer.{poreover,bonito,flipflop}.viterbi_decode(return_path=True)."""
    y = _f64(log_prob)
    T, S = y.shape
    path = np.zeros(T, dtype=np.int64)
    seq = np.zeros(max(T, 1), dtype=np.uint8)
    if kind == "flipflop":
        L = port().orc_viterbi_flipflop(_p(y, c_dp), T, _p(path, c_lp), _p(seq, c_bp))
    else:
        L = port().orc_viterbi_argmax(_p(y, c_dp), T, S, KINDS[kind], _p(path, c_lp), _p(seq, c_bp))
    return bases_to_str(seq[:L]), path


def sequence_mapping(path, kind):
    path = np.ascontiguousarray(path, dtype=np.int64)
    s2s = np.zeros(max(len(path), 1), dtype=np.int64)
    L = port().orc_sequence_mapping(_p(path, c_lp), len(path), KINDS[kind], _p(s2s, c_lp))
    return s2s[:L].copy()


def reverse_complement(log_prob, kind):
    """transducer.py:68-70, :79-81, :104-106"""
    y = np.asarray(log_prob)
    perm = [3, 2, 1, 0, 7, 6, 5, 4] if kind == "flipflop" else [3, 2, 1, 0, 4]
    return y[::-1, perm]


# ---------------------------------------------------------------- alignment / envelope
def global_pair_banded(seq1, seq2, band_width=500, match=2, mismatch=-1, gap_cost=-1):
    s1, s2 = seq1.encode(), seq2.encode()
    if len(s1) == 0:
        raise ZeroDivisionError("float division by zero")
    cap = len(s1) + len(s2) + 16
    a1 = C.create_string_buffer(cap)
    a2 = C.create_string_buffer(cap)
    n = port().orc_global_pair_banded(s1, len(s1), s2, len(s2), band_width, match, mismatch, gap_cost, a1, a2)
    return list(a1.raw[:n].decode()), list(a2.raw[:n].decode())


def global_pair(seq1, seq2, match=2, mismatch=-1, gap_cost=-1):
    s1, s2 = seq1.encode(), seq2.encode()
    cap = len(s1) + len(s2) + 16
    a1 = C.create_string_buffer(cap)
    a2 = C.create_string_buffer(cap)
    dp = np.zeros((len(s1) + 1, len(s2) + 1), dtype=np.int32)
    n = port().orc_global_pair(s1, len(s1), s2, len(s2), match, mismatch, gap_cost, a1, a2, _p(dp, c_ip))
    return list(a1.raw[:n].decode()), list(a2.raw[:n].decode()), dp


def alignment_columns(alignment):
    """alignment: 2 x C array/list of single characters. Returns list of (label, x_index, y_index)."""
    a1 = "".join(alignment[0]).encode()
    a2 = "".join(alignment[1]).encode()
    n = len(a1)
    lab = np.zeros(max(n, 1), dtype=np.int8)
    xi = np.zeros(max(n, 1), dtype=np.int32)
    yi = np.zeros(max(n, 1), dtype=np.int32)
    port().orc_alignment_columns(a1, a2, n, lab.ctypes.data_as(C.POINTER(C.c_int8)), _p(xi, c_ip), _p(yi, c_ip))
    return [("mid"[lab[c]], int(xi[c]), int(yi[c])) for c in range(n)]


def build_envelope(U, V, alignment_col, s2s1, s2s2, padding=150):
    n = len(alignment_col)
    xi = np.array([c[1] for c in alignment_col], dtype=np.int32).reshape(-1)
    yi = np.array([c[2] for c in alignment_col], dtype=np.int32).reshape(-1)
    a = np.ascontiguousarray(s2s1, dtype=np.int64)
    b = np.ascontiguousarray(s2s2, dtype=np.int64)
    env = np.zeros((U, 2), dtype=np.int64)
    port().orc_build_envelope(U, V, n, _p(xi, c_ip), _p(yi, c_ip), _p(a, c_lp), len(a), _p(b, c_lp), len(b),
                              padding, _p(env, c_lp))
    return env


# ---------------------------------------------------------------- beam searches
def _out(cap):
    return C.create_string_buffer(cap)


def beam_search(y, beam_width=25, model="ctc", backend="port", with_score=False):
    y = _f64(y)
    T, S = y.shape
    out = _out(T + 8)
    sc = C.c_double(0)
    if backend == "port":
        n = port().orc_beam_search(_p(y, c_dp), T, S, beam_width, MODELS[model], out, T + 8, C.byref(sc))
    elif backend == "ref":
        n = ref().ref_beam_search(_p(y, c_dp), T, S, beam_width, model.encode(), out, T + 8, C.byref(sc))
    else:  # "stock": the unmodified dispatch function, string only
        n = ref().ref_beam_search_stock(_p(y, c_dp), T, S, beam_width, model.encode(), out, T + 8)
    assert n >= 0, n
    s = out.raw[:n].decode()
    return (s, sc.value) if with_score else s


def beam_search_2d(y1, y2, envelope=None, beam_width=25, model="ctc", method="row", backend="port",
                   with_score=False, info=None):
    y1, y2 = _f64(y1), _f64(y2)
    U, S = y1.shape
    V = y2.shape[0]
    env = None if envelope is None else np.ascontiguousarray(np.asarray(envelope), dtype=np.int32)
    envp = None if env is None else _p(env, c_ip)
    cap = U + V + 8
    out = _out(cap)
    sc = C.c_double(0)
    if backend == "port":
        if method == "row":
            n = port().orc_beam_search_2d_row(_p(y1, c_dp), _p(y2, c_dp), U, V, S, envp, beam_width,
                                              MODELS[model], out, cap, C.byref(sc))
        elif method == "row_col":
            st = C.c_int(0)
            nu = C.c_int64(0)
            n = port().orc_beam_search_2d_row_col(_p(y1, c_dp), _p(y2, c_dp), U, V, S, envp, beam_width,
                                                  MODELS[model], out, cap, C.byref(sc), C.byref(st), C.byref(nu))
            if info is not None:
                info["status"] = st.value
                info["n_updates"] = nu.value
        else:
            raise ValueError(method)
    elif backend == "ref":
        n = ref().ref_beam_search_2d(_p(y1, c_dp), _p(y2, c_dp), U, V, S, envp, beam_width, model.encode(),
                                     method.encode(), out, cap, C.byref(sc))
    else:
        n = ref().ref_beam_search_2d_stock(_p(y1, c_dp), _p(y2, c_dp), U, V, S, envp, beam_width, model.encode(),
                                           method.encode(), out, cap)
    assert n >= 0, n
    s = out.raw[:n].decode()
    return (s, sc.value) if with_score else s


def forward(y, label, model="ctc", backend="port"):
    y = _f64(y)
    T, S = y.shape
    if backend == "port":
        lab = np.array(["ACGT".index(c) for c in label], dtype=np.uint8)
        return port().orc_forward(_p(y, c_dp), T, S, _p(lab, c_bp), len(lab), MODELS[model])
    return ref().ref_forward(_p(y, c_dp), T, S, label.encode(), model.encode())


def viterbi_acceptor(y, label, band_size=1000, backend="port"):
    """decoding_cpp.cpp_viterbi_acceptor (decoding_cpp.pyx:69-84): int array of length T, 4 = gap."""
    y = _f64(y)
    T, S = y.shape
    path = np.zeros(T, dtype=np.int8)
    if backend == "port":
        lab = np.array(["ACGT".index(c) for c in label], dtype=np.uint8)
        rc = port().orc_viterbi_acceptor(_p(y, c_dp), T, S, int(band_size), _p(lab, c_bp), len(lab),
                                         path.ctypes.data_as(C.POINTER(C.c_int8)))
        if rc != 0:
            raise RuntimeError("traceback leaves the matrix (the reference does not terminate here)")
    else:
        ref().ref_viterbi_acceptor(_p(y, c_dp), T, S, int(band_size), label.encode(),
                                   path.ctypes.data_as(C.POINTER(C.c_int8)))
    return path.astype(np.int64)


# ---------------------------------------------------------------- full pair path (pair_decode.py:305-531)
def pair_decode(lp1, lp2, kind="bonito", beam_width=25, padding=5, method="row_col", backend="port",
                band_width=500, with_score=False):
    """lp1 / lp2: log-prob arrays as the reference loader produced them; read 2 already reverse-complemented.
    Returns dict(basecall1, basecall2, identity, skipped, consensus[, score], envelope)."""
    model = {"poreover": "ctc", "bonito": "ctc_merge_repeats"}[kind]
    b1, p1 = viterbi(lp1, kind)
    b2, p2 = viterbi(lp2, kind)
    res = {"basecall1": b1, "basecall2": b2, "skipped": 0}
    if abs(len(b1) - len(b2)) > 1000:
        res["skipped"] = 1
        return res
    m1 = sequence_mapping(p1, kind)
    m2 = sequence_mapping(p2, kind)
    assert len(m1) == len(b1) and len(m2) == len(b2)
    if backend == "ref":
        a = ref_align_module().global_pair_banded(b1, b2)
    else:
        a = global_pair_banded(b1, b2, band_width)
    al = np.array([list(s) for s in a[:2]])
    ident = np.sum(al[0] == al[1]) / len(al[0])
    res["identity"] = ident
    if ident < 0.5:
        res["skipped"] = 1
        return res
    cols = alignment_columns(al)
    env = build_envelope(len(lp1), len(lp2), cols, m1, m2, padding)
    res["envelope"] = env
    r = beam_search_2d(lp1, lp2, env, beam_width, model, method, backend=backend, with_score=with_score)
    if with_score:
        res["consensus"], res["score"] = r
    else:
        res["consensus"] = r
    return res
