"""Batched host-side API over the C ABI: packing of reads and the calls the drivers actually make.

The single-item functions in poreover_b200.decoding / poreover_b200.align (which keep the reference's
signatures) are thin wrappers over these with a batch of one.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import ReadsT, check, get_ctx, lib, ptr

ALIGN_ROWS = 4  # float32 x 5 states: 4 rows = 80 B keeps every read 16-byte aligned

import threading  # noqa: E402

_tls = threading.local()


def pinned_scratch(tag, shape, dtype):
    """A zeroed numpy array over PINNED host memory (cudaMallocHost), owned by the calling thread and reused by its
    next call with the same tag (grow-only).  Result buffers of the batched calls live here: a device-to-host copy
    into pageable memory is synchronous and, with two GPU calls in flight per process, serialises them (measured:
    3.9k against 5.3k pairs/s end to end at 313-pair chunks).  The caller must be done with the array (results are
    turned into Python strings / copied) before its next call with that tag."""
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    cache = _tls.__dict__.setdefault("bufs", {})
    have = cache.get(tag)
    if have is None or have[1] < nbytes:
        if have is not None:
            lib().pob_free_host(_lib.vp(have[0]))
        cap = max(4096, int(nbytes * 1.25))
        ptr_ = _lib.vp()
        check(lib().pob_malloc_host(cap, C.byref(ptr_)), "pob_malloc_host(%d)" % cap)
        have = cache[tag] = (ptr_.value, cap)
    buf = (C.c_char * max(nbytes, 1)).from_address(have[0])
    a = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    a[...] = 0
    return a


class ReadBatch:
    """Packed log-probability matrices (host side). Keeps the numpy arrays alive for the C call."""

    def __init__(self, arrays, rc=None, layout=_lib.BLANK_LAST, dtype=None):
        arrays = [np.asarray(a) for a in arrays]
        n = len(arrays)
        if dtype is None:
            dtype = np.float32 if all(a.dtype == np.float32 for a in arrays) and n else np.float64
        self.np_dtype = np.dtype(dtype)
        self.dtype = {np.dtype(np.float32): _lib.F32, np.dtype(np.float64): _lib.F64, np.dtype(np.uint8): _lib.U8_TRACE}[self.np_dtype]
        self.n = n
        self.n_states = int(arrays[0].shape[1]) if n else 5
        self.layout = layout
        self.lens = np.array([a.shape[0] for a in arrays], dtype=np.int32)
        padded = (self.lens.astype(np.int64) + ALIGN_ROWS - 1) // ALIGN_ROWS * ALIGN_ROWS
        self.row_off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(padded, out=self.row_off[1:])
        total = int(self.row_off[-1])
        self.data = np.zeros((total, self.n_states), dtype=self.np_dtype)
        for a, o in zip(arrays, self.row_off[:-1]):
            if a.shape[1] != self.n_states:
                raise ValueError("all reads of a batch must have the same number of states")
            self.data[o:o + a.shape[0]] = a
        self.rc = None if rc is None else np.ascontiguousarray(np.broadcast_to(np.asarray(rc, dtype=np.uint8), (n,)))
        self.total_rows = total

    def struct(self):
        return ReadsT(self.data.ctypes.data, self.row_off.ctypes.data, self.lens.ctypes.data,
                      None if self.rc is None else self.rc.ctypes.data, self.n, self.n_states, self.dtype, self.layout)


def _unpack(buf, offs, lens):
    return [buf[o:o + l] for o, l in zip(offs, lens)]


def viterbi_batch(arrays, kind, rc=None, layout=_lib.BLANK_LAST, return_path=False, device=None, return_maps=True):
    """Best-path decode of many reads.  Returns (sequences, s2s lists or None, paths or None, status array).

    replaces transducer.{poreover,bonito}.viterbi_decode + get_sequence_mapping (see include/poreover_b200.h).
    return_maps=False (the `decode` command line only wants the sequences) skips the mapping's device-to-host copy
    and its per-read arrays."""
    b = arrays if isinstance(arrays, ReadBatch) else ReadBatch(arrays, rc=rc, layout=layout)
    ctx = get_ctx(device)
    rows = max(b.total_rows, 1)
    seq = np.zeros(rows, dtype=np.uint8)
    s2s = np.zeros(rows, dtype=np.int32) if return_maps else None
    path = np.zeros(rows, dtype=np.int8) if return_path else None
    ln = np.zeros(max(b.n, 1), dtype=np.int32)
    st = np.zeros(max(b.n, 1), dtype=np.int32)
    rs = b.struct()
    check(lib().pob_viterbi(ctx.h, _lib.HOST, C.byref(rs), _lib.KIND[kind], ptr(seq), ptr(s2s), ptr(path), ptr(ln),
                            ptr(st)), "pob_viterbi")
    offs = b.row_off[:-1].tolist()
    lens = ln[:b.n].tolist()
    text = seq.tobytes().decode("latin1")  # one decode, then string slices (offsets are byte offsets)
    seqs = [text[o:o + l] for o, l in zip(offs, lens)]
    maps = [s2s[o:o + l].astype(np.int64) for o, l in zip(offs, lens)] if return_maps else None
    paths = [path[o:o + t].astype(np.int64) for o, t in zip(offs, b.lens)] if return_path else None
    return seqs, maps, paths, st[:b.n].copy()


def _pack_bytes(strings):
    bs = [s.encode() if isinstance(s, str) else bytes(s) for s in strings]
    off = np.zeros(len(bs) + 1, dtype=np.int64)
    np.cumsum([len(b) for b in bs], out=off[1:])
    buf = np.frombuffer(b"".join(bs) + b"\0", dtype=np.uint8).copy()
    return buf, off


def align_banded_batch(seqs1, seqs2, band_width=500, match=2, mismatch=-1, gap_cost=-1, device=None):
    """Banded NW for many sequence pairs.  Returns list of (row1, row2, matches) with row strings.

    replaces align.global_pair_banded (align.pyx:100-178)."""
    n = len(seqs1)
    # align.pyx:120-171 with l1 == 0: the row loop never runs (no division happens), the traceback only drains seq2,
    # so the alignment is all gaps against seq2; those pairs never reach the kernel
    empty = [k for k in range(n) if len(seqs1[k]) == 0 or len(seqs2[k]) == 0]
    if empty:
        keep = [k for k in range(n) if k not in set(empty)]
        rest = align_banded_batch([seqs1[k] for k in keep], [seqs2[k] for k in keep], band_width, match, mismatch,
                                  gap_cost, device) if keep else []
        out = [None] * n
        for k, r in zip(keep, rest):
            out[k] = r
        text = lambda s: s if isinstance(s, str) else bytes(s).decode()  # noqa: E731
        for k in empty:
            s1, s2 = text(seqs1[k]), text(seqs2[k])
            out[k] = ('-' * len(s2), s2, 0) if not s1 else (s1, '-' * len(s1), 0)
        return out
    ctx = get_ctx(device)
    b1, o1 = _pack_bytes(seqs1)
    b2, o2 = _pack_bytes(seqs2)
    aln_off = o1 + o2 + 8 * np.arange(n + 1, dtype=np.int64)
    a1 = np.zeros(int(aln_off[-1]) + 1, dtype=np.uint8)
    a2 = np.zeros(int(aln_off[-1]) + 1, dtype=np.uint8)
    alen = np.zeros(max(n, 1), dtype=np.int32)
    mat = np.zeros(max(n, 1), dtype=np.int32)
    check(lib().pob_align_banded(ctx.h, _lib.HOST, ptr(b1), ptr(o1), ptr(b2), ptr(o2), n, band_width, match, mismatch,
                                 gap_cost, ptr(a1), ptr(a2), ptr(alen), ptr(mat)), "pob_align_banded")
    out = []
    for p in range(n):
        o, l = int(aln_off[p]), int(alen[p])
        out.append((a1[o:o + l].tobytes().decode(), a2[o:o + l].tobytes().decode(), int(mat[p])))
    return out


def build_envelope_batch(alignments, s2s1_list, s2s2_list, U_list, V_list, padding=5, device=None):
    """alignments: list of (row1, row2) gapped strings.  Returns list of int64 (U,2) envelopes.

    replaces envelope.get_alignment_columns + envelope.build_envelope (envelope.py:26-87)."""
    n = len(alignments)
    ctx = get_ctx(device)
    r1, aoff = _pack_bytes([a[0] for a in alignments])
    r2, _ = _pack_bytes([a[1] for a in alignments])
    alen = np.diff(aoff).astype(np.int32)

    def pack_i32(lst):
        off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum([len(x) for x in lst], out=off[1:])
        buf = np.zeros(int(off[-1]) + 1, dtype=np.int32)
        for x, o in zip(lst, off[:-1]):
            buf[o:o + len(x)] = np.asarray(x, dtype=np.int32)
        return buf, off, np.diff(off).astype(np.int32)

    m1, so1, l1 = pack_i32(s2s1_list)
    m2, so2, l2 = pack_i32(s2s2_list)
    U = np.asarray(U_list, dtype=np.int32)
    V = np.asarray(V_list, dtype=np.int32)
    eoff = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(U, out=eoff[1:])
    env = np.zeros((int(eoff[-1]) + 1, 2), dtype=np.int32)
    check(lib().pob_build_envelope(ctx.h, _lib.HOST, ptr(r1), ptr(r2), ptr(aoff), ptr(alen), ptr(m1), ptr(so1),
                                   ptr(l1), ptr(m2), ptr(so2), ptr(l2), ptr(U), ptr(V), ptr(eoff), n, padding,
                                   ptr(env)), "pob_build_envelope")
    return [env[eoff[p]:eoff[p + 1]].astype(np.int64) for p in range(n)]


def _as_batch(arrays, rc=None, layout=_lib.BLANK_LAST):
    return arrays if isinstance(arrays, ReadBatch) else ReadBatch(arrays, rc=rc, layout=layout)


def beam_search_batch(arrays, beam_width=25, model="ctc", rc=None, layout=_lib.BLANK_LAST, device=None):
    """CTC prefix beam search over many single reads.  Returns (sequences, scores, status).

    replaces decoding_cpp.cpp_beam_search (decoding_cpp.pyx:88-103)."""
    if model not in _lib.MODEL:
        raise ValueError("unknown model %r (the reference falls off the end of beam_search, BeamSearch.h:400-408)" % model)
    b = _as_batch(arrays, rc, layout)
    ctx = get_ctx(device)
    n = b.n
    seq = np.zeros(max(b.total_rows, 1) + 4, dtype=np.uint8)
    ln = np.zeros(max(n, 1), dtype=np.int32)
    sc = np.zeros(max(n, 1), dtype=np.float64)
    st = np.zeros(max(n, 1), dtype=np.int32)
    rs = b.struct()
    check(lib().pob_beam_search(ctx.h, _lib.HOST, C.byref(rs), int(beam_width), _lib.MODEL[model], ptr(seq), ptr(ln),
                                ptr(sc), ptr(st)), "pob_beam_search")
    seqs = [seq[o:o + l].tobytes().decode() for o, l in zip(b.row_off[:-1], ln[:n])]
    return seqs, sc[:n].copy(), st[:n].copy()


def beam_search_2d_batch(arrays1, arrays2, envelopes=None, beam_width=25, model="ctc", method="row", rc1=None,
                         rc2=None, layout=_lib.BLANK_LAST, device=None):
    """Joint two-read prefix beam search.  envelopes: list of (U,2) int arrays or None.

    replaces decoding_cpp.cpp_beam_search_2d (decoding_cpp.pyx:107-139).  Returns (sequences, scores, status)."""
    if model not in _lib.MODEL:
        raise ValueError("unknown model %r" % model)
    if method not in _lib.METHOD:
        raise ValueError("unknown method %r (only 'row' and 'row_col' are built; 'grid' is out of scope)" % method)
    b1 = _as_batch(arrays1, rc1, layout)
    b2 = _as_batch(arrays2, rc2, layout)
    n = b1.n
    if envelopes is None and method == "row_col":
        raise ValueError("method 'row_col' needs an envelope (the reference dereferences a null envelope)")
    ctx = get_ctx(device)
    env = env_off = None
    if envelopes is not None:
        env_off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(b1.lens, out=env_off[1:])
        env = np.zeros((int(env_off[-1]) + 1, 2), dtype=np.int32)
        for p, e in enumerate(envelopes):
            e = np.asarray(e, dtype=np.int32).reshape(-1, 2)  # decoding_cpp.pyx:121: np.intc
            if e.shape[0] < b1.lens[p]:
                raise ValueError("envelope of pair %d has %d rows, read 1 has %d timesteps" % (p, e.shape[0], b1.lens[p]))
            env[env_off[p]:env_off[p + 1]] = e[:b1.lens[p]]
    out_off = (b1.row_off + b2.row_off).astype(np.int64)
    seq = np.zeros(int(out_off[-1]) + 8, dtype=np.uint8)
    ln = np.zeros(max(n, 1), dtype=np.int32)
    sc = np.zeros(max(n, 1), dtype=np.float64)
    st = np.zeros(max(n, 1), dtype=np.int32)
    s1, s2 = b1.struct(), b2.struct()
    check(lib().pob_beam_search_2d(ctx.h, _lib.HOST, C.byref(s1), C.byref(s2), ptr(env), ptr(env_off), int(beam_width),
                                   _lib.MODEL[model], _lib.METHOD[method], ptr(out_off), ptr(seq), ptr(ln), ptr(sc),
                                   ptr(st)), "pob_beam_search_2d")
    seqs = [seq[o:o + l].tobytes().decode() for o, l in zip(out_off[:-1], ln[:n])]
    return seqs, sc[:n].copy(), st[:n].copy()


def pair_decode_batch(arrays1, arrays2, kind="bonito", beam_width=25, padding=5, band_width=500, method="row_col",
                      rc2=True, rc1=False, layout=_lib.BLANK_LAST, device=None):
    """The whole pair-decode hot path for many pairs in one call (pair_decode.py:361-398, :495-511).

    arrays: log-probabilities as the loader produced them; rc2=True is --reverse_complement.
    Returns one dict per pair: basecall1, basecall2, consensus, score, identity, skipped, status."""
    b1 = _as_batch(arrays1, rc1 if rc1 is not None else None, layout)
    b2 = _as_batch(arrays2, rc2 if rc2 is not None else None, layout)
    n = b1.n
    ctx = get_ctx(device)
    seq1 = pinned_scratch("pd_seq1", max(b1.total_rows, 1) + 4, np.uint8)
    seq2 = pinned_scratch("pd_seq2", max(b2.total_rows, 1) + 4, np.uint8)
    cons = pinned_scratch("pd_cons", b1.total_rows + b2.total_rows + 8, np.uint8)
    l1 = pinned_scratch("pd_l1", max(n, 1), np.int32)
    l2 = pinned_scratch("pd_l2", max(n, 1), np.int32)
    lc = pinned_scratch("pd_lc", max(n, 1), np.int32)
    sc = pinned_scratch("pd_sc", max(n, 1), np.float64)
    stats = pinned_scratch("pd_stats", (max(n, 1), 4), np.int32)
    st = pinned_scratch("pd_st", max(n, 1), np.int32)
    s1, s2 = b1.struct(), b2.struct()
    check(lib().pob_pair_decode(ctx.h, _lib.HOST, C.byref(s1), C.byref(s2), _lib.KIND[kind], int(beam_width),
                                int(padding), int(band_width), _lib.METHOD[method], ptr(seq1), ptr(l1), ptr(seq2),
                                ptr(l2), ptr(cons), ptr(lc), ptr(sc), ptr(stats), ptr(st)), "pob_pair_decode")
    out = []
    skip_mask = _lib.ST_SKIPPED_LENGTH | _lib.ST_SKIPPED_IDENTITY | _lib.ST_MAPPING_WRAP | _lib.ST_EMPTY
    # one decode per buffer, then string slices (offsets are byte offsets); plain python ints in the loop
    t1, t2, tc = (x.tobytes().decode("latin1") for x in (seq1, seq2, cons))
    off1, off2 = b1.row_off.tolist(), b2.row_off.tolist()
    len1, len2, lenc, stl, scl = l1.tolist(), l2.tolist(), lc.tolist(), st.tolist(), sc.tolist()
    for p in range(n):
        o1, o2, s_ = off1[p], off2[p], stl[p]
        r = {"basecall1": t1[o1:o1 + len1[p]], "basecall2": t2[o2:o2 + len2[p]],
             "status": s_, "skipped": 1 if (s_ & skip_mask) else 0, "length1": len1[p], "length2": len2[p]}
        if stats[p, 3] > 0:
            r["identity"] = stats[p, 2] / stats[p, 3]  # np.sum(a0 == a1) / len(a0)  (pair_decode.py:391)
        if not r["skipped"]:
            r["consensus"] = tc[o1 + o2:o1 + o2 + lenc[p]]
            r["score"] = scl[p]
        out.append(r)
    return out


FLIPFLOP_LUT = None


def flipflop_lut():
    """log((x + 1e-7) / (255 + 1e-7)) for x = 0..255, computed by numpy exactly as decode.py:92-93 does."""
    global FLIPFLOP_LUT
    if FLIPFLOP_LUT is None:
        eps = 0.0000001
        FLIPFLOP_LUT = np.log((np.arange(256, dtype=np.uint8) + eps) / (255 + eps))
    return FLIPFLOP_LUT


def flipflop_viterbi_batch(arrays, rc=None, return_path=False, device=None, return_maps=True):
    """Flip-flop Viterbi over many reads.  arrays: T x 8 float64 log-probabilities, or T x 8 uint8 traces
    (decoded through the host-computed table).  Returns (sequences, s2s lists or None, paths or None).

    replaces transducer.viterbi_decode for kind 'flipflop' (transducer.py:35-59, :94-103)."""
    arrays = [np.asarray(a) for a in arrays]
    u8 = all(a.dtype == np.uint8 for a in arrays) and len(arrays) > 0
    b = ReadBatch(arrays, rc=rc, dtype=np.uint8 if u8 else np.float64)
    if b.n and b.n_states != 8:
        raise ValueError("flip-flop traces have 8 states")
    ctx = get_ctx(device)
    rows = max(b.total_rows, 1)
    seq = np.zeros(rows + 4, dtype=np.uint8)
    s2s = np.zeros(rows + 4, dtype=np.int32) if return_maps else None
    path = np.zeros(rows + 4, dtype=np.int8) if return_path else None
    ln = np.zeros(max(b.n, 1), dtype=np.int32)
    lut = flipflop_lut() if u8 else None
    rs = b.struct()
    check(lib().pob_viterbi_flipflop(ctx.h, _lib.HOST, C.byref(rs), ptr(lut), ptr(seq), ptr(s2s), ptr(path), ptr(ln)),
          "pob_viterbi_flipflop")
    offs = b.row_off[:-1].tolist()
    lens = ln[:b.n].tolist()
    text = seq.tobytes().decode("latin1")  # one decode, then string slices (offsets are byte offsets)
    seqs = [text[o:o + l] for o, l in zip(offs, lens)]
    maps = [s2s[o:o + l].astype(np.int64) for o, l in zip(offs, lens)] if return_maps else None
    paths = [path[o:o + t].astype(np.int64) for o, t in zip(offs, b.lens)] if return_path else None
    return seqs, maps, paths


def forward_batch(arrays, labels, model="ctc", rc=None, layout=_lib.BLANK_LAST, device=None):
    """Exact forward log-probability of one label per read.  labels: strings over ACGT.

    replaces decoding_cpp.cpp_forward (decoding_cpp.pyx:49-65)."""
    if model not in _lib.MODEL:
        raise ValueError("unknown model %r" % model)
    b = _as_batch(arrays, rc, layout)
    ctx = get_ctx(device)
    n = b.n
    lab_off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum([len(l) for l in labels], out=lab_off[1:])
    lab = np.zeros(int(lab_off[-1]) + 8, dtype=np.uint8)
    for l, o in zip(labels, lab_off[:-1]):
        lab[o:o + len(l)] = ["ACGT".index(c) for c in l]
    out = np.zeros(max(n, 1), dtype=np.float64)
    rs = b.struct()
    check(lib().pob_forward(ctx.h, _lib.HOST, C.byref(rs), ptr(lab), ptr(lab_off), _lib.MODEL[model], ptr(out)),
          "pob_forward")
    return out[:n].copy()


def viterbi_acceptor_batch(arrays, labels, band_size=1000, rc=None, layout=_lib.BLANK_LAST, device=None):
    """Best alignment of each label string to its read.  Returns (paths, status): int64 arrays of length T with
    4 (n_states-1) for blank or the base index emitted at that timestep.

    replaces decoding_cpp.cpp_viterbi_acceptor (decoding_cpp.pyx:69-84, Forward.h:14-121)."""
    b = _as_batch(arrays, rc, layout)
    n = b.n
    ctx = get_ctx(device)
    for lab, T in zip(labels, b.lens):
        if len(lab) == 0:
            raise IndexError("viterbi_acceptor: empty label (the reference reads label_int[0], Forward.h:51)")
    idx = [np.frombuffer(l.encode(), dtype=np.uint8) for l in labels]
    lut = np.full(256, 255, dtype=np.uint8)
    lut[np.frombuffer(b"ACGT", dtype=np.uint8)] = np.arange(4, dtype=np.uint8)
    lab_off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum([len(x) for x in idx], out=lab_off[1:])
    lab = np.zeros(int(lab_off[-1]) + 8, dtype=np.uint8)
    for x, o in zip(idx, lab_off[:-1]):
        lab[o:o + len(x)] = lut[x]
    if (lab[:int(lab_off[-1])] > 3).any():
        raise KeyError("viterbi_acceptor: label characters must be in ACGT")
    path = np.zeros(max(b.total_rows, 1) + 4, dtype=np.int8)
    st = np.zeros(max(n, 1), dtype=np.int32)
    rs = b.struct()
    check(lib().pob_viterbi_acceptor(ctx.h, _lib.HOST, C.byref(rs), ptr(lab), ptr(lab_off), int(band_size), ptr(path),
                                     ptr(st)), "pob_viterbi_acceptor")
    return [path[o:o + t].astype(np.int64) for o, t in zip(b.row_off[:-1], b.lens)], st[:n].copy()


def align_global_batch(seqs1, seqs2, match=2, mismatch=-1, gap_cost=-1, return_dp=False, device=None):
    """Full Needleman-Wunsch for many pairs.  Returns list of (row1, row2, matches[, dp]).

    replaces align.global_pair (align.pyx:29-98)."""
    n = len(seqs1)
    ctx = get_ctx(device)
    b1, o1 = _pack_bytes(seqs1)
    b2, o2 = _pack_bytes(seqs2)
    aln_off = o1 + o2 + 8 * np.arange(n + 1, dtype=np.int64)
    a1 = np.zeros(int(aln_off[-1]) + 1, dtype=np.uint8)
    a2 = np.zeros(int(aln_off[-1]) + 1, dtype=np.uint8)
    alen = np.zeros(max(n, 1), dtype=np.int32)
    mat = np.zeros(max(n, 1), dtype=np.int32)
    dp_off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum((np.diff(o1) + 1) * (np.diff(o2) + 1), out=dp_off[1:])
    dp = np.zeros(int(dp_off[-1]) + 1, dtype=np.int32) if return_dp else None
    check(lib().pob_align_global(ctx.h, _lib.HOST, ptr(b1), ptr(o1), ptr(b2), ptr(o2), n, match, mismatch, gap_cost,
                                 ptr(a1), ptr(a2), ptr(alen), ptr(mat), ptr(dp_off), ptr(dp)), "pob_align_global")
    out = []
    for p in range(n):
        o, l = int(aln_off[p]), int(alen[p])
        r = (a1[o:o + l].tobytes().decode(), a2[o:o + l].tobytes().decode(), int(mat[p]))
        if return_dp:
            l1, l2 = int(o1[p + 1] - o1[p]), int(o2[p + 1] - o2[p])
            r = r + (dp[dp_off[p]:dp_off[p + 1]].reshape(l1 + 1, l2 + 1).copy(),)
        out.append(r)
    return out


# ---- legacy prefix search (decoding/prefix_search.py of the reference) -------------------------------------------
def _pack_f64(arrays):
    """float64 log-probability tables (rows x S, blank last) packed row-wise; returns (data, row_off, S)."""
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in arrays]
    if not arrs:
        return np.zeros(8, np.float64), np.zeros(1, np.int64), 0
    S = arrs[0].shape[1] if arrs[0].ndim == 2 else 0
    for a in arrs:
        if a.ndim != 2 or a.shape[1] != S:
            raise ValueError("every table must be rows x %d" % S)
    off = np.zeros(len(arrs) + 1, dtype=np.int64)
    np.cumsum([a.shape[0] for a in arrs], out=off[1:])
    data = np.zeros(int(off[-1]) * S + 8, dtype=np.float64)
    if off[-1]:
        data[:int(off[-1]) * S] = np.concatenate([a.reshape(-1) for a in arrs])
    return data, off, S


def _unpack_labels(lab, lab_off, lens, n):
    return [lab[int(lab_off[i]):int(lab_off[i]) + int(lens[i])].copy() for i in range(n)]


def prefix_search_batch(arrays, flavour=_lib.PREFIX_CY, device=None):
    """Legacy 1D prefix search of every table (one window each).  Returns (labels, scores, status): letter-index
    arrays, label log-probabilities, POB_ST_* bits.

    replaces prefix_search.prefix_search_log / prefix_search_log_cy (prefix_search.py:116-174 / :176-238)."""
    data, off, S = _pack_f64(arrays)
    n = len(off) - 1
    if n == 0:
        return [], np.zeros(0), np.zeros(0, np.int32)
    ctx = get_ctx(device)
    lab_off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.diff(off) + 2, out=lab_off[1:])
    lab = np.zeros(int(lab_off[-1]) + 8, dtype=np.uint8)
    lens = np.zeros(n, dtype=np.int32)
    score = np.zeros(n, dtype=np.float64)
    st = np.zeros(n, dtype=np.int32)
    check(lib().pob_prefix_search(ctx.h, _lib.HOST, ptr(data), ptr(off), n, S, flavour, ptr(lab_off), ptr(lab), ptr(lens),
                                  ptr(score), ptr(st)), "pob_prefix_search")
    return _unpack_labels(lab, lab_off, lens, n), score, st


def pair_gamma_batch(arrays1, arrays2, flavour=_lib.PREFIX_NUMPY, device=None):
    """Dense (U+1) x (V+1) gamma matrix of every pair (prefix_search.py:35-65 / decoding_cy.pyx:177-220)."""
    d1, o1, S = _pack_f64(arrays1)
    d2, o2, S2 = _pack_f64(arrays2)
    n = len(o1) - 1
    if n != len(o2) - 1 or (n and S != S2):
        raise ValueError("the two reads of a pair must have the same alphabet")
    if n == 0:
        return []
    ctx = get_ctx(device)
    goff = np.zeros(n + 1, dtype=np.int64)
    np.cumsum((np.diff(o1) + 1) * (np.diff(o2) + 1), out=goff[1:])
    out = np.zeros(int(goff[-1]) + 8, dtype=np.float64)
    check(lib().pob_pair_gamma(ctx.h, _lib.HOST, ptr(d1), ptr(o1), ptr(d2), ptr(o2), n, S, flavour, ptr(goff), ptr(out)),
          "pob_pair_gamma")
    return [out[int(goff[i]):int(goff[i + 1])].reshape(int(o1[i + 1] - o1[i]) + 1, int(o2[i + 1] - o2[i]) + 1).copy()
            for i in range(n)]


def pair_prefix_search_batch(arrays1, arrays2, flavour=_lib.PREFIX_NUMPY, device=None):
    """Legacy dense 2D prefix search of every pair.  Returns (labels, scores, status).

    replaces prefix_search.pair_prefix_search_log / _cy (prefix_search.py:247-310 / :312-385)."""
    d1, o1, S = _pack_f64(arrays1)
    d2, o2, S2 = _pack_f64(arrays2)
    n = len(o1) - 1
    if n != len(o2) - 1 or (n and S != S2):
        raise ValueError("the two reads of a pair must have the same alphabet")
    if n == 0:
        return [], np.zeros(0), np.zeros(0, np.int32)
    ctx = get_ctx(device)
    lab_off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.maximum(np.diff(o1), np.diff(o2)) + 3, out=lab_off[1:])
    lab = np.zeros(int(lab_off[-1]) + 8, dtype=np.uint8)
    lens = np.zeros(n, dtype=np.int32)
    score = np.zeros(n, dtype=np.float64)
    st = np.zeros(n, dtype=np.int32)
    check(lib().pob_pair_prefix_search(ctx.h, _lib.HOST, ptr(d1), ptr(o1), ptr(d2), ptr(o2), n, S, flavour, ptr(lab_off),
                                       ptr(lab), ptr(lens), ptr(score), ptr(st)), "pob_pair_prefix_search")
    return _unpack_labels(lab, lab_off, lens, n), score, st


def forward_vec(y, s, i, previous=None, flavour=_lib.PREFIX_NUMPY, device=None):
    """One column of the 1D forward algorithm (prefix_search.py:81-97 / decoding_cy.pyx:127-156)."""
    y = np.ascontiguousarray(y, dtype=np.float64)
    if y.ndim != 2:
        raise ValueError("expected a (time, alphabet + blank) table")
    assert i == 0 or previous is not None  # prefix_search.py:84
    prev = None if previous is None else np.ascontiguousarray(previous, dtype=np.float64)
    if prev is not None and len(prev) != len(y):
        raise ValueError("previous must have one entry per timestep")
    out = np.zeros(max(len(y), 1), dtype=np.float64)
    ctx = get_ctx(device)
    check(lib().pob_forward_vec(ctx.h, _lib.HOST, ptr(y), len(y), y.shape[1], flavour, int(s), int(i),
                                ptr(prev) if prev is not None else None, ptr(out)), "pob_forward_vec")
    return out[:len(y)].copy()
