"""Batched host-side API over the C ABI: packing of reads and the calls the drivers actually make.

The single-item functions in poreover_b200.decoding / poreover_b200.align (which keep the reference's
signatures) are thin wrappers over these with a batch of one.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import ReadsT, check, get_ctx, lib, ptr

ALIGN_ROWS = 4  # float32 x 5 states: 4 rows = 80 B keeps every read 16-byte aligned


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


def viterbi_batch(arrays, kind, rc=None, layout=_lib.BLANK_LAST, return_path=False, device=None):
    """Best-path decode of many reads.  Returns (sequences, s2s lists, paths or None, status array).

    replaces transducer.{poreover,bonito}.viterbi_decode + get_sequence_mapping (see include/poreover_b200.h)."""
    b = arrays if isinstance(arrays, ReadBatch) else ReadBatch(arrays, rc=rc, layout=layout)
    ctx = get_ctx(device)
    rows = max(b.total_rows, 1)
    seq = np.zeros(rows, dtype=np.uint8)
    s2s = np.zeros(rows, dtype=np.int32)
    path = np.zeros(rows, dtype=np.int8) if return_path else None
    ln = np.zeros(max(b.n, 1), dtype=np.int32)
    st = np.zeros(max(b.n, 1), dtype=np.int32)
    rs = b.struct()
    check(lib().pob_viterbi(ctx.h, _lib.HOST, C.byref(rs), _lib.KIND[kind], ptr(seq), ptr(s2s), ptr(path), ptr(ln),
                            ptr(st)), "pob_viterbi")
    offs = b.row_off[:-1]
    seqs = [seq[o:o + l].tobytes().decode() for o, l in zip(offs, ln[:b.n])]
    maps = [s2s[o:o + l].astype(np.int64) for o, l in zip(offs, ln[:b.n])]
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
    for s in seqs1:
        if len(s) == 0:
            raise ZeroDivisionError("float division by zero")  # align.pyx:122 with l1 == 0
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
