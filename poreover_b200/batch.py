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
