"""Batched ingest of basecaller output files for the command-line drivers.

The reference loads one file per pool task (decode.py:41-51, :67-112, pair_decode.py:321-356).  With the search on
the GPU the loader is what a `pair-decode` run waits for: np.load's header machinery, the float64 copy and the
all-ones transition table of every transducer object, the column permutation and the packing copy cost more host
time per pair than the GPU needs to decode it.  Here the files of a chunk are read by a thread pool (file reads and
numpy's ufunc loops release the GIL) and the logarithm is written straight into the packed buffer that crosses the
C ABI:

  * the arithmetic is still numpy's, on the same C-contiguous float32 array, so every value is bit-identical to what
    decode.load_logits returns (a CUDA logf would not be, and Viterbi ties would move: DESIGN.md section 1);
  * bonito files stay in file order (blank first): the kernels resolve the layout in their address computation
    (POB_BLANK_FIRST), which replaces the permutation copy of decode.py:79;
  * anything that is not a plain 2-D probability table (.csv, 3-D logits, float64 files, .hdf5/.fast5) goes through
    decode.model_from_trace unchanged and is copied into the batch.
"""
import ctypes as C
import os
import re
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import _lib
from .batch import ALIGN_ROWS, ReadBatch

_pool = None
_pool_lock = threading.Lock()


def n_threads():
    env = os.environ.get("POREOVER_B200_LOADER_THREADS")
    if env:
        return max(1, int(env))
    cores = os.cpu_count() or 4
    try:  # under torchrun the ranks of a box share its cores
        cores //= max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
    except ValueError:
        pass
    return max(2, min(32, cores))


def pool():
    """The loader's thread pool (created on first use, shared by all chunks of a run)."""
    global _pool
    with _pool_lock:
        if _pool is None:
            _pool = ThreadPoolExecutor(max_workers=n_threads(), thread_name_prefix="pob-load")
        return _pool


_HDR = re.compile(r"^\{'descr': '([<|=]?[fiu]\d)', 'fortran_order': False, 'shape': \((\d*(?:, ?\d+)*),?\), \}\s*$")


def read_npy(path):
    """np.load(path) for the plain little-endian C-order arrays bonito and PoreOverNet write, without the Python-level
    header parsing (most of np.load's time on a 100 KB file).  Anything unusual falls back to np.load."""
    with open(path, 'rb') as f:
        buf = f.read()
    if buf[:6] == b'\x93NUMPY' and len(buf) >= 12:
        major = buf[6]
        if major == 1:
            hlen, start = int.from_bytes(buf[8:10], 'little'), 10
        elif major in (2, 3):
            hlen, start = int.from_bytes(buf[8:12], 'little'), 12
        else:
            return np.load(path)
        m = _HDR.match(buf[start:start + hlen].decode('latin1'))
        if m:
            dtype = np.dtype(m.group(1))
            shape = tuple(int(x) for x in m.group(2).replace(' ', '').split(',') if x)
            count = int(np.prod(shape)) if shape else 1
            if len(buf) - start - hlen >= count * dtype.itemsize:
                return np.frombuffer(buf, dtype=dtype, count=count, offset=start + hlen).reshape(shape)
    return np.load(path)


def _is_probability_row(row):
    """np.isclose(np.sum(row), 1) of decode.py:43, deciding the clear cases without numpy's array machinery."""
    s = np.sum(row)
    d = abs(float(s) - 1.0)
    if d < 5e-6:
        return True
    if d > 2e-5 or d != d:  # the test is |s - 1| <= 1e-8 + 1e-5; NaN is never close
        return False
    return bool(np.isclose(s, 1))


class _Raw:
    """What pass 1 learned about one file."""
    __slots__ = ("arr", "direct", "T", "dtype", "kind")


def _stage1(path, basecaller):
    from .decoding import decode
    r = _Raw()
    ext = os.path.splitext(path)[1]
    if ext == '.npy' and basecaller in ('poreover', 'bonito'):
        raw = read_npy(path)
        if raw.ndim == 2 and raw.dtype == np.float32 and raw.flags.c_contiguous and _is_probability_row(raw[0]):
            assert raw.shape[1] == 5  # transducer.py:21
            r.arr, r.direct, r.T, r.dtype, r.kind = raw, True, raw.shape[0], np.dtype(np.float32), basecaller
            return r
    # everything else: the reference's loader path, then the model's own array (blank last)
    m = decode.model_from_trace(path, basecaller)
    a = m.device_array()
    r.arr, r.direct, r.T, r.dtype, r.kind = a, False, a.shape[0], a.dtype, m.kind
    return r


def _chunks(n, parts):
    step = max(1, -(-n // max(1, parts)))
    return [(i, min(n, i + step)) for i in range(0, n, step)]


def _load_reads_native(paths, basecaller, rc, alloc):
    """The common case -- every file a plain 2-D float32 .npy table of probabilities -- without a Python-level step
    per file: headers and payloads are read by native threads (pob_npy_probe / pob_npy_read, csrc/hostio.cu) straight
    into the packed buffer, then numpy's own log runs in place over it (decode.py:45: same ufunc, same values).
    Returns None when some file is anything else: the caller takes the general path."""
    n = len(paths)
    if n == 0 or basecaller not in ('poreover', 'bonito') or any(os.path.splitext(p)[1] != '.npy' for p in paths):
        return None
    L = _lib.lib()
    cpaths = (C.c_char_p * n)(*[os.fsencode(p) for p in paths])
    rows, cols, doff = (np.zeros(n, np.int64) for _ in range(3))
    flags, sums = np.zeros(n, np.int32), np.zeros(n, np.float32)
    nt = n_threads()
    _lib.check(L.pob_npy_probe(cpaths, n, nt, _lib.ptr(rows), _lib.ptr(cols), _lib.ptr(doff), _lib.ptr(flags),
                               _lib.ptr(sums)), "pob_npy_probe")
    bad = np.flatnonzero(flags == 2)
    if len(bad):
        raise OSError("cannot read %s" % paths[int(bad[0])])
    if (flags != 0).any() or (cols != 5).any() or (rows <= 0).any():
        return None
    # np.isclose(np.sum(arr[0]), 1) of decode.py:43: clear cases decided on the float32 sums, the rest by numpy itself
    d = np.abs(sums.astype(np.float64) - 1.0)
    prob = d < 5e-6
    unclear = ~prob & ~(d > 2e-5) & ~np.isnan(d)
    if unclear.any():
        prob |= unclear & np.isclose(sums, 1)
    if not prob.all():
        return None  # logits: the reference's log-softmax path (decode.py:47-50)
    b = ReadBatch.__new__(ReadBatch)
    b.np_dtype = np.dtype(np.float32)
    b.dtype, b.n, b.n_states = _lib.F32, n, 5
    b.kinds = [basecaller] * n
    b.layout = _lib.BLANK_FIRST if basecaller == 'bonito' else _lib.BLANK_LAST
    b.lens = rows.astype(np.int32)
    padded = (rows + ALIGN_ROWS - 1) // ALIGN_ROWS * ALIGN_ROWS
    b.row_off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(padded, out=b.row_off[1:])
    b.total_rows = int(b.row_off[-1])
    b.data = (alloc or np.empty)((b.total_rows, 5), b.np_dtype)
    ok = np.zeros(n, np.int32)
    _lib.check(L.pob_npy_read(cpaths, n, nt, _lib.ptr(rows), 5, _lib.ptr(doff), _lib.ptr(b.row_off), _lib.ptr(b.data),
                              _lib.ptr(ok)), "pob_npy_read")
    if not ok.all():
        raise OSError("cannot read %s" % paths[int(np.flatnonzero(ok == 0)[0])])
    data, off, lens = b.data, b.row_off, b.lens

    def logs(lohi):
        with np.errstate(divide="ignore"):
            for i in range(*lohi):
                v = data[off[i]:off[i] + lens[i]]
                np.log(v, out=v)  # the ufunc loop runs without the interpreter lock

    list(pool().map(logs, _chunks(n, 4 * nt)))
    b.rc = None if rc is None else np.ascontiguousarray(np.broadcast_to(np.asarray(rc, dtype=np.uint8), (n,)))
    return b


def load_reads(paths, basecaller, rc=None, alloc=None):
    """Load many files of one basecaller into one packed ReadBatch (host memory).

    Equivalent to ReadBatch([decode.model_from_trace(p, basecaller).device_array() for p in paths], rc=rc) up to the
    column layout: when every file is a plain bonito probability table the batch keeps the file order and says so
    (layout BLANK_FIRST).  alloc(shape, dtype) may supply the packed buffer (e.g. pinned memory, any contents).
    The batch also carries .kinds, the transducer kind of every read."""
    paths = list(paths)
    n = len(paths)
    if os.environ.get("POREOVER_B200_NATIVE_LOADER", "1") != "0":
        fast = _load_reads_native(paths, basecaller, rc, alloc)
        if fast is not None:
            return fast
    ex = pool()
    parts = _chunks(n, 4 * n_threads())
    raws = [None] * n

    def s1(lohi):
        for i in range(*lohi):
            raws[i] = _stage1(paths[i], basecaller)

    list(ex.map(s1, parts))
    all_direct = n > 0 and all(r.direct for r in raws)
    first_blank = all_direct and basecaller == 'bonito'
    f32 = n > 0 and all(r.dtype == np.float32 for r in raws)
    b = ReadBatch.__new__(ReadBatch)
    b.np_dtype = np.dtype(np.float32 if f32 else np.float64)
    b.dtype = _lib.F32 if f32 else _lib.F64
    b.n = n
    b.n_states = 5
    b.kinds = [r.kind for r in raws]  # transducer.kind of every read ('poreover' / 'bonito' / 'flipflop')
    if 'flipflop' in b.kinds:
        raise NotImplementedError("flip-flop traces do not go through the packed 5-state loader")
    for r in raws:
        if r.arr.ndim != 2 or r.arr.shape[1] != 5:
            raise ValueError("all reads of a batch must have the same number of states")
    b.layout = _lib.BLANK_FIRST if first_blank else _lib.BLANK_LAST
    b.lens = np.array([r.T for r in raws], dtype=np.int32)
    padded = (b.lens.astype(np.int64) + ALIGN_ROWS - 1) // ALIGN_ROWS * ALIGN_ROWS
    b.row_off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(padded, out=b.row_off[1:])
    b.total_rows = int(b.row_off[-1])
    b.data = (alloc or np.zeros)((b.total_rows, 5), b.np_dtype)
    data, off = b.data, b.row_off

    def s2(lohi):
        with np.errstate(divide="ignore"):
            for i in range(*lohi):
                r = raws[i]
                dst = data[off[i]:off[i] + r.T]
                if r.direct and first_blank:
                    np.log(r.arr, out=dst)  # decode.py:45, written where the kernels read it
                elif r.direct:
                    if basecaller == 'bonito':  # mixed batch: permute like decode.py:79
                        lg = np.log(r.arr)
                        dst[:, :4] = lg[:, 1:]
                        dst[:, 4] = lg[:, 0]
                    else:
                        np.log(r.arr, out=dst)
                else:
                    dst[...] = r.arr  # float32 -> float64 widening is exact
                data[off[i] + r.T:off[i + 1]] = 0  # alignment rows between reads
                raws[i] = None

    list(ex.map(s2, parts))
    b.rc = None if rc is None else np.ascontiguousarray(np.broadcast_to(np.asarray(rc, dtype=np.uint8), (n,)))
    return b


def load_models(paths, basecaller):
    """[decode.model_from_trace(p, basecaller) for p in paths] on the loader's thread pool."""
    from .decoding import decode
    paths = list(paths)
    out = [None] * len(paths)

    def work(lohi):
        for i in range(*lohi):
            out[i] = decode.model_from_trace(paths[i], basecaller)

    list(pool().map(work, _chunks(len(paths), 4 * n_threads())))
    return out


class Lookahead:
    """Pipeline over a stream of chunks: while the caller consumes chunk k (result formatting, file output), one
    background thread runs load(..) a chunk ahead and, when `work` is given, `workers` more threads run
    work(load's payload) -- the GPU call, which releases the GIL.  With two workers (each on its own context and
    stream, _lib.thread_ctx) the kernels of chunk k+1 fill the SMs that the draining last wave of chunk k leaves idle.
    Chunks are pulled from `source` (a callable returning the next chunk or None) on the CALLER's thread, so a
    distributed work queue is only ever touched from there, and results are yielded in chunk order:
    (chunk, work(load(chunk))), or (chunk, load(chunk)) without a work stage."""

    def __init__(self, source, load, work=None, workers=None):
        if workers is None:
            workers = int(os.environ.get("POREOVER_B200_GPU_LANES", "2"))
        self._source, self._load, self._work = source, load, work
        self._ex_load = ThreadPoolExecutor(max_workers=1, thread_name_prefix="pob-prefetch")
        self._ex_work = ThreadPoolExecutor(max_workers=max(1, workers), thread_name_prefix="pob-gpu") if work else None
        self._depth = (1 + max(1, workers)) if work else 1  # chunks in flight behind the one being consumed
        self._pending = []
        self._fill()

    def _fill(self):
        while len(self._pending) < self._depth:
            c = self._source()
            if c is None:
                return
            fut = self._ex_load.submit(self._load, c)
            if self._ex_work is not None:
                fut = self._ex_work.submit(lambda f=fut: self._work(f.result()))
            self._pending.append((c, fut))

    def __iter__(self):
        try:
            while self._pending:
                c, fut = self._pending.pop(0)
                self._fill()  # the following chunks are under way before this one is consumed
                yield c, fut.result()
                self._fill()
        finally:
            for _, fut in self._pending:
                fut.cancel()
            self._ex_load.shutdown(wait=True)
            if self._ex_work is not None:
                self._ex_work.shutdown(wait=True)


class PinnedPool:
    """Page-locked host buffers (pob_malloc_host) for the packed batches, reused from chunk to chunk: a buffer goes
    back to the pool when the last numpy view of it dies.  Pinning costs more than the copy it speeds up (0.2 s per
    GB), so nothing ever waits for it: a request that finds no free buffer gets None (the caller uses pageable
    memory for this chunk) and a background thread pins one more buffer of that size for the chunks to come."""

    def __init__(self, keep=8):
        self._free, self._lock, self._keep = [], threading.Lock(), keep
        self._owned = 0            # buffers pinned so far (free or in use)
        self._wanted = []          # sizes the background thread still has to pin
        self._thread = None
        self._closed = False

    def _give(self, cap, addr):
        with self._lock:
            self._free.append((cap, addr))

    def _pin_loop(self):
        while True:
            with self._lock:
                if not self._wanted or self._closed:
                    self._thread = None
                    return
                cap = self._wanted.pop(0)
            h = C.c_void_p()
            ok = _lib.lib().pob_malloc_host(C.c_size_t(cap), C.byref(h)) == 0 and h.value
            with self._lock:
                if ok:
                    self._free.append((cap, h.value))
                else:
                    self._owned -= 1   # no pinned memory to be had (no GPU): stay on pageable buffers

    def close(self):
        """Stop pinning (interpreter exit): wait for the allocation in flight, leave the rest to the OS."""
        with self._lock:
            self._closed = True
            t = self._thread
        if t is not None:
            t.join(timeout=5)

    def empty(self, shape, dtype):
        """np.empty(shape, dtype) in pinned memory, or None when no pinned buffer of that size is free right now."""
        import weakref
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(shape)) * dtype.itemsize
        if nbytes == 0:
            return np.zeros(shape, dtype)
        with self._lock:
            fit = [x for x in self._free if x[0] >= nbytes]
            if not fit:
                if self._owned < self._keep and not self._closed:
                    self._owned += 1
                    self._wanted.append((nbytes * 5 // 4 + 4095) & ~4095)  # headroom: chunks differ a little in size
                    if self._thread is None:
                        self._thread = threading.Thread(target=self._pin_loop, name="pob-pin", daemon=True)
                        self._thread.start()
                return None
            cap, addr = min(fit)
            self._free.remove((cap, addr))
        cbuf = (C.c_char * cap).from_address(addr)
        weakref.finalize(cbuf, self._give, cap, addr)
        return np.frombuffer(cbuf, dtype=np.uint8, count=nbytes).view(dtype).reshape(shape)


_pinned = None


def packed_alloc():
    """Allocator of load_reads' packed buffers: the pinned pool when POREOVER_B200_PINNED=1, else numpy's.  Opt-in:
    measured again in round 2 on a 40,960-pair run (10 chunks), the pool on by default did 4,077 pairs/s against 5,197
    with pageable batches -- cudaMallocHost in the background stalls the GPU threads' launches for longer than the
    faster copies save."""
    global _pinned
    if os.environ.get("POREOVER_B200_PINNED", "0") != "1":
        return None
    with _pool_lock:
        if _pinned is None:
            import atexit
            _pinned = PinnedPool()
            atexit.register(_pinned.close)

    def alloc(shape, dtype):
        a = _pinned.empty(shape, dtype)
        return a if a is not None else np.zeros(shape, dtype)

    return alloc
