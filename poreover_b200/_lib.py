"""ctypes binding of libporeover_b200.so (the C ABI declared in include/poreover_b200.h).

There is NO CPU fallback: if the CUDA library is missing or no GPU is present, every decoding call
raises.  Loading the library itself needs no GPU (used by the CPU-only symbol tests).
"""
import ctypes as C
import os
import threading

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("POB_DEBUG_LIB") or os.path.join(HERE, "libporeover_b200.so")  # override: instrumented debug builds (tools/)

HOST, DEVICE = 0, 1
F32, F64, U8_TRACE = 0, 1, 2
BLANK_LAST, BLANK_FIRST = 0, 1
KIND = {"poreover": 0, "bonito": 1, "flipflop": 2}
MODEL = {"ctc": 0, "ctc_merge_repeats": 1}
METHOD = {"row": 0, "row_col": 1}

ST_SHORT_BEAM_SKIP, ST_UNSET_BAND, ST_POOL_OVERFLOW, ST_MAPPING_WRAP = 1, 2, 4, 8
ST_SKIPPED_LENGTH, ST_SKIPPED_IDENTITY, ST_EMPTY = 16, 32, 64
ST_MAX_DEPTH = 128
PREFIX_NUMPY, PREFIX_CY = 0, 1  # arithmetic flavours of the legacy prefix search

K_NAMES = ["viterbi_ctc", "viterbi_flipflop", "nw_band_fill", "nw_traceback", "envelope", "beam_pair",
           "beam_single", "backtrace", "forward", "acceptor", "prefix_search", "pair_gamma", "pair_prefix_search"]

vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double

# name -> (restype, argtypes); must list every symbol of include/poreover_b200.h
SIGNATURES = {
    "pob_abi_version": (i32, []),
    "pob_strerror": (C.c_char_p, [i32]),
    "pob_last_cuda_error": (C.c_char_p, []),
    "pob_device_count": (i32, [C.POINTER(i32)]),
    "pob_ctx_create": (i32, [i32, C.POINTER(vp)]),
    "pob_ctx_destroy": (i32, [vp]),
    "pob_ctx_sync": (i32, [vp]),
    "pob_ctx_stream": (vp, [vp]),
    "pob_ctx_device": (i32, [vp]),
    "pob_malloc": (i32, [vp, C.c_size_t, C.POINTER(vp)]),
    "pob_free": (i32, [vp, vp]),
    "pob_malloc_host": (i32, [C.c_size_t, C.POINTER(vp)]),
    "pob_free_host": (i32, [vp]),
    "pob_memcpy_h2d": (i32, [vp, vp, vp, C.c_size_t]),
    "pob_memcpy_d2h": (i32, [vp, vp, vp, C.c_size_t]),
    "pob_timer_start": (i32, [vp]),
    "pob_timer_stop": (i32, [vp, C.POINTER(dbl)]),
    "pob_profile_enable": (i32, [vp, i32]),
    "pob_profile_reset": (i32, [vp]),
    "pob_profile_get": (i32, [vp, i32, C.POINTER(dbl), C.POINTER(i64)]),
    "pob_kernel_name": (C.c_char_p, [i32]),
    "pob_viterbi": (i32, [vp, i32, vp, i32, vp, vp, vp, vp, vp]),
    "pob_viterbi_flipflop": (i32, [vp, i32, vp, vp, vp, vp, vp, vp]),
    "pob_align_banded": (i32, [vp, i32, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp]),
    "pob_align_global": (i32, [vp, i32, vp, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp]),
    "pob_build_envelope": (i32, [vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, vp]),
    "pob_beam_search": (i32, [vp, i32, vp, i32, i32, vp, vp, vp, vp]),
    "pob_beam_search_2d": (i32, [vp, i32, vp, vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp]),
    "pob_forward": (i32, [vp, i32, vp, vp, vp, i32, vp]),
    "pob_viterbi_acceptor": (i32, [vp, i32, vp, vp, vp, i32, vp, vp]),
    "pob_pair_decode": (i32, [vp, i32, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "pob_counters": (i32, [vp, vp]),
    "pob_forward_vec": (i32, [vp, i32, vp, i32, i32, i32, i32, i32, vp, vp]),
    "pob_prefix_search": (i32, [vp, i32, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp]),
    "pob_pair_gamma": (i32, [vp, i32, vp, vp, vp, vp, i32, i32, i32, vp, vp]),
    "pob_pair_prefix_search": (i32, [vp, i32, vp, vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp]),
    "pob_npy_probe": (i32, [vp, i32, i32, vp, vp, vp, vp, vp]),
    "pob_npy_read": (i32, [vp, i32, i32, vp, i64, vp, vp, vp, vp]),
}


class ReadsT(C.Structure):
    """pob_reads_t"""
    _fields_ = [("data", vp), ("row_off", vp), ("row_len", vp), ("rc", vp), ("n", C.c_int32),
                ("n_states", C.c_int32), ("dtype", C.c_int32), ("layout", C.c_int32)]


class PoreoverB200Error(RuntimeError):
    pass


_lib = None
_lock = threading.Lock()
_ctx_lock = threading.Lock()
_ctxs = {}


def lib():
    """Load the CUDA library.  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise PoreoverB200Error(
                        "%s is missing: the CUDA extension has not been built (run "
                        "`python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback." % LIB_PATH)
                l = C.CDLL(LIB_PATH)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(l, name)  # AttributeError if the build lacks a declared symbol
                    fn.restype = res
                    fn.argtypes = args
                _lib = l
    return _lib


def check(status, what=""):
    if status != 0:
        l = lib()
        msg = l.pob_strerror(status).decode()
        if status == -2:
            msg += ": " + l.pob_last_cuda_error().decode()
        raise PoreoverB200Error("%s failed: %s" % (what or "poreover_b200 call", msg))


def device_count():
    n = i32(0)
    st = lib().pob_device_count(C.byref(n))
    return n.value if st == 0 else 0


class Context:
    """One per GPU: stream + scratch arena (pob_ctx)."""

    def __init__(self, device=0):
        self.device = device
        h = vp()
        check(lib().pob_ctx_create(device, C.byref(h)), "pob_ctx_create(device=%d)" % device)
        self.h = h

    def sync(self):
        check(lib().pob_ctx_sync(self.h), "pob_ctx_sync")

    def close(self):
        if self.h:
            lib().pob_ctx_destroy(self.h)
            self.h = None

    # device memory helpers -------------------------------------------------
    def malloc(self, nbytes):
        p = vp()
        check(lib().pob_malloc(self.h, nbytes, C.byref(p)), "pob_malloc(%d)" % nbytes)
        return p.value

    def free(self, ptr):
        check(lib().pob_free(self.h, vp(ptr)), "pob_free")

    def to_device(self, arr):
        arr = np.ascontiguousarray(arr)
        d = self.malloc(max(arr.nbytes, 1) + 64)
        check(lib().pob_memcpy_h2d(self.h, vp(d), arr.ctypes.data_as(vp), arr.nbytes), "h2d")
        self.sync()
        return d

    def from_device(self, ptr, shape, dtype):
        out = np.empty(shape, dtype=dtype)
        check(lib().pob_memcpy_d2h(self.h, out.ctypes.data_as(vp), vp(ptr), out.nbytes), "d2h")
        self.sync()
        return out

    def timer_start(self):
        check(lib().pob_timer_start(self.h), "pob_timer_start")

    def timer_stop(self):
        ms = dbl(0)
        check(lib().pob_timer_stop(self.h, C.byref(ms)), "pob_timer_stop")
        return ms.value

    # profiling ---------------------------------------------------------------
    def profile(self, on=True):
        check(lib().pob_profile_enable(self.h, 1 if on else 0))

    def profile_reset(self):
        check(lib().pob_profile_reset(self.h))

    def profile_get(self):
        out = {}
        for k, name in enumerate(K_NAMES):
            ms, n = dbl(0), i64(0)
            check(lib().pob_profile_get(self.h, k, C.byref(ms), C.byref(n)))
            if n.value:
                out[name] = {"ms": ms.value, "launches": n.value}
        return out

    def counters(self):
        a = np.zeros(3, dtype=np.int64)
        check(lib().pob_counters(self.h, a.ctypes.data_as(vp)))
        return {"cell_updates": int(a[0]), "steps": int(a[1]), "launches": int(a[2])}


def get_ctx(device=None):
    """Process-wide context for `device` (default: $LOCAL_RANK or 0); a Context passes through."""
    if isinstance(device, Context):
        return device
    if device is None:
        device = int(os.environ.get("POREOVER_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    lib()  # loaded before _ctx_lock is taken (lib() has its own lock)
    with _ctx_lock:  # creation is atomic: the command-line pipeline asks from two GPU threads at once
        c = _ctxs.get(device)
        if c is None:
            if device_count() <= device:
                raise PoreoverB200Error(
                    "no CUDA device %d visible: poreover_b200 has no CPU fallback (%s)"
                    % (device, lib().pob_last_cuda_error().decode()))
            c = _ctxs[device] = Context(device)
    return c


_free_ctxs = {}  # device -> contexts not in use by a borrow_ctx() block


class borrow_ctx:
    """`with borrow_ctx(device) as ctx:` -- a context (stream + scratch arena) nobody else is using right now.  The
    command-line pipeline runs two GPU threads per device so that consecutive batches overlap; each call borrows
    a context for its duration.  The process-wide context of get_ctx() is the first one handed out, further ones are
    created on demand and kept for the life of the process."""

    def __init__(self, device=None):
        if device is None:
            device = int(os.environ.get("POREOVER_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        self.device, self.ctx = device, None

    def __enter__(self):
        first = get_ctx(self.device)
        with _lock:
            free = _free_ctxs.setdefault(self.device, [first])
            if free:
                self.ctx = free.pop()
        if self.ctx is None:
            self.ctx = Context(self.device)
        return self.ctx

    def __exit__(self, *exc):
        with _lock:
            _free_ctxs[self.device].append(self.ctx)
        self.ctx = None
        return False


def ptr(a):
    return None if a is None else a.ctypes.data_as(vp)
