"""Loader of libfmcmcb200.so (the CUDA product library).  Fails loudly: there is no
CPU fallback anywhere in this package (north_star: a missing device path is an error)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

from . import _abi as A

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("FMCMC_B200_LIB") or os.path.join(_HERE, "libfmcmcb200.so")   # env override: profiling builds only
_lib = None


class FmcmcError(RuntimeError):
    """Error raised by the C ABI (status code + the library's message)."""

    def __init__(self, code: int, msg: str):
        super().__init__(msg)
        self.code = code


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA extension in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    src_dir = os.path.join(_HERE, "csrc")
    srcs = [os.path.join(src_dir, f) for f in os.listdir(src_dir)]
    srcs.append(os.path.join(_HERE, "..", "include", "fmcmc_b200.h"))
    stale = force or not os.path.exists(SO_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(SO_PATH) for s in srcs if os.path.isfile(s))
    if stale:
        r = subprocess.run(["make", "-C", src_dir], capture_output=True, text=True)
        if verbose:
            print(r.stdout[-4000:], r.stderr[-8000:])
        if r.returncode != 0:
            raise RuntimeError("nvcc build of libfmcmcb200.so failed:\n" + r.stderr[-4000:])
    return SO_PATH


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(
            f"{SO_PATH} is missing: the CUDA extension has not been built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'`). fmcmc_b200 has no CPU fallback.")
    L = C.CDLL(SO_PATH)
    dp = C.POINTER(C.c_double)
    u8p = C.POINTER(C.c_uint8)
    vp = C.c_void_p
    L.fmcmc_version.restype = C.c_int
    L.fmcmc_device_count.restype = C.c_int
    L.fmcmc_model_nparams.restype = C.c_int32
    L.fmcmc_model_nparams.argtypes = [C.POINTER(A.ModelDesc)]
    L.fmcmc_kernel_state_len.restype = C.c_int64
    L.fmcmc_kernel_state_len.argtypes = [C.c_int32, C.c_int32, C.c_int32]
    L.fmcmc_rows_kept.restype = C.c_int64
    L.fmcmc_rows_kept.argtypes = [C.c_int64, C.c_int64, C.c_int64]
    for name in ("fmcmc_model_create", "fmcmc_model_create_device"):
        f = getattr(L, name)
        f.restype = C.c_int
        f.argtypes = [C.POINTER(A.ModelDesc), C.c_int, C.POINTER(vp), C.c_char_p, C.c_size_t]
    L.fmcmc_model_free.restype = None
    L.fmcmc_model_free.argtypes = [vp]
    L.fmcmc_event_mark.restype = C.c_int
    L.fmcmc_event_mark.argtypes = [vp, C.c_int]
    L.fmcmc_event_elapsed_ms.restype = C.c_int
    L.fmcmc_event_elapsed_ms.argtypes = [vp, C.c_int, C.c_int, dp]
    L.fmcmc_model_trim.restype = C.c_int
    L.fmcmc_model_trim.argtypes = [vp, C.c_int, C.c_char_p, C.c_size_t]
    L.fmcmc_kernel_state_fetch.restype = C.c_int
    L.fmcmc_kernel_state_fetch.argtypes = [vp, C.POINTER(A.KernelState), C.c_char_p, C.c_size_t]
    L.fmcmc_store_ess.restype = C.c_int
    L.fmcmc_store_ess.argtypes = [vp, C.c_int64, C.c_int64, u8p, C.c_int32, dp, C.POINTER(C.c_int32), C.c_char_p, C.c_size_t]
    L.fmcmc_store_pooled.restype = C.c_int
    L.fmcmc_store_pooled.argtypes = [vp, u8p, dp, C.c_char_p, C.c_size_t]
    L.fmcmc_set_path.restype = C.c_int
    L.fmcmc_set_path.argtypes = [vp, C.c_int]
    L.fmcmc_run.restype = C.c_int
    L.fmcmc_run.argtypes = A.RUN_ARGTYPES
    L.fmcmc_logpost.restype = C.c_int
    L.fmcmc_logpost.argtypes = [vp, C.c_int32, dp, dp, C.c_char_p, C.c_size_t]
    L.fmcmc_store_reset.restype = C.c_int
    L.fmcmc_store_reset.argtypes = [vp, C.c_int32, C.c_int32, C.c_int64, C.c_char_p, C.c_size_t]
    L.fmcmc_store_rows.restype = C.c_int64
    L.fmcmc_store_rows.argtypes = [vp]
    L.fmcmc_gelman_partials.restype = C.c_int
    L.fmcmc_gelman_partials.argtypes = [vp, C.c_int64, C.c_int64, u8p, vp, vp, vp, C.c_int, C.c_char_p, C.c_size_t]
    L.fmcmc_gelman_finish.restype = C.c_int
    L.fmcmc_gelman_finish.argtypes = [vp, C.c_int64, C.c_int64, C.c_int32, vp, vp, vp, C.c_int, dp, dp,
                                      C.c_char_p, C.c_size_t]
    L.fmcmc_gelman.restype = C.c_int
    L.fmcmc_gelman.argtypes = [vp, u8p, C.c_int64, C.c_int64, dp, dp, C.POINTER(C.c_int64), C.c_char_p, C.c_size_t]
    L.fmcmc_gelman_window_begin.restype = C.c_int64
    L.fmcmc_gelman_window_begin.argtypes = [C.c_int64, C.c_int64, C.c_int64]
    L.fmcmc_host_sym_eigmax.restype = C.c_int
    L.fmcmc_host_sym_eigmax.argtypes = [C.c_int32, dp, dp]
    L.fmcmc_cov_recursive.restype = C.c_int
    L.fmcmc_cov_recursive.argtypes = [C.c_int, C.c_int32, C.c_int64, dp, dp, dp, C.c_double, C.c_double,
                                      C.c_double, dp, dp, dp, C.c_char_p, C.c_size_t]
    L.fmcmc_reflect.restype = C.c_int
    L.fmcmc_reflect.argtypes = [C.c_int, C.c_int32, C.c_int64, dp, dp, dp, u8p, C.c_char_p, C.c_size_t]
    L.fmcmc_shard_alloc.restype = C.c_int
    L.fmcmc_shard_alloc.argtypes = [vp, C.c_int, C.c_int, C.c_int64, C.POINTER(A.ShardHandles), C.c_char_p, C.c_size_t]
    L.fmcmc_shard_attach.restype = C.c_int
    L.fmcmc_shard_attach.argtypes = [vp, C.c_int, C.c_int, C.POINTER(A.ShardHandles), C.c_char_p, C.c_size_t]
    L.fmcmc_measure_fp64_peak.restype = C.c_int
    L.fmcmc_measure_fp64_peak.argtypes = [C.c_int, dp, C.c_char_p, C.c_size_t]
    L.fmcmc_test_softplus.restype = C.c_int
    L.fmcmc_test_softplus.argtypes = [C.c_int, C.c_int64, dp, dp, C.c_char_p, C.c_size_t]
    if L.fmcmc_version() != A.ABI_VERSION:
        raise ImportError("libfmcmcb200.so ABI version mismatch")
    _lib = L
    return L


EXPORTED_SYMBOLS = [
    "fmcmc_version", "fmcmc_device_count", "fmcmc_model_nparams", "fmcmc_kernel_state_len",
    "fmcmc_rows_kept", "fmcmc_model_create", "fmcmc_model_create_device", "fmcmc_model_free", "fmcmc_model_trim", "fmcmc_event_mark", "fmcmc_event_elapsed_ms",
    "fmcmc_set_path", "fmcmc_run", "fmcmc_logpost", "fmcmc_store_reset", "fmcmc_store_rows",
    "fmcmc_gelman_partials", "fmcmc_gelman_finish", "fmcmc_gelman", "fmcmc_gelman_window_begin", "fmcmc_store_pooled", "fmcmc_store_ess", "fmcmc_kernel_state_fetch", "fmcmc_host_sym_eigmax", "fmcmc_shard_alloc", "fmcmc_shard_attach", "fmcmc_cov_recursive", "fmcmc_reflect", "fmcmc_measure_fp64_peak", "fmcmc_test_softplus",
]


def check(rc: int, errbuf) -> None:
    if rc != 0:
        raise FmcmcError(rc, errbuf.value.decode(errors="replace"))


def errbuf():
    return C.create_string_buffer(1024)
