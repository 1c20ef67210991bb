"""Loader (and in-tree builder) of the sm_100a CUDA library behind the C-ABI in include/fbsnn_b200.h.

There is no CPU or PyTorch fallback: if the shared library is missing and cannot be built, or no CUDA device is
present when a compute entry point is called, the caller gets a RuntimeError.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.environ.get("FBSNN_LIB_PATH") or os.path.join(_HERE, "libfbsnn_b200.so")   # override: A/B of two builds
SOURCES = ["fbsnn_api.cu", "mc_pricer.cu"]
HEADERS = ["common.cuh", "gemm_simt.cuh", "gemm_tc.cuh", "gemm_tc2.cuh", "gemm_tc16.cuh", "gemm_tc2g.cuh", "gemm_chain.cuh", "kernels.cuh",
           "philox.cuh",
           os.path.join("..", "..", "include", "fbsnn_b200.h")]
NVCC_FLAGS = ["-O3", "-std=c++17", "--threads", "2", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]

_lock = threading.Lock()
_lib = None


HASH_PATH = LIB_PATH + ".srchash"


def _src_hash() -> str:
    """Content hash of everything the library is built from (sources, headers, flags): unlike mtimes it survives a
    copy of the tree (the GPU box gets a snapshot), so a shipped, matching .so is never rebuilt there."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for f in SOURCES + HEADERS:
        p = os.path.join(CSRC, f)
        if os.path.exists(p):
            with open(p, "rb") as fh:
                h.update(f.encode() + b"\0" + fh.read())
    return h.hexdigest()


def _stale() -> bool:
    if not os.path.exists(LIB_PATH) or not os.path.exists(HASH_PATH):
        return True
    with open(HASH_PATH) as fh:
        return fh.read().strip() != _src_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into libfbsnn_b200.so next to this file (nvcc cross-compiles without a GPU).
    Safe when several processes (the ranks of a torchrun job) find the library stale at the same time: one of them builds
    under an exclusive file lock, into a temporary file that is renamed into place; the others wait and find it fresh."""
    import fcntl
    if not force and not _stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libfbsnn_b200.so (and there is no CPU fallback)")
    with open(LIB_PATH + ".lock", "w") as lock_fh:
        fcntl.flock(lock_fh, fcntl.LOCK_EX)
        try:
            if not force and not _stale():       # another process built it while this one waited for the lock
                return LIB_PATH
            tmp = f"{LIB_PATH}.{os.getpid()}.tmp"
            cmd = [nvcc] + NVCC_FLAGS + ["-o", tmp] + SOURCES
            if verbose:
                print(" ".join(cmd))
            r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
            if r.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
            os.replace(tmp, LIB_PATH)
            with open(HASH_PATH + ".tmp", "w") as fh:
                fh.write(_src_hash())
            os.replace(HASH_PATH + ".tmp", HASH_PATH)
        finally:
            fcntl.flock(lock_fh, fcntl.LOCK_UN)
    return LIB_PATH


def _declare(lib):
    from . import spec as S
    c = ctypes
    vp, f32p, i64, u64, sz = c.c_void_p, c.c_void_p, c.c_int64, c.c_uint64, c.c_size_t
    lib.fbsnn_last_error.restype = c.c_char_p
    lib.fbsnn_last_error.argtypes = []
    lib.fbsnn_version.restype = c.c_int
    lib.fbsnn_set_option.restype = c.c_int
    lib.fbsnn_set_option.argtypes = [c.c_char_p, c.c_int]
    lib.fbsnn_launch_count.restype = c.c_longlong
    lib.fbsnn_launch_count.argtypes = []
    lib.fbsnn_dense_timing.restype = None
    lib.fbsnn_dense_timing.argtypes = [c.c_int]
    lib.fbsnn_dense_timing_read.restype = c.c_int
    lib.fbsnn_dense_timing_read.argtypes = [c.POINTER(c.c_double)]
    lib.fbsnn_dense_timing_entry.restype = c.c_char_p
    lib.fbsnn_dense_timing_entry.argtypes = [c.c_int, c.POINTER(c.c_double)]
    lib.fbsnn_debug_gemm.restype = c.c_int
    lib.fbsnn_debug_gemm.argtypes = [c.c_int] * 6 + [f32p, c.c_int, f32p, c.c_int, f32p, c.c_int, vp]
    lib.fbsnn_debug_ws_offset.restype = c.c_int
    lib.fbsnn_debug_ws_offset.argtypes = [c.POINTER(S.FbsnnSpec), i64, c.c_int, c.c_char_p, c.c_int, c.POINTER(i64),
                                          c.POINTER(c.c_int)]
    lib.fbsnn_workspace_bytes.restype = c.c_int
    lib.fbsnn_workspace_bytes.argtypes = [c.POINTER(S.FbsnnSpec), i64, c.c_int, c.POINTER(sz)]
    lib.fbsnn_fetch_minibatch.restype = c.c_int
    lib.fbsnn_fetch_minibatch.argtypes = [c.POINTER(S.FbsnnSpec), c.c_float, i64, i64, u64, u64, f32p, vp, sz,
                                          f32p, f32p, vp]
    lib.fbsnn_net_u.restype = c.c_int
    lib.fbsnn_net_u.argtypes = [c.POINTER(S.FbsnnSpec), f32p, f32p, f32p, i64, vp, sz, f32p, f32p, vp]
    lib.fbsnn_forward.restype = c.c_int
    lib.fbsnn_forward.argtypes = [c.POINTER(S.FbsnnSpec), f32p, f32p, f32p, f32p, i64, i64, vp, sz, f32p, f32p,
                                  f32p, f32p, vp]
    lib.fbsnn_loss_grad.restype = c.c_int
    lib.fbsnn_loss_grad.argtypes = [c.POINTER(S.FbsnnSpec), f32p, f32p, f32p, f32p, f32p, i64, i64, c.c_float,
                                    i64, u64, u64, f32p, vp, sz, f32p, f32p, f32p, f32p, vp]
    lib.fbsnn_loss_grad_step.restype = c.c_int
    lib.fbsnn_loss_grad_step.argtypes = [c.POINTER(S.FbsnnSpec), f32p, f32p, f32p, f32p, f32p, i64, i64, c.c_float,
                                         i64, u64, f32p, vp, vp, sz, f32p, f32p, f32p, vp]
    lib.fbsnn_track_min.restype = c.c_int
    lib.fbsnn_track_min.argtypes = [f32p, f32p, vp, f32p, f32p, i64, f32p, f32p, i64, vp]
    lib.fbsnn_adam_step.restype = c.c_int
    lib.fbsnn_adam_step.argtypes = [c.POINTER(S.FbsnnAdam), f32p, f32p, f32p, f32p, i64, vp, vp]
    lib.fbsnn_peer_allreduce_adam.restype = c.c_int
    lib.fbsnn_peer_allreduce_adam.argtypes = [c.POINTER(S.FbsnnAdam), f32p, vp, c.c_int, c.c_int, f32p, f32p, f32p,
                                              i64, vp, vp]
    lib.fbsnn_peer_buffer_floats.restype = c.c_int
    lib.fbsnn_peer_buffer_floats.argtypes = [i64, c.POINTER(i64), c.POINTER(i64)]
    lib.fbsnn_peer_wait.restype = c.c_int
    lib.fbsnn_peer_wait.argtypes = [f32p, i64, c.c_int, vp, vp]
    lib.fbsnn_train_step.restype = c.c_int
    lib.fbsnn_train_step.argtypes = [c.POINTER(S.FbsnnSpec), c.POINTER(S.FbsnnAdam), f32p, f32p, f32p, f32p, vp,
                                     f32p, f32p, f32p, i64, i64, c.c_float, i64, u64, u64, f32p, vp, sz, f32p,
                                     f32p, f32p, vp]
    lib.mc_launch_count.restype = c.c_longlong
    lib.mc_launch_count.argtypes = []
    lib.mc_scratch_bytes.restype = sz
    lib.mc_scratch_bytes.argtypes = []
    lib.mc_basket_price.restype = c.c_int
    lib.mc_basket_price.argtypes = [c.POINTER(S.McSpec), f32p, f32p, f32p, u64, u64, u64, vp, vp, vp]
    lib.mc_basket_price_delta.restype = c.c_int
    lib.mc_basket_price_delta.argtypes = [c.POINTER(S.McSpec), f32p, f32p, f32p, u64, u64, u64, vp, vp, vp, vp]
    lib.mc_hjb_exact.restype = c.c_int
    lib.mc_hjb_exact.argtypes = [c.c_int32, c.c_int32, f32p, f32p, c.c_float, u64, u64, vp, vp, vp]
    lib.mc_bs_comparator.restype = c.c_int
    lib.mc_bs_comparator.argtypes = [c.c_int32, vp, vp, i64, c.c_int32, c.c_int32, c.c_double, c.c_double, c.c_double,
                                     c.c_double, c.c_int32, vp, vp, vp]
    lib.mc_normal_rate_probe.restype = c.c_int
    lib.mc_normal_rate_probe.argtypes = [u64, u64, vp, c.POINTER(u64), vp]
    lib.mc_generate_paths.restype = c.c_int
    lib.mc_generate_paths.argtypes = [c.POINTER(S.McSpec), f32p, f32p, u64, u64, u64, f32p, vp]


EXPORTS = ["fbsnn_last_error", "fbsnn_version", "fbsnn_set_option", "fbsnn_launch_count", "fbsnn_dense_timing",
           "fbsnn_dense_timing_read", "fbsnn_dense_timing_entry", "fbsnn_debug_gemm", "fbsnn_debug_ws_offset", "fbsnn_workspace_bytes", "fbsnn_fetch_minibatch", "fbsnn_net_u",
           "fbsnn_forward", "fbsnn_loss_grad", "fbsnn_loss_grad_step", "fbsnn_track_min", "fbsnn_adam_step", "fbsnn_peer_buffer_floats", "fbsnn_peer_wait",
           "fbsnn_peer_allreduce_adam", "fbsnn_train_step", "mc_scratch_bytes", "mc_launch_count",
           "mc_basket_price", "mc_basket_price_delta", "mc_hjb_exact", "mc_bs_comparator", "mc_normal_rate_probe", "mc_generate_paths"]


def load():
    """dlopen the library.  A missing or STALE library (a csrc / header file newer than the .so) is rebuilt first when
    nvcc is available, so that an edited source can never silently run old kernels; the ABI version the ctypes struct
    mirrors in spec.py were written against is checked after loading."""
    global _lib
    with _lock:
        if _lib is None:
            from . import spec as S
            have_nvcc = bool(shutil.which("nvcc")) or os.path.exists("/usr/local/cuda/bin/nvcc")
            overridden = bool(os.environ.get("FBSNN_LIB_PATH"))
            if not os.path.exists(LIB_PATH) or (_stale() and have_nvcc and not overridden):
                build()
            lib = ctypes.CDLL(LIB_PATH)
            _declare(lib)
            if lib.fbsnn_version() != S.ABI_VERSION:
                raise RuntimeError(f"{LIB_PATH} reports ABI version {lib.fbsnn_version()}, spec.py expects "
                                   f"{S.ABI_VERSION}: rebuild with _lib.build(force=True)")
            _lib = lib
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().fbsnn_last_error().decode() or f"code {rc}"
        if rc == -2:
            raise NotImplementedError(f"{what}: {msg}")
        if rc == -1:
            raise ValueError(f"{what}: {msg}")
        raise RuntimeError(f"{what}: {msg}")
