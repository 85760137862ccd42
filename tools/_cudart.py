"""Minimal ctypes face of the CUDA runtime for the torch-free tools (device memory, copies, events on the legacy default
stream).  The library links its own static cudart; both runtimes share the device's primary context, so pointers and the
default stream are common."""
import ctypes as C
import glob
import os
import sys

import numpy as np

_rt = None


def rt():
    global _rt
    if _rt is None:
        cands = ["libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so.12",
                 "/usr/local/cuda/targets/x86_64-linux/lib/libcudart.so.12"]
        cands += glob.glob(os.path.join(sys.prefix, "lib/python*/site-packages/nvidia/cuda_runtime/lib/libcudart.so.12"))
        for c in cands:
            try:
                _rt = C.CDLL(c)
                break
            except OSError:
                continue
        if _rt is None:
            raise SystemExit("libcudart.so.12 not found")
    return _rt


def ck(e, what="cuda"):
    if e != 0:
        raise SystemExit(f"{what}: cuda error {e}")


def malloc(nbytes: int) -> C.c_void_p:
    p = C.c_void_p()
    ck(rt().cudaMalloc(C.byref(p), C.c_size_t(nbytes)), "cudaMalloc")
    return p


def memset(p, value: int, nbytes: int):
    ck(rt().cudaMemset(p, value, C.c_size_t(nbytes)), "cudaMemset")


def h2d(p, arr: np.ndarray):
    a = np.ascontiguousarray(arr)
    ck(rt().cudaMemcpy(p, a.ctypes.data_as(C.c_void_p), C.c_size_t(a.nbytes), 1), "H2D")


def d2h(p, nbytes: int, dtype=np.uint8) -> np.ndarray:
    out = np.empty(nbytes // np.dtype(dtype).itemsize, dtype)
    ck(rt().cudaMemcpy(out.ctypes.data_as(C.c_void_p), p, C.c_size_t(nbytes), 2), "D2H")
    return out


def to_device(arr: np.ndarray) -> C.c_void_p:
    a = np.ascontiguousarray(arr)
    p = malloc(a.nbytes)
    h2d(p, a)
    return p


def sync():
    ck(rt().cudaDeviceSynchronize(), "sync")


class Timer:
    """CUDA events on the legacy default stream."""

    def __init__(self):
        self.e0, self.e1 = C.c_void_p(), C.c_void_p()
        ck(rt().cudaEventCreate(C.byref(self.e0)))
        ck(rt().cudaEventCreate(C.byref(self.e1)))

    def start(self):
        ck(rt().cudaEventRecord(self.e0, None))

    def stop_ms(self) -> float:
        ck(rt().cudaEventRecord(self.e1, None))
        ck(rt().cudaEventSynchronize(self.e1))
        ms = C.c_float()
        ck(rt().cudaEventElapsedTime(C.byref(ms), self.e0, self.e1))
        return ms.value
