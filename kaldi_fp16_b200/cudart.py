"""Minimal ctypes view of the CUDA runtime for host-side plumbing: events, streams, sync.
(Timing and stream handles only -- no compute goes through here.)"""
from __future__ import annotations

import ctypes as C

_rt = None


def rt() -> C.CDLL:
    global _rt
    if _rt is None:
        last = None
        for name in ("libcudart.so", "libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
            try:
                _rt = C.CDLL(name)
                break
            except OSError as e:  # pragma: no cover
                last = e
        if _rt is None:
            raise ImportError(f"CUDA runtime not found: {last}")
        _rt.cudaGetErrorString.restype = C.c_char_p
    return _rt


def check(code: int, what: str) -> None:
    if code != 0:
        raise RuntimeError(f"{what}: {rt().cudaGetErrorString(code).decode()}")


def device_count() -> int:
    n = C.c_int(0)
    return n.value if rt().cudaGetDeviceCount(C.byref(n)) == 0 else 0


def set_device(i: int) -> None:
    check(rt().cudaSetDevice(i), "cudaSetDevice")


def synchronize() -> None:
    check(rt().cudaDeviceSynchronize(), "cudaDeviceSynchronize")


class Stream:
    def __init__(self, non_blocking: bool = True):
        self.handle = C.c_void_p()
        check(rt().cudaStreamCreateWithFlags(C.byref(self.handle), 1 if non_blocking else 0), "cudaStreamCreate")

    @property
    def ptr(self) -> int:
        return self.handle.value or 0

    def synchronize(self) -> None:
        check(rt().cudaStreamSynchronize(self.handle), "cudaStreamSynchronize")

    def destroy(self) -> None:
        if self.handle:
            rt().cudaStreamDestroy(self.handle)
            self.handle = C.c_void_p()


class Event:
    def __init__(self):
        self.handle = C.c_void_p()
        check(rt().cudaEventCreate(C.byref(self.handle)), "cudaEventCreate")

    def record(self, stream_ptr: int = 0) -> None:
        check(rt().cudaEventRecord(self.handle, C.c_void_p(stream_ptr)), "cudaEventRecord")

    def synchronize(self) -> None:
        check(rt().cudaEventSynchronize(self.handle), "cudaEventSynchronize")

    def elapsed_ms(self, end: "Event") -> float:
        ms = C.c_float(0)
        check(rt().cudaEventElapsedTime(C.byref(ms), self.handle, end.handle), "cudaEventElapsedTime")
        return ms.value

    def destroy(self) -> None:
        if self.handle:
            rt().cudaEventDestroy(self.handle)
            self.handle = C.c_void_p()


def memset(ptr: int, value: int, nbytes: int, stream_ptr: int = 0) -> None:
    check(rt().cudaMemsetAsync(C.c_void_p(ptr), value, C.c_size_t(nbytes), C.c_void_p(stream_ptr)), "cudaMemsetAsync")
