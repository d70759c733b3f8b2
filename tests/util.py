"""Shared helpers for the parity tests."""
from __future__ import annotations

import ctypes as C

import numpy as np

from kaldi_fp16_b200 import _lib, gpu
from kaldi_fp16_b200._lib import GemmDesc, K_MAJOR, MN_MAJOR


def rand_f16(rng: np.random.Generator, shape, scale=1.0) -> np.ndarray:
    """fp16-representable float32 values"""
    return (rng.standard_normal(shape).astype(np.float32) * np.float32(scale)).astype(np.float16).astype(np.float32)


def gemm_tol(A: np.ndarray, B: np.ndarray, want: np.ndarray, alpha: float = 1.0, extra_abs=0.0) -> np.ndarray:
    """Elementwise bound on |kernel - oracle|: both accumulate in fp32 (order differs) and round to
    fp16 once -> 1 fp16 ulp of the result + fp32 accumulation noise on sum |a||b|."""
    absprod = np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64)
    return 2.0 ** -10 * np.abs(want) * 1.01 + 2.0 ** -19 * abs(alpha) * absprod + 1e-7 + extra_abs


def assert_close(got, want, tol, what=""):
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    assert got.shape == want.shape, f"{what}: shape {got.shape} vs {want.shape}"
    bad = ~(np.abs(got - want) <= tol)
    if bad.any():
        idx = np.argwhere(bad)
        i = tuple(idx[0])
        raise AssertionError(
            f"{what}: {bad.sum()} / {bad.size} elements out of tolerance; first at {i}: got {got[i]} want {want[i]} "
            f"tol {tol[i] if np.ndim(tol) else tol}; max abs err {np.nanmax(np.abs(got - want))}")


def make_desc(M, N, K, A: gpu.Tensor, B: gpu.Tensor, D: gpu.Tensor, a_major=K_MAJOR, b_major=MN_MAJOR, flags=0,
              alpha=1.0, **kw) -> GemmDesc:
    d = GemmDesc()
    d.M, d.N, d.K = M, N, K
    d.a_major, d.b_major = a_major, b_major
    d.A.ptr, d.A.rows, d.A.cols, d.A.ld, d.A.halo = A.Ptr, A.Rows, A.Cols, A.Cols, 0
    d.B.ptr, d.B.rows, d.B.cols, d.B.ld, d.B.halo = B.Ptr, B.Rows, B.Cols, B.Cols, 0
    d.groups, d.kslabs, d.kslab_len = 1, 1, K
    d.D[0] = D.Ptr
    d.ldd = D.Cols
    d.flags = flags
    d.alpha = alpha
    for k, v in kw.items():
        setattr(d, k, v)
    return d


def run_desc(handle, d: GemmDesc):
    lib = _lib.load()
    rc = lib.kfp16_gemm_ex(handle.ptr, C.byref(d))
    if rc != 0:
        raise RuntimeError(f"kfp16_gemm_ex failed: {_lib.last_error()}")
    gpu.Sync()
