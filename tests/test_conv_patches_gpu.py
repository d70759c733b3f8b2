"""Patch gather / scatter (kfp16_im2col, kfp16_col2im) against a numpy restatement of the reference's CPU im2col
(internal/nnet/forward.go:429-455: P[(t*Hout+ho), off*Fin+f] = X[t+dt, (ho*sub+dh)*Fin+f], zero outside) and its adjoint,
on the padded minibatch layout (per-sequence zero padding in time).  Covers the shared-memory staged kernels for few input
filters (the first conv layer: 6 feature maps, K = 54 -> 64) and the generic kernels."""
from __future__ import annotations

import ctypes as C

import numpy as np
import pytest

from kaldi_fp16_b200 import gpu

pytestmark = pytest.mark.gpu

TAPS9 = [(dt, dh) for dt in (-1, 0, 1) for dh in (-1, 0, 1)]


def ref_patches(x, n_seq, L, halo, hin, hout, sub, fin, Kp, taps):
    """x: [n_seq*(L+2*halo), hin*fin] -> P [rows*hout, Kp] (float32)"""
    blk = L + 2 * halo
    rows = n_seq * blk
    xs = x.reshape(n_seq, blk, hin, fin)
    P = np.zeros((n_seq, blk, hout, Kp), np.float32)
    for k, (dt, dh) in enumerate(taps):
        for ho in range(hout):
            hs = ho * sub + dh
            if hs < 0 or hs >= hin:
                continue
            for t in range(L):
                ts = t + dt
                if 0 <= ts < L:
                    P[:, halo + t, ho, k * fin:(k + 1) * fin] = xs[:, halo + ts, hs, :]
    return P.reshape(rows * hout, Kp)


@pytest.mark.parametrize("n_seq,L,halo,hin,sub,fin,taps", [
    (3, 11, 3, 40, 1, 6, TAPS9),                 # the benchmark's cnn1 geometry: K = 54 -> 64 (staged kernels)
    (2, 9, 1, 16, 1, 3, TAPS9),                  # K = 27 -> 32
    (5, 2, 3, 8, 1, 2, [(-1, 0), (0, 1), (1, -1), (0, 0)]),   # sequences shorter than the frame group of a block
    (2, 13, 2, 16, 2, 3, TAPS9),                 # height subsampling: generic kernels
    (2, 10, 1, 8, 1, 32, TAPS9),                 # 16-byte generic path
    (1, 7, 3, 40, 1, 6, [(-3, 0), (0, 0), (3, 0)]),            # wide time span
])
def test_im2col_col2im_against_numpy(lib, handle, n_seq, L, halo, hin, sub, fin, taps):
    rng = np.random.default_rng(n_seq * 100 + L + fin)
    hout = hin // sub
    blk = L + 2 * halo
    rows = n_seq * blk
    K = len(taps) * fin
    Kp = (K + 15) // 16 * 16
    x = rng.standard_normal((rows, hin * fin)).astype(np.float16).astype(np.float32)      # halo rows hold data too: must be ignored
    tx = gpu.TensorFromFP16(x)
    tP = gpu.TensorFromFP16(np.full((rows * hout, Kp), 7.0, np.float32))                  # every element must be overwritten
    dt = (C.c_int * len(taps))(*[t[0] for t in taps])
    dh = (C.c_int * len(taps))(*[t[1] for t in taps])
    assert lib.kfp16_im2col(handle.ptr, tx.Ptr, tP.Ptr, Kp, n_seq, L, halo, hin, hout, sub, fin, len(taps), dt, dh) == 0
    want = ref_patches(x, n_seq, L, halo, hin, hout, sub, fin, Kp, taps)
    got = tP.ToFP32()
    assert np.array_equal(got, want), f"im2col: {np.sum(got != want)} elements differ"
    # adjoint: <P(x), dP> == <x, col2im(dP)> holds exactly in exact arithmetic; compare with the explicit scatter in fp32
    dP = rng.standard_normal((rows * hout, Kp)).astype(np.float16).astype(np.float32)
    tdP = gpu.TensorFromFP16(dP)
    tdx = gpu.TensorFromFP16(np.full((rows, hin * fin), 7.0, np.float32))
    assert lib.kfp16_col2im(handle.ptr, tdP.Ptr, Kp, tdx.Ptr, n_seq, L, halo, hin, hout, sub, fin, len(taps), dt, dh) == 0
    dps = dP.reshape(n_seq, blk, hout, Kp)
    wx = np.zeros((n_seq, blk, hin, fin), np.float32)
    for k, (d_t, d_h) in enumerate(taps):       # tap order = the kernels' accumulation order
        for ho in range(hout):
            hs = ho * sub + d_h
            if hs < 0 or hs >= hin:
                continue
            for t in range(L):
                ts = t + d_t
                if 0 <= ts < L:
                    wx[:, halo + ts, hs, :] += dps[:, halo + t, ho, k * fin:(k + 1) * fin]
    want_dx = wx.reshape(rows, hin * fin).astype(np.float16).astype(np.float32)
    got_dx = tdx.ToFP32()
    # one fp16 rounding of an fp32 sum of <= 9 terms: equal up to the summation order of ties
    assert np.abs(got_dx - want_dx).max() <= 2.0 ** -9 * max(1.0, np.abs(want_dx).max())
    blkrow = np.arange(rows) % blk
    halo_rows = (blkrow < halo) | (blkrow >= halo + L)
    assert not got_dx[halo_rows].any(), "halo rows of the input gradient must be zero"
    for t in (tx, tP, tdP, tdx):
        t.Free()
