"""Second operator surface (go/kaldibridge): kaldi_* opaque tensors and launch_* conv / batch-norm entry
points, checked on the GPU against the REFERENCE's own compiled library (oracle/_ref, cuBLAS + its scalar
kernels) on the same inputs, and against float64 numpy where the reference kernel is known-buggy."""
import ctypes as C

import numpy as np
import pytest

from kaldi_fp16_b200 import gpu
from oracle import kaldi_oracle as O
from tests.refbind import RefBuf

pytestmark = pytest.mark.gpu


def f16(rng, shape, scale=1.0):
    return (rng.standard_normal(shape).astype(np.float32) * np.float32(scale)).astype(np.float16)


def dev(x16):
    return gpu.TensorFromBits(np.ascontiguousarray(x16).view(np.uint16).reshape(1, -1))


def host(t, shape):
    return t.ToBits().view(np.float16).astype(np.float32).reshape(shape)


def refdev(reflib, x16):
    return RefBuf(reflib, np.ascontiguousarray(x16).view(np.uint16))


@pytest.mark.parametrize("transA,transB", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_kaldi_gemm_vs_reference(lib, reflib, transA, transB):
    M, N, K = 96, 72, 128
    rng = np.random.default_rng(3 + transA * 2 + transB)
    A = f16(rng, (K, M) if transA else (M, K)).astype(np.float32)
    B = f16(rng, (N, K) if transB else (K, N), 0.1).astype(np.float32)
    C0 = f16(rng, (M, N)).astype(np.float32)
    outs = []
    for L in (lib, reflib):
        h = L.kaldi_cublas_create()
        assert h
        tA, tB, tC = L.kaldi_tensor_create(*A.shape), L.kaldi_tensor_create(*B.shape), L.kaldi_tensor_create(M, N)
        L.kaldi_tensor_copy_from_host_fp32(tA, A.ctypes.data, A.size)
        L.kaldi_tensor_copy_from_host_fp32(tB, B.ctypes.data, B.size)
        L.kaldi_tensor_copy_from_host_fp32(tC, C0.ctypes.data, C0.size)
        L.kaldi_gemm(h, tA, tB, tC, 0.5, 1.0, transA, transB)
        out = np.empty((M, N), np.float32)
        L.kaldi_tensor_copy_to_host_fp32(tC, out.ctypes.data, out.size)
        outs.append(out)
        for t in (tA, tB, tC):
            L.kaldi_tensor_free(t)
        L.kaldi_cublas_destroy(h)
    want = 0.5 * ((A.T if transA else A).astype(np.float64) @ (B.T if transB else B).astype(np.float64)) + C0
    # the reference accumulates in FP16 (cublasHgemm): it is the LESS accurate of the two
    assert O.max_err_vs_scale(outs[0], want) < 2e-3
    assert O.max_err_vs_scale(outs[1], want) < 8e-3
    assert O.max_err_vs_scale(outs[0], outs[1]) < 8e-3


def test_kaldi_tensor_api_and_elementwise(lib):
    assert lib.kaldi_tensor_rows(None) == 0 and lib.kaldi_tensor_size(None) == 0
    t = lib.kaldi_tensor_ones(5, 8)
    assert (lib.kaldi_tensor_rows(t), lib.kaldi_tensor_cols(t), lib.kaldi_tensor_size(t)) == (5, 8, 40)
    out = np.zeros(40, np.float32)
    lib.kaldi_tensor_copy_to_host_fp32(t, out.ctypes.data, 40)
    assert (out == 1).all()
    x = np.linspace(-2, 2, 40, dtype=np.float32)
    lib.kaldi_tensor_copy_from_host_fp32(t, x.ctypes.data, 40)
    lib.kaldi_scale(t, 2.0)
    lib.kaldi_relu(t)
    z = lib.kaldi_tensor_zeros(5, 8)
    lib.kaldi_add(z, t)
    lib.kaldi_tensor_copy_to_host_fp32(z, out.ctypes.data, 40)
    assert np.allclose(out, np.maximum(2 * x.astype(np.float16).astype(np.float32), 0), atol=2e-3)
    lib.kaldi_softmax(z)
    lib.kaldi_tensor_copy_to_host_fp32(z, out.ctypes.data, 40)
    assert np.allclose(out.reshape(5, 8).sum(1), 1.0, atol=5e-3)
    lib.kaldi_tensor_free(t)
    lib.kaldi_tensor_free(z)


@pytest.mark.parametrize("stride,padding,dilation,K", [(1, 1, 1, 3), (2, 2, 2, 3), (1, 0, 1, 5)])
@pytest.mark.parametrize("Cin,Cout", [(16, 24), (5, 7), (16, 10), (3, 8)])
def test_conv1d_forward_and_backward_vs_reference(lib, reflib, stride, padding, dilation, K, Cin, Cout):
    """(16, 24): the tcgen05 / TMA lowering; the other channel counts (Cin*K or Cout not a multiple of 8) must be accepted
    like the reference's kernels accept them (cnn_kernels.cu:19-65) and run the dense SIMT lowering"""
    B, T = 3, 50
    Tout = (T + 2 * padding - dilation * (K - 1) - 1) // stride + 1
    rng = np.random.default_rng(K * 10 + stride)
    x, w, b = f16(rng, (B, T, Cin)), f16(rng, (Cout, Cin, K), 0.2), f16(rng, (Cout,), 0.1)
    go = f16(rng, (B, Tout, Cout))
    # ---- forward: ours vs the reference's scalar kernel (cnn_kernels.cu:19-63)
    dx, dw, db, dy = dev(x), dev(w), dev(b), gpu.ZeroTensor(1, B * Tout * Cout)
    lib.launch_conv1d_forward_fp16(dx.Ptr, dw.Ptr, db.Ptr, dy.Ptr, B, T, Cin, Cout, K, stride, padding, dilation, None)
    gpu.Sync()
    got = host(dy, (B, Tout, Cout))
    rx, rw, rb, ry = refdev(reflib, x), refdev(reflib, w), refdev(reflib, b), RefBuf(reflib, np.zeros((B, Tout, Cout), np.uint16))
    reflib.launch_conv1d_forward_fp16(rx.ptr, rw.ptr, rb.ptr, ry.ptr, B, T, Cin, Cout, K, stride, padding, dilation, None)
    reflib.bridge_gpu_sync()
    ref = ry.f32()
    # float64 definition
    xp = np.zeros((B, T + 2 * padding, Cin))
    xp[:, padding:padding + T] = x
    want = np.zeros((B, Tout, Cout))
    for k in range(K):
        seg = xp[:, k * dilation: k * dilation + (Tout - 1) * stride + 1: stride]
        want += seg @ w[:, :, k].astype(np.float64).T
    want += b.astype(np.float64)
    assert O.max_err_vs_scale(ref, want) < 2e-3
    assert O.max_err_vs_scale(got, want) < 2e-3
    assert np.mean(got == ref) > 0.95 and O.max_err_vs_scale(got, ref) < 2e-3
    # ---- backward (cnn_kernels.cu:126-229 semantics) vs the float64 definition
    dgo = dev(go)
    dgi, dgw, dgb = gpu.ZeroTensor(1, B * T * Cin), gpu.ZeroTensor(1, Cout * Cin * K), gpu.ZeroTensor(1, Cout)
    lib.launch_conv1d_backward_fp16(dx.Ptr, dgo.Ptr, dw.Ptr, dgi.Ptr, dgw.Ptr, dgb.Ptr, B, T, Cin, Cout, K, stride, padding, dilation, None)
    gpu.Sync()
    gi, gw, gb = host(dgi, (B, T, Cin)), host(dgw, (Cout, Cin, K)), host(dgb, (Cout,))
    want_gi = np.zeros((B, T + 2 * padding, Cin))
    want_gw = np.zeros((Cout, Cin, K))
    for k in range(K):
        sl = slice(k * dilation, k * dilation + (Tout - 1) * stride + 1, stride)
        want_gi[:, sl] += go.astype(np.float64) @ w[:, :, k].astype(np.float64)
        want_gw[:, :, k] = np.einsum("bto,bti->oi", go.astype(np.float64), xp[:, sl])
    want_gi = want_gi[:, padding:padding + T]
    assert O.max_err_vs_scale(gi, want_gi) < 3e-3
    assert O.max_err_vs_scale(gw, want_gw) < 3e-3
    assert O.max_err_vs_scale(gb, go.astype(np.float64).sum((0, 1))) < 3e-3
    # (the reference's own launch_conv1d_backward_fp16 is NOT run: its weight-gradient kernel does a float atomicAdd on
    #  half-aligned addresses and faults with a misaligned-address error on B200, poisoning the CUDA context)


def test_pointwise_conv_vs_reference(lib, reflib):
    B, T, Cin, Cout = 4, 33, 64, 40
    rng = np.random.default_rng(8)
    x, w, b = f16(rng, (B, T, Cin)), f16(rng, (Cout, Cin), 0.1), f16(rng, (Cout,), 0.1)
    dx, dw, db, dy = dev(x), dev(w), dev(b), gpu.ZeroTensor(1, B * T * Cout)
    lib.launch_pointwise_conv1d_fp16(dx.Ptr, dw.Ptr, db.Ptr, dy.Ptr, B, T, Cin, Cout, None)
    gpu.Sync()
    rx, rw, rb, ry = refdev(reflib, x), refdev(reflib, w), refdev(reflib, b), RefBuf(reflib, np.zeros((B, T, Cout), np.uint16))
    reflib.launch_pointwise_conv1d_fp16(rx.ptr, rw.ptr, rb.ptr, ry.ptr, B, T, Cin, Cout, None)
    reflib.bridge_gpu_sync()
    got, ref = host(dy, (B, T, Cout)), ry.f32()
    assert np.mean(got == ref) > 0.97 and O.max_err_vs_scale(got, ref) < 1e-3


@pytest.mark.parametrize("training", [True, False])
def test_batchnorm1d_vs_reference(lib, reflib, training):
    B, T, Cc = 4, 60, 48
    rng = np.random.default_rng(12)
    x = f16(rng, (B, T, Cc), 2.0)
    gamma, beta = f16(rng, (Cc,), 0.5) + np.float16(1), f16(rng, (Cc,), 0.1)
    rm, rv = f16(rng, (Cc,), 0.1), (rng.random(Cc) + 0.5).astype(np.float16)
    ours = [dev(a) for a in (x, gamma, beta, rm, rv)] + [gpu.ZeroTensor(1, x.size), gpu.ZeroTensor(1, Cc), gpu.ZeroTensor(1, Cc)]
    lib.launch_batchnorm1d_forward_fp16(ours[0].Ptr, ours[1].Ptr, ours[2].Ptr, ours[3].Ptr, ours[4].Ptr, ours[5].Ptr, ours[6].Ptr,
                                        ours[7].Ptr, B, T, Cc, 0.1, 1e-5, training, None)
    gpu.Sync()
    refs = [refdev(reflib, a) for a in (x, gamma, beta, rm, rv)] + [RefBuf(reflib, np.zeros(s, np.uint16)) for s in ((B, T, Cc), (Cc,), (Cc,))]
    reflib.launch_batchnorm1d_forward_fp16(refs[0].ptr, refs[1].ptr, refs[2].ptr, refs[3].ptr, refs[4].ptr, refs[5].ptr, refs[6].ptr,
                                           refs[7].ptr, B, T, Cc, 0.1, 1e-5, training, None)
    reflib.bridge_gpu_sync()
    y, ry = host(ours[5], (B, T, Cc)), refs[5].f32()
    assert O.max_err_vs_scale(y, ry) < 2e-3
    if training:
        for i, name in [(3, "running_mean"), (4, "running_var"), (6, "save_mean"), (7, "save_invstd")]:
            assert O.max_err_vs_scale(host(ours[i], (Cc,)), refs[i].f32()) < 2e-3, name
        xs = x.astype(np.float64).reshape(-1, Cc)
        assert np.allclose(host(ours[6], (Cc,)), xs.mean(0), atol=2e-3)


def test_maxpool_vs_reference(lib, reflib):
    B, T, Cc, K, S = 2, 41, 24, 3, 2
    Tout = (T - K) // S + 1
    rng = np.random.default_rng(4)
    x, go = f16(rng, (B, T, Cc)), f16(rng, (B, Tout, Cc))
    dx, dy, di = dev(x), gpu.ZeroTensor(1, B * Tout * Cc), gpu.DeviceF32(n=B * Tout * Cc)
    lib.launch_maxpool1d_forward_fp16(dx.Ptr, dy.Ptr, di.Ptr, B, T, Cc, K, S, None)
    gpu.Sync()
    win = np.stack([x[:, k: k + (Tout - 1) * S + 1: S] for k in range(K)], 0).astype(np.float32)
    assert np.array_equal(host(dy, (B, Tout, Cc)), win.max(0))
    idx = di.ToHost().view(np.int32).reshape(B, Tout, Cc)
    assert np.array_equal(idx, win.argmax(0) + np.arange(Tout)[None, :, None] * S)
    dgo, dgi = dev(go), gpu.ZeroTensor(1, B * T * Cc)
    lib.launch_maxpool1d_backward_fp16(dgo.Ptr, di.Ptr, dgi.Ptr, B, T, Tout, Cc, None)
    gpu.Sync()
    want = np.zeros((B, T, Cc))
    for b in range(B):
        for t in range(Tout):
            for c in range(Cc):
                want[b, idx[b, t, c], c] += float(go[b, t, c])
    assert O.max_err_vs_scale(host(dgi, (B, T, Cc)), want) < 2e-3
