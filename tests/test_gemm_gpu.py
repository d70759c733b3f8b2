"""GPU parity of the tcgen05 GEMM against the numpy oracle (oracle/kaldi_oracle.py :: gemm, which
restates ops_gemm, /root/reference/cpp/cuda/ops.cu:366-400) and, when present, against the
reference's own compiled library (cublasGemmEx) on the same inputs.  Tolerance: FP16 storage with
FP32 accumulation -> 1 fp16 ulp of the result + fp32 accumulation noise (tests/util.py::gemm_tol)."""
import ctypes as C

import numpy as np
import pytest

from kaldi_fp16_b200 import _lib, gpu
from kaldi_fp16_b200._lib import (EPI_BETA, EPI_BIAS, EPI_BN, EPI_DROPOUT, EPI_MASK, EPI_REF_ROUND, EPI_RELU, EPI_RESID,
                                  K_MAJOR, MN_MAJOR)
from oracle import kaldi_oracle as O
from tests.util import assert_close, gemm_tol, make_desc, rand_f16, run_desc

pytestmark = pytest.mark.gpu

# (M, N, K): layer shapes of SURVEY 8(d) scaled down in M, plus ragged / tiny cases
SHAPES = [
    (128, 64, 64), (128, 256, 128), (256, 160, 320), (300, 1536, 320), (1000, 160, 3072), (777, 256, 1536),
    (150, 6016, 256), (64, 200, 104), (1, 8, 8), (129, 72, 88), (515, 3080, 256), (2400, 40, 40), (1234, 128, 1152),
]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_ops_gemm_matches_oracle(handle, M, N, K):
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    A, B = rand_f16(rng, (M, K)), rand_f16(rng, (K, N), 0.1)
    tA, tB, tC = gpu.TensorFromFP16(A), gpu.TensorFromFP16(B), gpu.ZeroTensor(M, N)
    gpu.GEMM(handle, M, N, K, 1.0, tA, tB, 0.0, tC)
    gpu.Sync()
    want = O.gemm(A, B)
    assert_close(tC.ToFP32(), want, gemm_tol(A, B, want), f"ops_gemm {M}x{N}x{K}")
    for t in (tA, tB, tC):
        t.Free()


@pytest.mark.parametrize("M,N,K", [(256, 160, 320), (777, 256, 1536), (150, 6016, 256), (129, 72, 88)])
def test_ops_gemm_matches_reference_library(handle, reflib, M, N, K):
    """same inputs through the reference's cublasGemmEx path and through the new kernel"""
    from tests.refbind import ref_half

    rng = np.random.default_rng(11 + M + N + K)
    A, B, C0 = rand_f16(rng, (M, K)), rand_f16(rng, (K, N), 0.1), rand_f16(rng, (M, N))
    for alpha, beta in ((1.0, 0.0), (0.5, 1.0), (2.0, -0.25)):
        rh = reflib.ops_cublas_create()
        rA, rB, rC = ref_half(reflib, A), ref_half(reflib, B), ref_half(reflib, C0)
        assert reflib.ops_gemm(rh, M, N, K, alpha, rA.ptr, K, rB.ptr, N, beta, rC.ptr, N) == 0
        reflib.bridge_gpu_sync()
        ref = rC.f32()
        tA, tB, tC = gpu.TensorFromFP16(A), gpu.TensorFromFP16(B), gpu.TensorFromFP16(C0)
        gpu.GEMM(handle, M, N, K, alpha, tA, tB, beta, tC)
        gpu.Sync()
        got = tC.ToFP32()
        want = O.gemm(A, B, alpha, beta, C0)
        tol = gemm_tol(A, B, want, alpha) + 2.0 ** -10 * np.abs(beta * C0)
        if beta != 0.0:   # cuBLAS rounds alpha*acc to fp16 before adding beta*C (see oracle.gemm): one more rounding of the product
            tol = tol + 2.0 ** -10 * np.abs(alpha * O.gemm_f64(A, B))
        assert_close(ref, want, tol, f"oracle vs reference cuBLAS a={alpha} b={beta}")   # pins the oracle
        assert_close(got, ref, tol, f"kernel vs reference cuBLAS a={alpha} b={beta}")
        for b in (rA, rB, rC):
            b.free()
        for t in (tA, tB, tC):
            t.Free()
        reflib.ops_cublas_destroy(rh)


@pytest.mark.parametrize("transA,transB", [(0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(256, 128, 192), (300, 160, 1000), (1536, 320, 2000), (136, 1536, 520)])
def test_gemm_operand_majors(handle, lib, M, N, K, transA, transB):
    """dgrad (NT) and wgrad (TN) without transpose kernels (backward_ops.go:162-225)"""
    rng = np.random.default_rng(5 + M + N + K + transA * 2 + transB)
    A = rand_f16(rng, (K, M) if transA else (M, K))
    B = rand_f16(rng, (N, K) if transB else (K, N), 0.1)
    tA, tB, tC = gpu.TensorFromFP16(A), gpu.TensorFromFP16(B), gpu.ZeroTensor(M, N)
    assert lib.kfp16_gemm(handle.ptr, M, N, K, 1.0, tA.Ptr, transA, tB.Ptr, transB, 0.0, tC.Ptr) == 0, _lib.last_error()
    gpu.Sync()
    want = O.gemm(A, B, transA=bool(transA), transB=bool(transB))
    a = A.T if transA else A
    b = B.T if transB else B
    assert_close(tC.ToFP32(), want, gemm_tol(a, b, want), f"gemm tA={transA} tB={transB}")
    for t in (tA, tB, tC):
        t.Free()


@pytest.mark.parametrize("bn", [64, 128, 160, 256])
def test_gemm_tile_widths_multi_tile_per_cta(handle, lib, bn):
    """force each tile width and a small persistent grid so every CTA walks several tiles"""
    M, N, K = 1100, bn, 448
    rng = np.random.default_rng(bn)
    A, B = rand_f16(rng, (M, K)), rand_f16(rng, (K, N), 0.1)
    tA, tB, tC = gpu.TensorFromFP16(A), gpu.TensorFromFP16(B), gpu.ZeroTensor(M, N)
    lib.kfp16_ctx_set_max_ctas(handle.ptr, 3)
    try:
        d = make_desc(M, N, K, tA, tB, tC, force_bn=bn)
        run_desc(handle, d)
    finally:
        lib.kfp16_ctx_set_max_ctas(handle.ptr, 0)
    want = O.gemm(A, B)
    assert_close(tC.ToFP32(), want, gemm_tol(A, B, want), f"bn={bn}")
    for t in (tA, tB, tC):
        t.Free()


def test_fused_epilogue_affine_relu_bn_bypass(handle, lib):
    """TDNN-F affine tail (forward.go:640-693): h(h(h(relu(h(h(acc)+b)))*s+t) ... + 0.66*X with the
    reference's intermediate FP16 stores (EPI_REF_ROUND) -> must equal the unfused op sequence."""
    M, N, K = 520, 256, 320
    rng = np.random.default_rng(3)
    A, W = rand_f16(rng, (M, K)), rand_f16(rng, (K, N), 0.08)
    bias = rand_f16(rng, (N,), 0.1)
    X = rand_f16(rng, (M, N))
    mean, var = rng.standard_normal(N).astype(np.float32) * 0.1, (rng.random(N).astype(np.float32) + 0.5)
    gamma, beta = (rng.random(N).astype(np.float32) + 0.5), rng.standard_normal(N).astype(np.float32) * 0.1
    eps = 1e-3
    # oracle: the reference's op sequence
    z = O.gemm(A, W)
    z = O.add_bias(z, bias)
    z = O.relu(z)
    z_relu = z
    z = O.batchnorm_forward(z, mean, var, gamma, beta, eps)
    want = O.add_scaled(z, X, 0.66, 1.0)
    # fused
    tA, tW, tX, tD = gpu.TensorFromFP16(A), gpu.TensorFromFP16(W), gpu.TensorFromFP16(X), gpu.ZeroTensor(M, N)
    tb = gpu.TensorFromFP16(bias.reshape(1, -1))
    dm, dv, dg, dbt = gpu.DeviceF32(mean), gpu.DeviceF32(var), gpu.DeviceF32(gamma), gpu.DeviceF32(beta)
    sc, sh = gpu.DeviceF32(n=N), gpu.DeviceF32(n=N)
    assert lib.kfp16_bn_fold(handle.ptr, dm.Ptr, dv.Ptr, dg.Ptr, dbt.Ptr, eps, 1.0, N, sc.Ptr, sh.Ptr) == 0
    mask_ld = (N + 31) // 32
    mask = gpu.DeviceF32(n=M * mask_ld)
    d = make_desc(M, N, K, tA, tW, tD, flags=EPI_BIAS | EPI_RELU | EPI_BN | EPI_RESID | EPI_REF_ROUND | EPI_MASK,
                  bias=tb.Ptr, bn_scale=sc.Ptr, bn_shift=sh.Ptr, res_scale=0.66, mask_out=mask.Ptr, mask_ld=mask_ld,
                  ldr=N)
    d.R[0] = tX.Ptr
    run_desc(handle, d)
    got = tD.ToFP32()
    # scale/shift folding differs from gamma*(x-mean)/sqrt(var+eps)+beta by fp32 rounding: 1 extra ulp
    tol = gemm_tol(A, W, want) * 3 + 2.0 ** -9 * np.abs(want) + 2e-3
    assert_close(got, want, tol, "fused affine epilogue")
    bits = mask.ToHost().view(np.uint32).reshape(M, mask_ld)
    got_mask = ((bits[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(M, -1)[:, :N].astype(bool)
    want_mask = z_relu > 0
    # elements whose pre-activation is within rounding of 0 may flip
    near0 = np.abs(O.add_bias(O.gemm(A, W), bias)) < 1e-3
    assert ((got_mask == want_mask) | near0).all()


def test_spliced_k_slabs(handle, lib):
    """TDNN-F splice as two K-slabs with shifted row coordinates (forward.go:699-790):
    Y = [X(t-s) | X(t)] * W  without materialising the spliced matrix; rows outside [0,T) read the
    replicated halo rows of the padded buffer."""
    T, D, N, s, halo = 300, 192, 160, 3, 3
    rng = np.random.default_rng(9)
    X = rand_f16(rng, (T, D))
    W = rand_f16(rng, (2 * D, N), 0.08)
    Xp = np.concatenate([np.repeat(X[:1], halo, 0), X, np.repeat(X[-1:], halo, 0)], 0)
    idx = np.clip(np.arange(T) - s, 0, T - 1)
    want = O.gemm(np.concatenate([X[idx], X], 1), W)
    tXp, tW, tD = gpu.TensorFromFP16(Xp), gpu.TensorFromFP16(W), gpu.ZeroTensor(T, N)
    d = make_desc(T, N, 2 * D, tXp, tW, tD)
    d.A.ptr = tXp.Ptr + halo * D * 2
    d.A.rows, d.A.halo = T, halo
    d.kslabs, d.kslab_len = 2, D
    d.a_row_off[0][0], d.a_row_off[0][1] = -s, 0
    d.b_row_off[0][0], d.b_row_off[0][1] = 0, D
    run_desc(handle, d)
    S = np.concatenate([X[idx], X], 1)
    assert_close(tD.ToFP32(), want, gemm_tol(S, W, want), "spliced GEMM")


def test_split_k_fp32_accumulate(handle, lib):
    """weight-gradient shape: tiny output, long reduction -> split-K partials added in FP32"""
    M, N, K = 320, 256, 5000
    rng = np.random.default_rng(21)
    At, B = rand_f16(rng, (K, M)), rand_f16(rng, (K, N), 0.05)
    tA, tB = gpu.TensorFromFP16(At), gpu.TensorFromFP16(B)
    ws = gpu.DeviceF32(n=M * N)
    tD = gpu.ZeroTensor(M, N)
    d = make_desc(M, N, K, tA, tB, tD, a_major=MN_MAJOR, b_major=MN_MAJOR, split_k=8, ws_ld=N)
    d.ws[0] = ws.Ptr
    run_desc(handle, d)
    got = ws.ToHost().reshape(M, N)
    want = (At.T.astype(np.float64) @ B.astype(np.float64))
    absprod = np.abs(At.T).astype(np.float64) @ np.abs(B).astype(np.float64)
    assert_close(got, want, 2.0 ** -19 * absprod + 1e-6, "split-K fp32")


def test_gemm_errors_are_reported(handle, lib):
    """error convention of ops.h: -1 + message from ops_last_error (ops.cu:12-20,393-397)"""
    lib.ops_clear_error()
    assert lib.ops_gemm(None, 8, 8, 8, 1.0, None, 8, None, 8, 0.0, None, 8) == -1
    assert lib.ops_last_error() and b"null" in lib.ops_last_error()
    lib.ops_clear_error()
    assert lib.ops_last_error() is None
    # empty problems are a no-op, not an error
    assert lib.ops_gemm(handle.ptr, 0, 8, 8, 1.0, None, 8, None, 8, 0.0, None, 8) == 0


def test_unaligned_shapes_take_the_simt_path(handle, lib):
    """K=1 / odd leading dimensions (the reference's AddBias trick, ops.go:335-351) are legal"""
    rng = np.random.default_rng(2)
    for (M, N, K) in [(37, 23, 1), (50, 31, 77), (64, 40, 3)]:
        A, B, C0 = rand_f16(rng, (M, K)), rand_f16(rng, (K, N)), rand_f16(rng, (M, N))
        tA, tB, tC = gpu.TensorFromFP16(A), gpu.TensorFromFP16(B), gpu.TensorFromFP16(C0)
        gpu.GEMM(handle, M, N, K, 1.0, tA, tB, 1.0, tC)
        gpu.Sync()
        want = O.gemm(A, B, 1.0, 1.0, C0)
        assert_close(tC.ToFP32(), want, gemm_tol(A, B, want) + 2.0 ** -10 * np.abs(C0), f"simt {M}x{N}x{K}")


# ------------------------------------------------------------------ CTA pairs (cta_group::2) and the shared splice tile
@pytest.mark.parametrize("cg", [1, 2])
@pytest.mark.parametrize("M,N,K,b_k_major", [(512, 160, 384, False), (1000, 256, 320, False), (384, 160, 1536, True),
                                             (2000, 1536, 192, False), (777, 64, 200, True), (260, 128, 64, False)])
def test_cta_pair_plain_gemm(handle, lib, cg, M, N, K, b_k_major):
    """256-row tiles over a CTA pair (each CTA stages half of B) == one CTA per 128-row tile == oracle"""
    rng = np.random.default_rng(M + N + K)
    A, B = rand_f16(rng, (M, K)), rand_f16(rng, (K, N), 0.1)
    want = O.gemm(A, B)
    tA, tD = gpu.TensorFromFP16(A), gpu.ZeroTensor(M, N)
    tB = gpu.TensorFromFP16(np.ascontiguousarray(B.T) if b_k_major else B)
    d = make_desc(M, N, K, tA, tB, tD, b_major=K_MAJOR if b_k_major else MN_MAJOR, force_cg=cg)
    lib.kfp16_ctx_set_max_ctas(handle.ptr, 6)          # several tiles per CTA / pair: exercises every ring
    try:
        run_desc(handle, d)
    finally:
        lib.kfp16_ctx_set_max_ctas(handle.ptr, 0)
    assert_close(tD.ToFP32(), want, gemm_tol(A, B, want), f"cg={cg} {M}x{N}x{K}")
    for t in (tA, tB, tD):
        t.Free()


@pytest.mark.parametrize("cg,no_share", [(1, 1), (2, 1), (2, 0)])
@pytest.mark.parametrize("T,D,N,s0,s1,b_k_major", [(600, 192, 160, -3, 0, False), (600, 160, 256, 0, 3, False),
                                                   (515, 256, 160, 0, -3, True), (900, 160, 1536, 3, 0, True),
                                                   (300, 64, 64, -1, 0, False)])
def test_spliced_gemm_shared_tile(handle, lib, cg, no_share, T, D, N, s0, s1, b_k_major):
    """Y = X(t+s0)*W0 + X(t+s1)*W1 (forward.go:699-790 and its transposes): separate A tiles per slab vs ONE
    A tile of 128+|s| rows read through row-shifted UMMA descriptors, on one CTA and on a CTA pair"""
    halo = 3
    rng = np.random.default_rng(T + D + N)
    X = rand_f16(rng, (T, D))
    W = rand_f16(rng, (2 * D, N), 0.08)
    Xp = np.concatenate([np.repeat(X[:1], halo, 0), X, np.repeat(X[-1:], halo, 0)], 0)
    i0, i1 = np.clip(np.arange(T) + s0, 0, T - 1), np.clip(np.arange(T) + s1, 0, T - 1)
    S = np.concatenate([X[i0], X[i1]], 1)
    want = O.gemm(S, W)
    tXp, tD = gpu.TensorFromFP16(Xp), gpu.ZeroTensor(T, N)
    # K-major B: stored [2N x D]: rows [0,N) = W0^T, rows [N,2N) = W1^T  (the layout dgrad sees: W stored [out x in])
    Bst = np.concatenate([W[:D].T, W[D:].T], 0) if b_k_major else W
    tW = gpu.TensorFromFP16(np.ascontiguousarray(Bst))
    d = make_desc(T, N, 2 * D, tXp, tW, tD, b_major=K_MAJOR if b_k_major else MN_MAJOR, force_cg=cg, no_share=no_share)
    d.A.ptr = tXp.Ptr + halo * D * 2
    d.A.rows, d.A.halo = T, halo
    d.kslabs, d.kslab_len = 2, D
    d.a_row_off[0][0], d.a_row_off[0][1] = s0, s1
    d.b_row_off[0][0], d.b_row_off[0][1] = 0, (N if b_k_major else D)
    run_desc(handle, d)
    assert_close(tD.ToFP32(), want, gemm_tol(S, W, want), f"spliced cg={cg} no_share={no_share}")
    for t in (tXp, tW, tD):
        t.Free()


@pytest.mark.parametrize("cg", [1, 2])
def test_fused_epilogues_on_cta_pairs(handle, lib, cg):
    """the specialised affine epilogue (bias, relu, bn, mask, bypass) on 128- and 256-row tiles vs the generic one"""
    M, N, K = 700, 512, 192
    rng = np.random.default_rng(17)
    A, W, X = rand_f16(rng, (M, K)), rand_f16(rng, (K, N), 0.08), rand_f16(rng, (M, N))
    bias = rand_f16(rng, (N,), 0.1)
    scale, shift = (rng.random(N).astype(np.float32) + 0.5), rng.standard_normal(N).astype(np.float32) * 0.1
    z = A.astype(np.float64) @ W.astype(np.float64) + bias
    relu = np.maximum(z, 0)
    want = relu * scale + shift + 0.66 * X
    tA, tW, tX = gpu.TensorFromFP16(A), gpu.TensorFromFP16(W), gpu.TensorFromFP16(X)
    tb = gpu.TensorFromFP16(bias.reshape(1, -1))
    sc, sh = gpu.DeviceF32(scale), gpu.DeviceF32(shift)
    mask_ld = (N + 31) // 32
    outs = []
    for generic in (0, 1):
        tD, mask = gpu.ZeroTensor(M, N), gpu.DeviceF32(n=M * mask_ld)
        d = make_desc(M, N, K, tA, tW, tD, flags=EPI_BIAS | EPI_RELU | EPI_BN | EPI_RESID | EPI_MASK, bias=tb.Ptr,
                      bn_scale=sc.Ptr, bn_shift=sh.Ptr, res_scale=0.66, mask_out=mask.Ptr, mask_ld=mask_ld, ldr=N,
                      force_cg=cg, force_generic=generic)
        d.R[0] = tX.Ptr
        run_desc(handle, d)
        got = tD.ToFP32()
        assert_close(got, want, 2.0 ** -10 * np.abs(want) * 1.01 + 2e-3, f"affine epilogue cg={cg} generic={generic}")
        bits = mask.ToHost().view(np.uint32).reshape(M, mask_ld)
        got_mask = ((bits[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(M, -1)[:, :N].astype(bool)
        assert ((got_mask == (z > 0)) | (np.abs(z) < 1e-3)).all()
        outs.append(got)
        tD.Free()
    assert np.array_equal(outs[0], outs[1])       # specialised == generic, bit for bit


@pytest.mark.parametrize("M,N,K,groups", [(1536, 160, 2000, 2), (512, 256, 900, 1), (304, 64, 700, 1)])
def test_split_k_transposed_target(handle, lib, M, N, K, groups):
    """weight gradient computed as dY^T * X (M = the long side) and accumulated TRANSPOSED into the [N x M] fp32
    bucket section, the second group reading B shifted by s rows (the TDNN-F splice): dW_g[n, m] += sum_k A[k, m] B[k+off_g, n]"""
    s = 3
    rng = np.random.default_rng(M + N)
    At = rand_f16(rng, (K, M))
    Bfull = rand_f16(rng, (K + 2 * s, N), 0.05)
    tA, tB = gpu.TensorFromFP16(At), gpu.TensorFromFP16(Bfull)
    ws = gpu.DeviceF32(n=groups * N * M)
    tD = gpu.ZeroTensor(8, 8)
    d = make_desc(M, N, K, tA, tB, tD, a_major=MN_MAJOR, b_major=MN_MAJOR, split_k=6, ws_ld=M, ws_transposed=1)
    d.groups = groups
    d.B.ptr, d.B.rows, d.B.halo = tB.Ptr + s * N * 2, K, s
    d.b_row_off[0][0], d.b_row_off[1][0] = 0, s
    d.ws[0], d.ws[1] = ws.Ptr, ws.Ptr + N * M * 4
    run_desc(handle, d)
    got = ws.ToHost().reshape(groups, N, M)
    for g, off in zip(range(groups), (0, s)):
        Bg = Bfull[s + off: s + off + K].astype(np.float64)
        want = Bg.T @ At.astype(np.float64)
        absprod = np.abs(Bg).T @ np.abs(At).astype(np.float64)
        assert_close(got[g], want, 2.0 ** -19 * absprod + 1e-6, f"transposed split-K group {g}")


# ---------------------------------------------------------------- implicit-GEMM convolution (kfp16_conv_addr)
def conv_ref_patches(x4, hout, sub, taps):
    """P[(t*hout+ho), tap*C+c] = x[t+dt, ho*sub+dh, c], zero outside (internal/nnet/forward.go:429-455) -- float64"""
    T, H, C = x4.shape
    P = np.zeros((T, hout, len(taps), C), np.float64)
    for k, (dt, dh) in enumerate(taps):
        for ho in range(hout):
            hs = ho * sub + dh
            if hs < 0 or hs >= H:
                continue
            t0, t1 = max(0, -dt), min(T, T - dt)
            P[t0:t1, ho, k, :] = x4[t0 + dt:t1 + dt, hs, :]
    return P.reshape(T * hout, len(taps) * C)


def conv_addr(d, mode, x_t, T, H, P, C, rows_h, taps_addr):
    c = d.conv
    c.mode, c.x, c.T, c.H, c.P, c.C, c.rows_h, c.ntaps = mode, x_t.Ptr, T, H, P, C, rows_h, len(taps_addr)
    for i, (dt, hq, par, brow) in enumerate(taps_addr):
        c.dt[i], c.hq[i], c.par[i], c.brow[i] = dt, hq, par, brow


TAPS9 = [(dt, dh) for dt in (-1, 0, 1) for dh in (-1, 0, 1)]


@pytest.mark.parametrize("max_ctas,no_share", [(0, 0), (6, 0), (0, 1), (0, 5)])     # 1: a box per tap; 5: shared box also for 256-wide tiles
@pytest.mark.parametrize("T,hin,sub,C,N,taps", [
    (300, 40, 1, 64, 64, TAPS9),        # 3 frames x 40 heights = 120-row tiles (the benchmark's cnn2 geometry)
    (190, 40, 2, 64, 128, TAPS9),       # height subsampling: parity planes of the input (cnn3)
    (156, 10, 1, 256, 256, TAPS9),      # 12 x 10 rows, 4 k-blocks per tap (cnn6)
    (97, 8, 1, 128, 64, [(-2, 0), (0, 1), (3, -1)]),   # 16 x 8 = full 128-row tiles, ragged T, irregular taps
    (40, 16, 2, 64, 72, [(0, 0), (1, 1)]),             # N not a tile multiple; the odd parity plane only via dh = 1
])
def test_implicit_conv_forward_gemm(handle, lib, max_ctas, no_share, T, hin, sub, C, N, taps):
    """conv.mode 1: Y[(t,h), n] = sum_taps x[t+dt, h*sub+dh, :] W[tap] with bias+ReLU+BN+mask epilogue, no patch matrix"""
    hout = hin // sub
    rng = np.random.default_rng(T * hin + C + N)
    x = rand_f16(rng, (T, hin, C))
    W = rand_f16(rng, (len(taps) * C, N), 0.05)
    bias = rand_f16(rng, (1, N), 0.1)
    scale = (rng.random(N) + 0.5).astype(np.float32)
    shift = (rng.standard_normal(N) * 0.1).astype(np.float32)
    Pm = conv_ref_patches(x.astype(np.float64), hout, sub, taps)
    z = np.maximum(Pm @ W.astype(np.float64) + bias, 0.0)
    want = z * scale + shift
    tx, tW, tb = gpu.TensorFromFP16(x.reshape(T, hin * C)), gpu.TensorFromFP16(W), gpu.TensorFromFP16(bias)
    tD = gpu.ZeroTensor(T * hout, N)
    tsc, tsh = gpu.DeviceF32(scale), gpu.DeviceF32(shift)
    mask_ld = (N + 31) // 32
    tmask = gpu.DeviceF32(n=T * hout * mask_ld)
    d = make_desc(T * hout, N, len(taps) * C, tx, tW, tD, flags=EPI_BIAS | EPI_RELU | EPI_BN | EPI_MASK, no_share=no_share)
    d.A.ptr = None
    d.bias, d.bn_scale, d.bn_shift = tb.Ptr, tsc.Ptr, tsh.Ptr
    d.mask_out, d.mask_ld = tmask.Ptr, mask_ld
    addr = []
    for k, (dt, dh) in enumerate(taps):
        par = dh % 2 if sub == 2 else 0
        addr.append((dt, (dh - par) // 2 if sub == 2 else dh, par, k * C))
    conv_addr(d, 1, tx, T, hin // sub, sub, C, hout, addr)
    lib.kfp16_ctx_set_max_ctas(handle.ptr, max_ctas)
    try:
        run_desc(handle, d)
    finally:
        lib.kfp16_ctx_set_max_ctas(handle.ptr, 0)
    tol = gemm_tol(Pm, W, want) * float(scale.max()) + 2.0 ** -10 * np.abs(want) + 1e-3
    assert_close(tD.ToFP32(), want, tol, f"implicit conv forward T={T} hin={hin} sub={sub} C={C} N={N}")
    bits = tmask.ToHost().view(np.uint32).reshape(T * hout, mask_ld)
    got_mask = ((bits[:, np.arange(N) >> 5] >> (np.arange(N) & 31)) & 1).astype(bool)
    zf = Pm @ W.astype(np.float64) + bias
    sure = np.abs(zf) > 1e-2
    assert np.array_equal(got_mask[sure], (zf > 0)[sure])
    for t in (tx, tW, tb, tD):
        t.Free()


@pytest.mark.parametrize("no_share", [0, 1, 5])
@pytest.mark.parametrize("T,hin,sub,Cin,Cout,taps", [(300, 40, 1, 64, 64, TAPS9), (190, 40, 2, 64, 128, TAPS9),
                                                     (156, 10, 1, 256, 128, TAPS9), (156, 10, 1, 256, 256, TAPS9),
                                                     (60, 16, 2, 64, 64, [(0, 0), (1, 1)])])
def test_implicit_conv_input_gradient_gemm(handle, lib, no_share, T, hin, sub, Cin, Cout, taps):
    """the adjoint: dX[t, hi, :] = sum_taps dZ[t-dt, (hi-dh)/sub, :] W[tap]^T as conv.mode 1 over dZ with mirrored taps
    (one launch per input-height parity when the layer subsamples by 2), K-major B = the forward's weight matrix"""
    hout = hin // sub
    rng = np.random.default_rng(T + hin + Cin + Cout)
    dZ = rand_f16(rng, (T, hout, Cout))
    W = rand_f16(rng, (len(taps) * Cin, Cout), 0.05)
    # oracle: dP = dZ W^T scattered back (transpose of conv_ref_patches)
    dP = (dZ.reshape(T * hout, Cout).astype(np.float64) @ W.astype(np.float64).T).reshape(T, hout, len(taps), Cin)
    want = np.zeros((T, hin, Cin), np.float64)
    absw = np.zeros_like(want)
    adP = (np.abs(dZ).reshape(T * hout, Cout).astype(np.float64) @ np.abs(W).astype(np.float64).T).reshape(T, hout, len(taps), Cin)
    for k, (dt, dh) in enumerate(taps):
        for ho in range(hout):
            hs = ho * sub + dh
            if hs < 0 or hs >= hin:
                continue
            t0, t1 = max(0, -dt), min(T, T - dt)
            want[t0 + dt:t1 + dt, hs, :] += dP[t0:t1, ho, k, :]
            absw[t0 + dt:t1 + dt, hs, :] += adP[t0:t1, ho, k, :]
    tz, tW = gpu.TensorFromFP16(dZ.reshape(T, hout * Cout)), gpu.TensorFromFP16(W)
    tD = gpu.TensorFromFP16(np.full((T, hin * Cin), 7.0, np.float32))      # every element must be overwritten
    for par in range(sub):
        addr = [(-dt, (par - dh) // sub, 0, k * Cin) for k, (dt, dh) in enumerate(taps) if (par - dh) % sub == 0]
        if not addr:
            continue
        d = make_desc(T * hout, Cin, len(addr) * Cout, tz, tW, tD, b_major=K_MAJOR, no_share=no_share)
        d.A.ptr = None
        d.D[0] = tD.Ptr + par * Cin * 2
        d.ldd = sub * Cin
        conv_addr(d, 1, tz, T, hout, 1, Cout, hout, addr)
        run_desc(handle, d)
    got = tD.ToFP32().reshape(T, hin, Cin)
    covered = np.zeros(hin, bool)
    for par in range(sub):
        if any((par - dh) % sub == 0 for _, dh in taps):
            covered[par::sub] = True
    tol = 2.0 ** -10 * np.abs(want) * 1.01 + 2.0 ** -19 * absw + 1e-6
    assert_close(got[:, covered], want[:, covered], tol[:, covered], f"implicit conv input gradient T={T} hin={hin} sub={sub}")
    for t in (tz, tW, tD):
        t.Free()


@pytest.mark.parametrize("split_k", [1, 5, 24])
@pytest.mark.parametrize("T,hin,sub,Cin,Cout,taps", [(312, 40, 1, 64, 64, TAPS9), (200, 40, 2, 64, 128, TAPS9),
                                                     (157, 10, 1, 256, 256, TAPS9), (90, 16, 1, 64, 136, [(-2, 0), (0, 1), (3, -1)])])
def test_implicit_conv_weight_gradient_gemm(handle, lib, split_k, T, hin, sub, Cin, Cout, taps):
    """conv.mode 2: dW[(tap, c), n] = sum over (t, h) of x[t+dt, h*sub+dh, c] dZ[(t,h), n]: MN-major operands, k-blocks of up to
    80 (frame, height) rows, the A chunks loaded as shifted 4-D boxes of the layer input; fp32 split-K accumulation"""
    hout = hin // sub
    rng = np.random.default_rng(T + hin + Cin + Cout + split_k)
    x = rand_f16(rng, (T, hin, Cin))
    dZ = rand_f16(rng, (T * hout, Cout), 0.05)
    Pm = conv_ref_patches(x.astype(np.float64), hout, sub, taps)
    want = Pm.T @ dZ.astype(np.float64)
    absprod = np.abs(Pm).T @ np.abs(dZ).astype(np.float64)
    tx, tz = gpu.TensorFromFP16(x.reshape(T, hin * Cin)), gpu.TensorFromFP16(dZ)
    M = len(taps) * Cin
    ws = gpu.DeviceF32(n=M * Cout)
    tD = gpu.ZeroTensor(8, 8)
    d = make_desc(M, Cout, T * hout, tx, tz, tD, a_major=MN_MAJOR, b_major=MN_MAJOR, split_k=split_k, ws_ld=Cout)
    d.A.ptr = None
    d.ws[0] = ws.Ptr
    addr = []
    for k, (dt, dh) in enumerate(taps):
        par = dh % 2 if sub == 2 else 0
        addr.append((dt, (dh - par) // 2 if sub == 2 else dh, par, k * Cin))
    conv_addr(d, 2, tx, T, hin // sub, sub, Cin, hout, addr)
    run_desc(handle, d)
    got = ws.ToHost().reshape(M, Cout)
    assert_close(got, want, 2.0 ** -19 * absprod + 1e-5, f"implicit conv weight gradient split={split_k}")
    for t in (tx, tz):
        t.Free()


@pytest.mark.parametrize("K", [2000, 9984])
@pytest.mark.parametrize("shift_a,transposed,N", [(False, True, 160), (True, False, 160), (True, False, 144)])
@pytest.mark.parametrize("merge", [True, False])
def test_spliced_weight_gradient_merged_groups(handle, lib, K, shift_a, transposed, N, merge):
    """the two groups of a spliced weight gradient, dW_g = sum_k A[k + a_off_g]^T B[k + b_off_g], share ONE A and ONE B
    tile per k-block (row-shifted UMMA descriptors on MN-major operands, one TMEM accumulator stage per group):
    shift on the B side (dWaff = dZ^T [B(t) | B(t+s)], transposed target) or on the A side (dWlin = [X(t-s) | X(t)]^T dB);
    merge=False runs the two groups as separate tiles (no_share = 1) for comparison"""
    s, M = 3, 1536
    rng = np.random.default_rng(K + N + int(shift_a))
    Afull = rand_f16(rng, (K + 2 * s, M))
    Bfull = rand_f16(rng, (K + 2 * s, N), 0.05)
    tA, tB = gpu.TensorFromFP16(Afull), gpu.TensorFromFP16(Bfull)
    ws = gpu.DeviceF32(n=2 * N * M)
    tD = gpu.ZeroTensor(8, 8)
    d = make_desc(M, N, K, tA, tB, tD, a_major=MN_MAJOR, b_major=MN_MAJOR, split_k=6, ws_ld=M if transposed else N,
                  ws_transposed=int(transposed), no_share=4 if merge else 1)
    d.groups = 2
    d.A.ptr, d.A.rows, d.A.halo = tA.Ptr + s * M * 2, K, s
    d.B.ptr, d.B.rows, d.B.halo = tB.Ptr + s * N * 2, K, s
    a_off, b_off = ((-s, 0), (0, 0)) if shift_a else ((0, 0), (0, s))
    d.a_row_off[0][0], d.a_row_off[1][0] = a_off
    d.b_row_off[0][0], d.b_row_off[1][0] = b_off
    d.ws[0], d.ws[1] = ws.Ptr, ws.Ptr + N * M * 4
    run_desc(handle, d)
    got = ws.ToHost().reshape((2, N, M) if transposed else (2, M, N))
    for g in range(2):
        Ag = Afull[s + a_off[g]: s + a_off[g] + K].astype(np.float64)
        Bg = Bfull[s + b_off[g]: s + b_off[g] + K].astype(np.float64)
        want = Ag.T @ Bg
        absprod = np.abs(Ag).T @ np.abs(Bg)
        if transposed:
            want, absprod = want.T, absprod.T
        assert_close(got[g], want, 2.0 ** -19 * absprod + 1e-6, f"merged={merge} group {g}")
    for t in (tA, tB, tD, ws):
        t.Free()


@pytest.mark.parametrize("N,count,K", [(160, 3, 2000), (64, 2, 912), (160, 5, 9984)])
def test_grouped_weight_gradients_equal_the_oracle(handle, lib, N, count, K):
    """kfp16_wgrad_group_*: `count` spliced weight-gradient problems (two row-shifted groups each, shift on the A or
    on the B side, plain or transposed target) in ONE persistent split-K launch with a device-resident problem table;
    N = 160 takes the merged-tile kernel (one A/B tile per k-block for both groups), N = 64 the two-tiles-per-problem one.
    Launched twice: the targets accumulate.  (K is a multiple of 16: a partial UMMA K step relies on TMA zero-fill past
    the matrix edge, and here the operands carry readable halo rows beyond K.)"""
    import ctypes as C
    from kaldi_fp16_b200._lib import WgradProb
    s, M = 3, 512
    rng = np.random.default_rng(N + count + K)
    probs = (WgradProb * count)()
    keep, want = [], []
    for i in range(count):
        transposed = i % 2 == 0                      # shift on the B side + transposed target, or shift on the A side
        Af = rand_f16(rng, (K + 2 * s, M))
        Bf = rand_f16(rng, (K + 2 * s, N), 0.05)
        tA, tB = gpu.TensorFromFP16(Af), gpu.TensorFromFP16(Bf)
        ws = gpu.DeviceF32(n=2 * N * M)
        keep += [tA, tB, ws]
        a_off, b_off = ((0, 0), (0, s)) if transposed else ((-s, 0), (0, 0))
        q = probs[i]
        q.A.ptr, q.A.rows, q.A.cols, q.A.ld, q.A.halo = tA.Ptr + s * M * 2, K, M, M, s
        q.B.ptr, q.B.rows, q.B.cols, q.B.ld, q.B.halo = tB.Ptr + s * N * 2, K, N, N, s
        q.a_row_off[0], q.a_row_off[1] = a_off
        q.b_row_off[0], q.b_row_off[1] = b_off
        q.ws[0], q.ws[1] = ws.Ptr, ws.Ptr + N * M * 4
        q.ws_ld, q.ws_transposed = (M if transposed else N), int(transposed)
        w = []
        for g in range(2):
            Ag = Af[s + a_off[g]: s + a_off[g] + K].astype(np.float64)
            Bg = Bf[s + b_off[g]: s + b_off[g] + K].astype(np.float64)
            d, ap = Ag.T @ Bg, np.abs(Ag).T @ np.abs(Bg)
            w.append((d.T, ap.T) if transposed else (d, ap))
        want.append((ws, transposed, w))
    grp = lib.kfp16_wgrad_group_create(handle.ptr, M, N, K, probs, count)
    assert grp, _lib_err()
    for _ in range(2):
        assert lib.kfp16_wgrad_group_launch(handle.ptr, grp) == 0, _lib_err()
    gpu.Sync()
    for i, (ws, transposed, w) in enumerate(want):
        got = ws.ToHost().reshape((2, N, M) if transposed else (2, M, N))
        for g in range(2):
            assert_close(got[g], 2.0 * w[g][0], 2.0 * (2.0 ** -19 * w[g][1] + 1e-6), f"problem {i} group {g}")
    lib.kfp16_wgrad_group_destroy(grp)
    assert not lib.kfp16_wgrad_group_create(handle.ptr, 64, N, K, probs, count)      # needs CTA-pair tiles
    for t in keep:
        t.Free()


def _lib_err():
    from kaldi_fp16_b200 import _lib
    return _lib.last_error()


# ---------------------------------------------------------------- fused dropout epilogue
@pytest.mark.parametrize("generic,resid", [(0, False), (0, True), (1, True)])
@pytest.mark.parametrize("M,N,K,p", [(900, 512, 320, 0.1), (300, 1536, 128, 0.5), (129, 72, 88, 0.25)])
def test_fused_dropout_epilogue(handle, lib, generic, resid, M, N, K, p):
    """KFP16_EPI_DROPOUT (north_star: fused dropout epilogue; semantics go/gotorch/layers.go:365-399): after
    bias + ReLU + batch-norm the element (row, col) is kept iff u(seed, row, col) > p and scaled by 1/(1-p), then the bypass is
    added; the emitted mask bit is ReLU-active AND kept.  Checked against the numpy oracle (same counter hash), for the
    specialised kinds (EK_AFFINE_DROP / EK_AFFINE_RES_DROP) and the generic run-time-flag body, with the seed taken from
    a device word as the training step does."""
    rng = np.random.default_rng(M + N + int(p * 100))
    A, W = rand_f16(rng, (M, K)), rand_f16(rng, (K, N), 0.08)
    bias = rand_f16(rng, (1, N), 0.1)
    R = rand_f16(rng, (M, N))
    scale = (rng.random(N) + 0.5).astype(np.float32)
    shift = (rng.standard_normal(N) * 0.1).astype(np.float32)
    seed_host, seed_word = 0x1234ABCD, 77
    u = O.dropout_uniform(seed_host ^ seed_word, np.arange(M), np.arange(N))
    assert abs(lib.kfp16_dropout_uniform(seed_host ^ seed_word, 5, 9) - u[5, 9]) == 0.0
    keep = u > np.float32(p)
    assert abs(keep.mean() - (1 - p)) < 0.02
    pre = A.astype(np.float64) @ W.astype(np.float64) + bias
    z = np.maximum(pre, 0.0) * scale + shift
    want = np.where(keep, z / (1.0 - p), 0.0) + (0.66 * R if resid else 0.0)
    tA, tW, tb, tR = gpu.TensorFromFP16(A), gpu.TensorFromFP16(W), gpu.TensorFromFP16(bias), gpu.TensorFromFP16(R)
    tD = gpu.ZeroTensor(M, N)
    tsc, tsh = gpu.DeviceF32(scale), gpu.DeviceF32(shift)
    mask_ld = (N + 31) // 32
    tmask = gpu.DeviceF32(n=M * mask_ld)
    tseed = gpu.DeviceF32(np.array([seed_word], np.uint32).view(np.float32))
    d = make_desc(M, N, K, tA, tW, tD, flags=EPI_BIAS | EPI_RELU | EPI_BN | EPI_MASK | EPI_DROPOUT | (EPI_RESID if resid else 0),
                  force_generic=generic, res_scale=0.66, ldr=N, drop_p=p, drop_seed=seed_host)
    d.R[0] = tR.Ptr
    d.bias, d.bn_scale, d.bn_shift = tb.Ptr, tsc.Ptr, tsh.Ptr
    d.mask_out, d.mask_ld = tmask.Ptr, mask_ld
    d.drop_seed_dev = tseed.Ptr
    kinds0 = [lib.kfp16_gemm_kind_launches(k) for k in range(11)]
    run_desc(handle, d)
    ran = [lib.kfp16_gemm_kind_launches(k) - kinds0[k] for k in range(11)]
    assert ran[0 if generic else (10 if resid else 9)] == 1, ran
    got = tD.ToFP32()
    tol = (gemm_tol(A, W, want) * float(scale.max()) + 2.0 ** -10 * np.abs(z)) / (1.0 - p) + 2.0 ** -10 * np.abs(want) + 1e-3
    assert_close(got, want, tol, f"dropout p={p} generic={generic} resid={resid}")
    if not resid:
        assert np.array_equal(got[~keep], np.zeros_like(got[~keep]))
    bits = tmask.ToHost().view(np.uint32).reshape(M, mask_ld)
    got_mask = ((bits[:, np.arange(N) >> 5] >> (np.arange(N) & 31)) & 1).astype(bool)
    sure = np.abs(pre) > 1e-2
    assert np.array_equal(got_mask[sure], ((pre > 0) & keep)[sure])
    for t in (tA, tW, tb, tR, tD):
        t.Free()
