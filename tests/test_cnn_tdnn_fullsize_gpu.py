"""The FULL-SIZE CNN-TDNN of the benchmark (BASELINE configs[2]: bench.cnn_tdnn_xconfig, 64 sequences x 150 frames,
6016 pdfs) with the epilogues bench.py times (ref_round = False, i.e. the specialised EpiKinds), layer by layer against
the REFERENCE'S OWN GPU operator library (oracle/_ref/libkaldi_fp16_ref.so, compiled unmodified from /root/reference/cpp):

  conv-relu-batchnorm (cnn1..cnn6, GEMM M = 384 000 .. 96 000): the patch matrix the reference's Go layer builds on the
      CPU (internal/nnet/forward.go:429-455; here numpy, per-sequence zero padding, Cartesian taps) -> ops_gemm ->
      AddBias (K = 1 GEMM, beta = 1: internal/gpu/ops.go:335-351) -> ops_relu -> ops_batchnorm_forward per filter;
      weight gradient: ops_batchnorm_backward -> ops_relu_backward -> ops_transpose + ops_gemm
      (AffineBackwardWeights, backward_ops.go:195-225), i.e. the transpose of that forward
  prefinal-chain (forward.go:912-968) and output (971-1001) at N = 1536 / 256 / 6016, and the prefinal BigW gradient
      (bn + relu-mask fused into the input-gradient GEMM here)

Each layer is fed THIS library's own input activation / output gradient of that layer, so every comparison isolates one
layer's kernels (EK_AFFINE on conv + prefinal affine, EK_BN, EK_BIAS, EK_BN_GRADMASK, the conv weight / input
gradient path) at the benchmark's sizes.  Tolerances: activations <= 2e-3 of the tensor's max-abs,
weight gradients <= 5e-3."""
import numpy as np
import pytest

import bench
from kaldi_fp16_b200 import gpu, nnet
from oracle import kaldi_oracle as O
from oracle.nnet_oracle import OracleNet
from tests.refbind import RefBuf, ref_half

pytestmark = pytest.mark.gpu

N_SEQ, L, PDFS = 64, 150, 6016
T = N_SEQ * L
CONV = [("cnn1", "combine_inputs"), ("cnn2", "cnn1"), ("cnn3", "cnn2"), ("cnn4", "cnn3"), ("cnn5", "cnn4"), ("cnn6", "cnn5")]


class Ref:
    def __init__(self, lib):
        self.lib = lib
        self.h = lib.ops_cublas_create()
        self.bufs = []

    def up(self, x):
        b = ref_half(self.lib, x)
        self.bufs.append(b)
        return b

    def f32(self, x):
        b = RefBuf(self.lib, np.ascontiguousarray(x, np.float32))
        self.bufs.append(b)
        return b

    def zeros(self, rows, cols):
        b = RefBuf(self.lib, np.zeros((rows, cols), np.uint16))
        self.bufs.append(b)
        return b

    def gemm(self, M, N, K, A, B, C, alpha=1.0, beta=0.0):
        assert self.lib.ops_gemm(self.h, M, N, K, alpha, A.ptr, K, B.ptr, N, beta, C.ptr, N) == 0

    def transpose(self, X, rows, cols):
        out = self.zeros(cols, rows)
        assert self.lib.ops_transpose(X.ptr, out.ptr, rows, cols) == 0
        return out

    def free_all(self):
        assert self.lib.bridge_gpu_sync() == 0
        for b in self.bufs:
            b.free()
        self.bufs = []

    def close(self):
        self.free_all()
        self.lib.ops_cublas_destroy(self.h)


def ref_affine_relu_bn(R, M, N, K, A, W, bias, bn):
    """ops_gemm -> AddBias -> ops_relu -> ops_batchnorm_forward; returns (Y, post-ReLU activation)"""
    Z = R.zeros(M, N)
    R.gemm(M, N, K, A, W, Z)
    ones = R.up(np.ones((M, 1), np.float32))
    R.gemm(M, N, 1, ones, bias, Z, beta=1.0)
    assert R.lib.ops_relu(Z.ptr, M * N) == 0
    relu = R.zeros(M, N)
    assert R.lib.ops_copy(relu.ptr, Z.ptr, M * N) == 0
    assert R.lib.ops_batchnorm_forward(Z.ptr, M, N, bn["mean"].ptr, bn["var"].ptr, bn["gamma"].ptr, bn["beta"].ptr, 1e-3) == 0
    return Z, relu


def up_bn(R, bn):
    return {k: R.f32(bn[k]) for k in ("mean", "var", "gamma", "beta")}


def test_full_size_cnn_tdnn_layers_match_the_reference_operators(handle, lib, reflib):
    rng = np.random.default_rng(31)
    xconfig = bench.cnn_tdnn_xconfig(PDFS)
    net = nnet.NewNetwork(nnet.BuildModelFromString(xconfig), handle, N_SEQ, L, train=True, ref_round=False, seed=42)
    on = OracleNet(xconfig, N_SEQ, L)          # geometry + patch builder only (its parameters are not used)
    bns = {}

    def rand_bn(layer, which, dim):
        bn = dict(mean=(rng.standard_normal(dim) * 0.1).astype(np.float32), var=(rng.random(dim) + 0.5).astype(np.float32),
                  gamma=(rng.random(dim) + 0.5).astype(np.float32), beta=(rng.standard_normal(dim) * 0.1).astype(np.float32))
        net.SetBN(layer, which, bn["mean"], bn["var"], bn["gamma"], bn["beta"], 1e-3)
        bns[(layer, which)] = bn

    for name, _ in CONV:
        fout = net.params[f"{name}.Bias"][1]
        net.SetParam(f"{name}.Bias", O.to_f16_trunc((rng.standard_normal((1, fout)) * 0.1).astype(np.float32)))
        rand_bn(name, "BN", fout)
    net.SetParam("prefinal-chain.BigBias", O.to_f16_trunc((rng.standard_normal((1, 1536)) * 0.1).astype(np.float32)))
    net.SetParam("output.Bias", O.to_f16_trunc((rng.standard_normal((1, PDFS)) * 0.1).astype(np.float32)))
    rand_bn("prefinal-chain", "PfBN", 1536)
    rand_bn("prefinal-chain", "BN", 256)

    feats = rng.standard_normal((T, 40)).astype(np.float32) * (10.0 * 0.9 ** np.arange(40, dtype=np.float32))
    feats[:, 0] = np.clip(60 + 20 * rng.standard_normal(T), -20, 105)
    ivecs = np.clip(rng.standard_normal((N_SEQ, 100)), -3, 3).astype(np.float32)
    net.MarkPerSequence("ivector", "ivector-linear", "ivector-batchnorm")
    net.SetInputF32("input", feats)            # FP32 -> FP16 (RNE) on the device
    net.SetInputF32("ivector", ivecs)
    assert lib.kfp16_net_forward(net.ptr) == 0
    gpu.Sync()
    out = net.Output("output")
    assert np.isfinite(out).all()
    # dY = Y / 64 keeps the FP16 activation gradients of this random-init net in range
    dy = O.to_f16_rne(out * np.float32(1.0 / 64.0))
    # backward twice: with one elementwise batch-norm / ReLU backward pass per conv layer (the gradient buffers then hold
    # dY, which the reference operators below start from), and -- the default, what bench.py times -- with that pass
    # folded into the consumer's input-gradient epilogue (the buffers hold dZ)
    net.SetFuseConvBackward(False)
    net.ZeroGrads()
    net.Backward(dy)
    gpu.Sync()
    wg_unfused = net.WeightGrads()
    douts = {name: net.Grad(name).astype(np.float16) for name, _ in CONV}
    net.SetFuseConvBackward(True)
    net.ZeroGrads()
    net.Backward(dy)
    gpu.Sync()
    wg = net.WeightGrads()

    R = Ref(reflib)
    # ------------------------------------------------------------------ conv layers, forward + weight gradient
    for name, src in CONV:
        l = on.by_name[name]
        x = net.Output(src)
        P, (hin, hout, sub, fin, taps) = on._patches(l, x)
        M, K = P.shape
        W = net.GetParam(f"{name}.W")
        fout = W.shape[1]
        assert K == W.shape[0] and M == T * hout
        Pr, Wr, br = R.up(P), R.up(W), R.up(net.GetParam(f"{name}.Bias"))
        del P
        bn = up_bn(R, bns[(name, "BN")])
        Z, relu = ref_affine_relu_bn(R, M, fout, K, Pr, Wr, br, bn)
        want = Z.f32().reshape(T, hout * fout)           # [(t, h) x f] == height-major [T x H*F]
        got = net.Output(name)
        err = O.max_err_vs_scale(got, want)
        assert err <= 2e-3, f"{name} forward (M={M} K={K} N={fout}) vs reference operators: err {err:.2e}"
        # weight gradient for THIS library's gradient wrt the layer output
        dout = douts[name].astype(np.float32).reshape(M, fout)
        assert np.isfinite(dout).all() and np.abs(dout).max() > 0, f"{name}: degenerate output gradient"
        dZ = R.zeros(M, fout)
        dO = R.up(dout)
        assert R.lib.ops_batchnorm_backward(dO.ptr, dZ.ptr, bn["gamma"].ptr, bn["var"].ptr, 1e-3, M, fout) == 0
        assert R.lib.ops_relu_backward(relu.ptr, dZ.ptr, M * fout) == 0
        # fused form: the layer's gradient buffer holds dZ itself (ReLU masks may differ where the pre-activation rounds
        # across zero: a small fraction of elements, each by its full value)
        dz_ref, dz_fused = dZ.f32(), net.Grad(name).reshape(M, fout)
        scale_dz = max(float(np.abs(dz_ref).max()), 1e-12)
        bad = np.abs(dz_fused - dz_ref) > 4e-3 * scale_dz
        assert bad.mean() < 2e-3, f"{name}: fused dZ differs from ops_batchnorm_backward + ops_relu_backward on {bad.mean():.2%} of the elements"
        del dz_ref, dz_fused, bad
        err = O.max_err_vs_scale(wg[f"{name}.W"], wg_unfused[f"{name}.W"])
        assert err <= 2e-3, f"{name} weight gradient, fused vs unfused backward: err {err:.2e}"
        Pt = R.transpose(Pr, M, K)
        ours = wg[f"{name}.W"]
        alpha = float(2.0 ** -np.ceil(np.log2(max(np.abs(ours).max() / 1024.0, 1.0))))   # keep the FP16 result in range
        dW = R.zeros(K, fout)
        R.gemm(K, fout, M, Pt, dZ, dW, alpha=alpha)
        want_dw = dW.f32()
        assert np.isfinite(want_dw).all()
        err = O.max_err_vs_scale(ours * np.float32(alpha), want_dw)
        assert err <= 5e-3, f"{name} weight gradient (reduction over {M} rows) vs reference operators: err {err:.2e}"
        # bias gradient: column sum of dZ (AffineBackwardBias, M = 1 GEMM)
        ones_row = R.up(np.ones((1, M), np.float32))
        db = R.zeros(1, fout)
        ob = wg[f"{name}.Bias"]
        beta_s = float(2.0 ** -np.ceil(np.log2(max(np.abs(ob).max() / 1024.0, 1.0))))
        R.gemm(1, fout, M, ones_row, dZ, db, alpha=beta_s)
        err = O.max_err_vs_scale(ob * np.float32(beta_s), db.f32())
        assert err <= 1e-2, f"{name} bias gradient: err {err:.2e}"
        R.free_all()

    # ------------------------------------------------------------------ prefinal-chain + output at full width
    x = net.Output("prefinal-l")
    Xr = R.up(x)
    BigW, SmallW = net.GetParam("prefinal-chain.BigW"), net.GetParam("prefinal-chain.SmallW")
    bn1, bn2 = up_bn(R, bns[("prefinal-chain", "PfBN")]), up_bn(R, bns[("prefinal-chain", "BN")])
    G, relu = ref_affine_relu_bn(R, T, 1536, 256, Xr, R.up(BigW), R.up(net.GetParam("prefinal-chain.BigBias")), bn1)
    Ys = R.zeros(T, 256)
    SmallWr = R.up(SmallW)
    R.gemm(T, 256, 1536, G, SmallWr, Ys)
    assert R.lib.ops_batchnorm_forward(Ys.ptr, T, 256, bn2["mean"].ptr, bn2["var"].ptr, bn2["gamma"].ptr, bn2["beta"].ptr, 1e-3) == 0
    pf = net.Output("prefinal-chain")
    err = O.max_err_vs_scale(pf, Ys.f32())
    assert err <= 2e-3, f"prefinal-chain forward vs reference operators: err {err:.2e}"
    Yo = R.zeros(T, PDFS)
    R.gemm(T, PDFS, 256, R.up(pf), R.up(net.GetParam("output.W")), Yo)
    ones = R.up(np.ones((T, 1), np.float32))
    R.gemm(T, PDFS, 1, ones, R.up(net.GetParam("output.Bias")), Yo, beta=1.0)
    err = O.max_err_vs_scale(out, Yo.f32())
    assert err <= 2e-3, f"output layer (N = {PDFS}) vs reference operators: err {err:.2e}"
    # prefinal backward: dYs = BN2 backward, dG = dYs * SmallW^T, BN1 + ReLU backward, dBigW = X^T dG
    dpf = R.up(net.Grad("prefinal-chain"))
    dYs = R.zeros(T, 256)
    assert R.lib.ops_batchnorm_backward(dpf.ptr, dYs.ptr, bn2["gamma"].ptr, bn2["var"].ptr, 1e-3, T, 256) == 0
    SWt = R.transpose(SmallWr, 1536, 256)
    dG = R.zeros(T, 1536)
    R.gemm(T, 1536, 256, dYs, SWt, dG)
    dGz = R.zeros(T, 1536)
    assert R.lib.ops_batchnorm_backward(dG.ptr, dGz.ptr, bn1["gamma"].ptr, bn1["var"].ptr, 1e-3, T, 1536) == 0
    assert R.lib.ops_relu_backward(relu.ptr, dGz.ptr, T * 1536) == 0
    Xt = R.transpose(Xr, T, 256)
    ours = wg["prefinal-chain.BigW"]
    alpha = float(2.0 ** -np.ceil(np.log2(max(np.abs(ours).max() / 1024.0, 1.0))))
    dBig = R.zeros(256, 1536)
    R.gemm(256, 1536, T, Xt, dGz, dBig, alpha=alpha)
    err = O.max_err_vs_scale(ours * np.float32(alpha), dBig.f32())
    assert err <= 5e-3, f"prefinal-chain.BigW gradient vs reference operators: err {err:.2e}"
    R.close()
    net.Free()
