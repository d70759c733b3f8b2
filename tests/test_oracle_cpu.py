"""CPU tests (no GPU): pin the numpy oracle against
  * the reference's own golden bit patterns (internal/fp16/fp16_test.go, restated in tests/golden/fp16_golden.json)
  * outputs of the REFERENCE's compiled operator library (cuBLAS-backed, built unmodified from
    /root/reference/cpp by oracle/Makefile) captured on a B200 by scripts/gen_ref_golden.py
    -> tests/golden/ref_ops.npz.
"""
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import kaldi_oracle as O

GOLD = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def ref():
    return np.load(GOLD / "ref_ops.npz")


def f(bits):
    return np.asarray(bits, np.uint16).view(np.float16).astype(np.float32)


def ulp16(x):
    """one fp16 ulp at |x| (normal range), as float64"""
    ax = np.maximum(np.abs(np.asarray(x, np.float64)), 2.0 ** -14)
    return 2.0 ** (np.floor(np.log2(ax)) - 10)


# ------------------------------------------------------------------ converters (fp16_test.go)
def test_rne_converter_golden_bits():
    g = json.loads((GOLD / "fp16_golden.json").read_text())
    for val, want in g["rne_exact"]:
        got = int(O.fp16_from_float32_rne(np.array([float(val)], np.float32))[0])
        assert got == int(want, 16), f"{val} -> {got:04x}, expected {want}"
    nan = int(O.fp16_from_float32_rne(np.array([np.nan], np.float32))[0])
    assert (nan >> 10) & 0x1F == 31 and nan & 0x3FF != 0                         # TestFromFloat32_NaN
    sn = np.array([g["smallest_normal"]], np.float32)
    assert O.fp16_bits_to_float32(O.fp16_from_float32_rne(sn))[0] == sn[0]       # TestFromFloat32_SmallestNormal
    sd = np.array([g["smallest_denorm"]], np.float32)
    assert O.fp16_from_float32_rne(sd)[0] != 0                                   # TestFromFloat32_Denormalized
    v = np.array(g["speech_features"], np.float32)
    back = O.fp16_bits_to_float32(O.fp16_from_float32_rne(v))
    assert np.max(np.abs(v - back) / np.abs(v)) < g["speech_features_max_rel_err"]


def test_rne_converter_equals_ieee_everywhere():
    """the branch-by-branch restatement of fp16.go:13-70 == IEEE RNE (numpy) on every finite half-range value"""
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.standard_normal(200000).astype(np.float32) * s for s in (1e-8, 1e-5, 1e-3, 1.0, 100.0, 3e4)])
    with np.errstate(over="ignore"):
        want = x.astype(np.float16).view(np.uint16)
    assert np.array_equal(O.fp16_from_float32_rne(x), want)
    allh = np.arange(65536, dtype=np.uint16)
    fin = ~np.isnan(allh.view(np.float16))
    assert np.array_equal(O.fp16_from_float32_rne(O.fp16_bits_to_float32(allh[fin])), allh[fin])   # round trip


def test_truncating_converter():
    """internal/gpu/tensor.go:158-174: mantissa >> 13, flush subnormals, exp > 15 -> Inf"""
    x = np.array([0.0, -0.0, 1.0, 1.0009765625, 1.0 + 2 ** -11 + 2 ** -12, 65504.0, 65520.0, 70000.0, 6.0e-5, 2.0 ** -14,
                  -3.14159, np.inf], np.float32)
    got = O.float32_to_fp16_bits_trunc(x)
    want = [0x0000, 0x8000, 0x3C00, 0x3C01, 0x3C00, 0x7BFF, 0x7BFF, 0x7C00, 0x0000, 0x0400, 0xC248, 0x7C00]
    assert [int(v) for v in got] == want
    rng = np.random.default_rng(1)
    y = rng.standard_normal(100000).astype(np.float32) * 10
    t = O.to_f16_trunc(y)
    big = np.abs(y) >= 2.0 ** -14
    assert np.all(np.abs(t) <= np.abs(y)) and np.all(np.abs(y - t)[big] <= ulp16(y)[big])   # toward zero, < 1 ulp
    assert np.all(t[~big] == 0)                                                             # subnormals flushed


# ------------------------------------------------------------------ ops_gemm vs the reference (cublasGemmEx)
@pytest.mark.parametrize("i", range(6))
def test_gemm_oracle_matches_reference_cublas(ref, i):
    A, B, C0 = f(ref[f"gemm{i}_A"]), f(ref[f"gemm{i}_B"]), f(ref[f"gemm{i}_C0"])
    alpha, beta = [float(v) for v in ref[f"gemm{i}_ab"]]
    want = f(ref[f"gemm{i}_out"])
    got = O.gemm(A, B, alpha, beta, C0)
    # same arithmetic (fp32 accumulate, one fp16 rounding); only the summation order differs
    exact = alpha * O.gemm_f64(A, B) + beta * C0.astype(np.float64)
    # (with beta != 0 cuBLAS rounds alpha*acc to fp16 before adding beta*C: one more half-ulp, see oracle.gemm)
    tol = ulp16(exact) * 0.51 + 2.0 ** -20 * (np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64))
    if beta != 0:
        tol = tol + 0.51 * ulp16(alpha * O.gemm_f64(A, B))
    assert np.all(np.abs(want - exact) <= tol), "reference itself outside the fp32-accumulate bound?"
    assert np.all(np.abs(got - exact) <= tol)
    assert np.mean(got == want) > 0.99           # and bit-identical on almost every element
    extra = 1.01 * ulp16(alpha * O.gemm_f64(A, B)) if beta != 0 else 0.0
    assert np.all(np.abs(got - want) <= ulp16(want) * 1.01 + extra + 2.0 ** -19 * (np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64)))


def test_kaldi_gemm_surface(ref):
    """kaldi_gemm = cublasHgemm (cgo_interface.cu:206-243): fp16 accumulate is allowed to differ more"""
    A, B, want = ref["kgemm_A"], ref["kgemm_B"], ref["kgemm_out"]
    got = O.gemm(A, B)
    assert O.max_err_vs_scale(got, want) < 4e-3


# ------------------------------------------------------------------ elementwise kernels, bit-exact where the math is exact
def same_bits(got_f32, want_bits, allow_nan=True):
    g = np.asarray(got_f32, np.float32).astype(np.float16).view(np.uint16)
    w = np.asarray(want_bits, np.uint16)
    eq = g == w
    if allow_nan:
        eq |= np.isnan(g.view(np.float16)) & np.isnan(w.view(np.float16))
    return eq


def test_exact_ops_bit_identical(ref):
    x, y, g = f(ref["x"]), f(ref["y"]), f(ref["g"])
    assert same_bits(O.relu(x), ref["relu"]).all()                              # incl. -0 and NaN pass-through
    assert same_bits(O.add(x, y), ref["add"]).all()
    assert same_bits(O.add_scaled(x, y, 0.66, 1.0), ref["add_scaled"]).mean() > 0.999   # fma contraction may flip a tie
    assert same_bits(O.transpose(x), ref["transpose"]).all()
    assert same_bits(O.subsample_rows(x, 3, 1), ref["subsample_3_1"]).all()
    assert same_bits(O.combine_feature_maps(f(ref["combine_x"]), 8, 1, 5), ref["combine"]).all()
    assert same_bits(O.relu_backward(f(ref["relu_backward_act"]), g), ref["relu_backward"]).all()
    assert same_bits(O.clipped_relu(x, 1.5), ref["clipped_relu"], allow_nan=True).mean() > 0.999


def test_transcendental_ops_within_one_ulp(ref):
    x, xs, g = f(ref["x"]), f(ref["xs"]), f(ref["g"])
    ok = ~np.isnan(x)
    for name, fn in [("sigmoid", O.sigmoid), ("tanh", O.tanh_act)]:
        got, want = fn(x), f(ref[name])
        assert np.all(np.abs(got - want)[ok] <= 2 * ulp16(want)[ok] + 1e-4), name    # --use_fast_math in the reference build
    for name, fn in [("softmax", O.softmax), ("log_softmax", O.log_softmax)]:
        got, want = fn(xs), f(ref[name])
        assert np.max(np.abs(got - want)) <= 4e-3, name
    for name, fn in [("sigmoid_backward", O.sigmoid_backward), ("tanh_backward", O.tanh_backward)]:
        got, want = fn(f(ref[name + "_act"]), g), f(ref[name])
        okk = ~np.isnan(want)
        assert np.array_equal(got[okk], want[okk]), name          # half arithmetic restated op by op: bit-identical


def test_batchnorm_ops(ref):
    x, g = f(ref["x"]), f(ref["g"])
    m, v, ga, be = ref["bn_mean"], ref["bn_var"], ref["bn_gamma"], ref["bn_beta"]
    ok = ~np.isnan(x)
    for got, want in [(O.batchnorm_forward(x, m, v, ga, be, 1e-3), f(ref["bn_fwd"])),
                      (O.batchnorm_forward_rms(x, m, v, 0.025, 1e-3), f(ref["bn_rms"])),
                      (O.batchnorm_backward(g, ga, v, 1e-5), f(ref["bn_bwd"]))]:
        o = ok & ~np.isnan(want)
        assert np.all(np.abs(got - want)[o] <= 1.01 * ulp16(want)[o])     # rsqrtf vs 1/sqrt: at most 1 ulp


def test_sgd_update_with_fp32_masters(ref):
    """ops_sgd_update (backward_wrappers.cu:129-142), 3 steps lr 0.01 momentum 0.9 (cmd/sgdtest tests 1-2)"""
    w32 = f(ref["sgd_w0"])
    vel = np.zeros_like(w32)
    for gbits in ref["sgd_grads"]:
        w32, w16, vel = O.sgd_update(w32, f(gbits), vel, 0.01, 0.9)
    assert np.max(np.abs(vel - ref["sgd_vel"])) <= 1e-6
    assert np.max(np.abs(w32 - ref["sgd_w32"])) <= 1e-6
    assert (w16.astype(np.float16).view(np.uint16) == ref["sgd_w16"]).mean() > 0.995


def test_sgd_fp32_master_precision():
    """cmd/sgdtest/main.go:142-193: 100 steps lr=1e-4 with grad 1.0 from w=1.0 -> 0.99 +- 0.002; an FP16-only
    update would never move (1e-4 < half ulp at 1.0)"""
    w32, vel = np.array([1.0], np.float32), np.zeros(1, np.float32)
    for _ in range(100):
        w32, w16, vel = O.sgd_update(w32, np.array([1.0], np.float32), vel, 1e-4, 0.0)
    assert abs(float(w32[0]) - 0.99) < 0.002 and abs(float(w16[0]) - 0.99) < 0.002


# ------------------------------------------------------------------ backward formulas (cmd/backtest restated)
def test_affine_backward_vs_float64():
    """cmd/backtest/main.go:148-174: AffineBackwardData vs float64 loop on FP16-rounded inputs, T=16 M=32 K=24"""
    rng = np.random.default_rng(5)
    T, M, K = 16, 32, 24
    go = O.to_f16_trunc(rng.uniform(-1, 1, (T, K)).astype(np.float32))
    W = O.to_f16_trunc(rng.uniform(-1, 1, (M, K)).astype(np.float32))
    got = O.affine_backward_data(go, W)
    want = go.astype(np.float64) @ W.astype(np.float64).T
    assert O.max_rel_err(got, want, floor=1e-2) < 2e-3
    X = O.to_f16_trunc(rng.uniform(-1, 1, (T, M)).astype(np.float32))
    assert O.max_err_vs_scale(O.affine_backward_weights(X, go), X.astype(np.float64).T @ go.astype(np.float64)) < 2e-3
    assert O.max_err_vs_scale(O.affine_backward_bias(go), go.astype(np.float64).sum(0, keepdims=True)) < 2e-3


def test_dropout_hash_matches_the_library_host_function():
    """oracle dropout_uniform == the integer hash compiled into the library (kfp16_dropout_uniform is a host function:
    no GPU needed); inverted-dropout forward / backward follow go/gotorch/layers.go:365-399"""
    from kaldi_fp16_b200 import _lib
    lib = _lib.load()
    rows, cols = np.arange(0, 4000, 37), np.arange(0, 1536, 11)
    for seed in (0, 0xC0FFEE, 0xFFFFFFFF):
        u = O.dropout_uniform(seed, rows, cols)
        assert u.min() >= 0.0 and u.max() < 1.0
        for i in (0, 5, len(rows) - 1):
            for j in (0, 3, len(cols) - 1):
                assert lib.kfp16_dropout_uniform(seed, int(rows[i]), int(cols[j])) == u[i, j]
    u = O.dropout_uniform(123, np.arange(512), np.arange(512))
    assert abs(u.mean() - 0.5) < 5e-3 and abs((u > 0.2).mean() - 0.8) < 5e-3
    x = np.arange(12, dtype=np.float32).reshape(3, 4)
    keep = np.array([[1, 0, 1, 1], [0, 0, 1, 1], [1, 1, 1, 0]], bool)
    y = O.dropout_forward(x, keep, 0.5)
    assert np.array_equal(y, np.where(keep, 2 * x, 0))
    assert np.array_equal(O.dropout_backward(np.ones_like(x), keep, 0.5), np.where(keep, 2.0, 0.0).astype(np.float32))


def test_compressed_matrix_decode_matches_the_scalar_restatement():
    """oracle decode_cm / cm2 / cm3 (vectorised) == internal/parser/matrix.go's loops restated element by element"""
    rng = np.random.default_rng(2)
    rows, cols, gmin, grange = 7, 5, np.float32(-13.25), np.float32(97.5)
    hdr = np.sort(rng.integers(0, 65536, size=(cols, 4)).astype(np.uint16), axis=1)
    data = rng.integers(0, 256, size=(cols, rows)).astype(np.uint8)
    data[0, :4] = [0, 64, 192, 255]
    data[1, :4] = [65, 193, 1, 128]
    got = O.decode_cm(hdr.tobytes() + data.tobytes(), rows, cols, gmin, grange)

    def u16(v):
        return np.float32(gmin) + (np.float32(grange) * np.float32(1.52590218966964e-05)) * np.float32(v)

    for c in range(cols):
        p0, p25, p75, p100 = (u16(hdr[c, k]) for k in range(4))
        for r in range(rows):
            v = int(data[c, r])
            if v <= 64:
                want = p0 + (p25 - p0) * np.float32(v) * np.float32(1.0 / 64.0)
            elif v <= 192:
                want = p25 + (p75 - p25) * np.float32(v - 64) * np.float32(1.0 / 128.0)
            else:
                want = np.float32(float(p75) + float((p100 - p75) * np.float32(v - 192)) / 63.0)
            assert got[r, c] == want, (r, c, v)
    raw16 = rng.integers(0, 65536, size=(rows, cols)).astype(np.uint16)
    inc = np.float32(grange) / np.float32(65535.0)
    assert np.array_equal(O.decode_cm2(raw16.tobytes(), rows, cols, gmin, grange), np.float32(gmin) + raw16.astype(np.float32) * inc)
    raw8 = rng.integers(0, 256, size=(rows, cols)).astype(np.uint8)
    inc = np.float32(grange) / np.float32(255.0)
    assert np.array_equal(O.decode_cm3(raw8.tobytes(), rows, cols, gmin, grange), np.float32(gmin) + raw8.astype(np.float32) * inc)
    x = rng.standard_normal((rows, cols)).astype("<f4")
    assert np.array_equal(O.decode_fm(x.tobytes(), rows, cols), x)
