"""CPU oracle for the FP16 GEMM hot path of djeday123/kaldi-fp16 -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; nothing under kaldi_fp16_b200/ does.  It is a numpy restatement of what the
reference computes on this path, every function citing the reference file:line it follows
(paths relative to /root/reference).

Numerics contract restated here: every tensor the reference stores is FP16; every matrix product
accumulates in FP32 (cublasGemmEx CUBLAS_COMPUTE_32F, cpp/cuda/ops.cu:381-392) and is rounded to
FP16 once on store; every elementwise kernel converts to float, computes in FP32 and rounds back
(cpp/cuda/ops.cu:26-320).  h(x) below = that round-to-nearest-even FP16 store.

Pinning (see DESIGN.md "Oracle"): the converters are checked against the reference's golden bit
patterns (internal/fp16/fp16_test.go:12-165, restated in tests/golden/fp16_golden.json); the GEMM
and elementwise restatements are checked on the GPU box against the reference's own compiled
library (oracle/_ref/libkaldi_fp16_ref.so, built by oracle/Makefile from the unmodified sources),
and against fixtures produced by that library (tests/golden/ref_ops_*.npz, generator:
scripts/gen_ref_golden.py).
"""
from __future__ import annotations

import numpy as np

f32 = np.float32
f16 = np.float16


def h(x: np.ndarray) -> np.ndarray:
    """FP16 store of an FP32 value (round to nearest even), returned as float32."""
    with np.errstate(over="ignore", invalid="ignore"):
        return np.asarray(x, dtype=f32).astype(f16).astype(f32)


# ----------------------------------------------------------------------------- converters
def float32_to_fp16_bits_trunc(x: np.ndarray) -> np.ndarray:
    """internal/gpu/tensor.go:158-174 (float32ToFP16Bits): TRUNCATING conversion used for weights
    and all test inputs.  exp>15 (incl. Inf/NaN) -> +-Inf, exp<-14 -> +-0, else mantissa>>13."""
    bits = np.ascontiguousarray(x, dtype=f32).view(np.uint32)
    out = np.empty(bits.shape, dtype=np.uint16)
    flat_in, flat_out = bits.reshape(-1), out.reshape(-1)
    sign = ((flat_in >> 16) & 0x8000).astype(np.uint32)
    exp = ((flat_in >> 23) & 0xFF).astype(np.int64) - 127
    frac = flat_in & 0x7FFFFF
    normal = sign | (((exp + 15) & 0x1F).astype(np.uint32) << 10) | (frac >> 13)
    res = np.where(exp > 15, sign | 0x7C00, np.where(exp < -14, sign, normal))
    flat_out[:] = res.astype(np.uint16)
    return out


def fp16_bits_to_float32(bits: np.ndarray) -> np.ndarray:
    """internal/gpu/tensor.go:176-203 / internal/fp16/fp16.go:72-104: exact widening."""
    return np.ascontiguousarray(bits, dtype=np.uint16).view(f16).astype(f32)


def fp16_from_float32_rne(x: np.ndarray) -> np.ndarray:
    """internal/fp16/fp16.go:13-70 (FromFloat32), restated branch by branch: round-to-nearest-even
    incl. subnormals; NaN keeps the top mantissa bits; used for FEATURES
    (internal/gpu/bridge.go:141).  Returns uint16 bit patterns."""
    b = np.ascontiguousarray(x, dtype=f32).view(np.uint32).reshape(-1).astype(np.uint64)
    sign = (b >> 31) & 1
    exp = ((b >> 23) & 0xFF).astype(np.int64)
    frac = b & 0x7FFFFF
    out = np.zeros(b.shape, dtype=np.uint64)

    # exp == 255: Inf / NaN
    m = exp == 255
    out[m] = (sign[m] << 15) | 0x7C00 | np.where(frac[m] == 0, 0, frac[m] >> 13)
    # exp > 142: overflow -> Inf
    m_over = (exp > 142) & (exp != 255)
    out[m_over] = (sign[m_over] << 15) | 0x7C00
    # 112 < exp <= 142: normal
    m = (exp > 112) & (exp <= 142)
    if m.any():
        ne = (exp[m] - 112).astype(np.uint64)
        fr = frac[m]
        rnd = fr & 0x1FFF
        fr = fr >> 13
        up = (rnd > 0x1000) | ((rnd == 0x1000) & ((fr & 1) != 0))
        fr = fr + up.astype(np.uint64)
        carry = fr > 0x3FF
        fr = np.where(carry, 0, fr)
        ne = ne + carry.astype(np.uint64)
        val = (sign[m] << 15) | (ne << 10) | fr
        val = np.where(ne > 30, (sign[m] << 15) | 0x7C00, val)
        out[m] = val
    # 101 < exp <= 112: subnormal half
    m = (exp > 101) & (exp <= 112)
    if m.any():
        shift = (113 - exp[m]).astype(np.uint64)
        fr = frac[m] | 0x800000
        rnd = fr & ((np.uint64(1) << (shift + 13)) - 1)
        half = np.uint64(1) << (shift + 12)
        fr = fr >> (shift + 13)
        up = (rnd > half) | ((rnd == half) & ((fr & 1) != 0))
        fr = fr + up.astype(np.uint64)
        out[m] = (sign[m] << 15) | fr
    # exp <= 101: +-0
    m = exp <= 101
    out[m] = sign[m] << 15
    return out.astype(np.uint16).reshape(np.shape(x))


def to_f16_trunc(x: np.ndarray) -> np.ndarray:
    """float32 -> fp16-representable float32 through the truncating converter."""
    return fp16_bits_to_float32(float32_to_fp16_bits_trunc(x))


def to_f16_rne(x: np.ndarray) -> np.ndarray:
    return fp16_bits_to_float32(fp16_from_float32_rne(x))


# ----------------------------------------------------------------------------- GEMM
def gemm(A: np.ndarray, B: np.ndarray, alpha: float = 1.0, beta: float = 0.0, C: np.ndarray | None = None,
         transA: bool = False, transB: bool = False) -> np.ndarray:
    """ops_gemm (cpp/cuda/ops.cu:366-400): C = alpha*A*B + beta*C, FP16 operands (passed here as
    fp16-representable float32), FP32 accumulate, alpha/beta float, one FP16 rounding on store.
    transA/transB follow kaldi_gemm (cpp/src/cgo_interface.cu:206-243)."""
    a = np.asarray(A, dtype=f32)
    b = np.asarray(B, dtype=f32)
    if transA:
        a = a.T
    if transB:
        b = b.T
    acc = a @ b  # fp32 accumulate (summation order differs from cuBLAS: covered by the tolerance)
    out = f32(alpha) * acc
    if beta != 0.0:
        # measured on the reference library itself (tests/golden/ref_ops.npz, gemm2): with beta != 0
        # cublasGemmEx returns h(h(alpha*acc) + beta*C) -- the product is rounded to FP16 before C is
        # added (99.95% of elements bit-identical to this form, 71% to the single-rounding form).
        out = h(out) + f32(beta) * np.asarray(C, dtype=f32)
    return h(out)


def gemm_f64(A: np.ndarray, B: np.ndarray) -> np.ndarray:
    """float64 product of fp16-representable inputs: the 'exact' value used to bound both the
    oracle's and the kernel's FP32 accumulation error (cmd/backtest/main.go:148-174 does the same
    with a float64 CPU loop)."""
    return np.asarray(A, dtype=np.float64) @ np.asarray(B, dtype=np.float64)


# ----------------------------------------------------------------------------- elementwise forward
def relu(x):
    """kernel_relu (ops.cu:26-37): only x<0 is replaced, NaN and -0 pass through."""
    x = np.asarray(x, dtype=f32)
    return np.where(x < 0, f32(0), x)


def sigmoid(x):
    """kernel_sigmoid (ops.cu:39-47)"""
    x = np.asarray(x, dtype=f32)
    with np.errstate(over="ignore"):
        return h(f32(1) / (f32(1) + np.exp(-x)))


def tanh_act(x):
    """kernel_tanh (ops.cu:49-57)"""
    return h(np.tanh(np.asarray(x, dtype=f32)))


def clipped_relu(x, ceiling):
    """kernel_clipped_relu (ops.cu:59-68): fmaxf(0, fminf(x, ceiling)) -- fminf/fmaxf drop NaN, so
    NaN -> ceiling (pinned by tests/golden/ref_ops.npz)"""
    return h(np.fmax(f32(0), np.fmin(np.asarray(x, dtype=f32), f32(ceiling))))


def softmax(x):
    """kernel_softmax (ops.cu:70-116) with the TRUE row maximum (the reference's atomicMax on the
    float bit pattern is wrong when the row maximum is negative -- documented deviation); keeps
    the reference's intermediate FP16 store of exp(x-max) before normalising (96-110)."""
    x = np.asarray(x, dtype=f32)
    m = x.max(axis=1, keepdims=True)
    e = np.exp(x - m).astype(f32)
    s = e.sum(axis=1, keepdims=True, dtype=f32)
    return h(h(e) * (f32(1) / s))


def log_softmax(x):
    """kernel_log_softmax (ops.cu:118-166), true maximum as above."""
    x = np.asarray(x, dtype=f32)
    m = x.max(axis=1, keepdims=True)
    s = np.exp(x - m).astype(f32).sum(axis=1, keepdims=True, dtype=f32)
    return h(x - (m + np.log(s)))


def batchnorm_forward(x, mean, var, gamma, beta, eps):
    """kernel_batchnorm_forward (ops.cu:171-187): gamma*(x-mean)/sqrt(var+eps)+beta in FP32."""
    x = np.asarray(x, dtype=f32)
    norm = (x - f32(mean)) / np.sqrt(np.asarray(var, f32) + f32(eps))
    return h(np.asarray(gamma, f32) * norm + np.asarray(beta, f32))


def batchnorm_forward_rms(x, mean, var, target_rms, eps):
    """kernel_batchnorm_rms (ops.cu:191-204)"""
    x = np.asarray(x, dtype=f32)
    return h((x - np.asarray(mean, f32)) / np.sqrt(np.asarray(var, f32) + f32(eps)) * f32(target_rms))


def add_scaled(dst, src, alpha, beta):
    """kernel_add_scaled (ops.cu:207-217): dst = alpha*src + beta*dst"""
    return h(f32(alpha) * np.asarray(src, f32) + f32(beta) * np.asarray(dst, f32))


def add(dst, src):
    """kernel_add (ops.cu:219-228)"""
    return h(np.asarray(dst, f32) + np.asarray(src, f32))


def add_bias(x, bias):
    """gpu.AddBias (internal/gpu/ops.go:335-351): ones[Tx1]*bias[1xD] GEMM with beta=1 -> h(x+b)"""
    return h(np.asarray(x, f32) + np.asarray(bias, f32).reshape(1, -1))


def concat_cols(dst, src, off):
    """kernel_concat_cols (ops.cu:241-254)"""
    out = np.array(dst, dtype=f32, copy=True)
    out[:, off:off + src.shape[1]] = src
    return out


def slice_cols(src, off, cols):
    """kernel_slice_cols (ops.cu:308-320)"""
    return np.array(src[:, off:off + cols], dtype=f32, copy=True)


def combine_feature_maps(x, height, nf1, nf2):
    """kernel_combine_feature_maps (ops.cu:258-287): [T x (H*F1 | H*F2)] -> [T x H*(F1+F2)]"""
    x = np.asarray(x, dtype=f32)
    T = x.shape[0]
    a = x[:, :height * nf1].reshape(T, height, nf1)
    b = x[:, height * nf1:].reshape(T, height, nf2)
    return np.concatenate([a, b], axis=2).reshape(T, height * (nf1 + nf2))


def subsample_rows(x, stride, row_offset):
    """kernel_subsample_rows (ops.cu:290-304, 628-640)"""
    return np.array(x[row_offset::stride], dtype=f32, copy=True)


# ----------------------------------------------------------------------------- backward ops
# ----------------------------------------------------------------------------- egs feature decode
def _u16_to_float(gmin, grange, v):
    """internal/parser/matrix.go:11-14 uint16ToFloat, float32 arithmetic left to right"""
    return f32(gmin) + (f32(grange) * f32(1.52590218966964e-05)) * v.astype(f32)


def decode_cm(payload: bytes, rows: int, cols: int, gmin: float, grange: float) -> np.ndarray:
    """ReadCompressedMatrix (matrix.go:28-84): cols x 4 uint16 percentiles, then bytes COLUMN-major; charToFloat (17-26)"""
    hdr = np.frombuffer(payload, np.uint16, cols * 4).reshape(cols, 4)
    data = np.frombuffer(payload, np.uint8, rows * cols, offset=cols * 8).reshape(cols, rows).T      # [rows x cols]
    p = _u16_to_float(gmin, grange, hdr)                                                            # [cols x 4] float32
    p0, p25, p75, p100 = p[:, 0], p[:, 1], p[:, 2], p[:, 3]
    v = data.astype(f32)
    b1 = p0 + ((p25 - p0) * v) * f32(1.0 / 64.0)
    b2 = p25 + ((p75 - p25) * (v - f32(64))) * f32(1.0 / 128.0)
    b3 = (p75.astype(np.float64) + ((p100 - p75) * (v - f32(192))).astype(np.float64) / 63.0).astype(f32)
    return np.where(data <= 64, b1, np.where(data <= 192, b2, b3)).astype(f32)


def decode_cm2(payload: bytes, rows: int, cols: int, gmin: float, grange: float) -> np.ndarray:
    """ReadCompressedMatrix2 (matrix.go:86-113): uint16 row-major, min + value * (range / 65535)"""
    v = np.frombuffer(payload, np.uint16, rows * cols).reshape(rows, cols).astype(f32)
    return (f32(gmin) + v * (f32(grange) / f32(65535.0))).astype(f32)


def decode_cm3(payload: bytes, rows: int, cols: int, gmin: float, grange: float) -> np.ndarray:
    """ReadCompressedMatrix3 (matrix.go:115-142): uint8 row-major, min + value * (range / 255)"""
    v = np.frombuffer(payload, np.uint8, rows * cols).reshape(rows, cols).astype(f32)
    return (f32(gmin) + v * (f32(grange) / f32(255.0))).astype(f32)


def decode_fm(payload: bytes, rows: int, cols: int) -> np.ndarray:
    """ReadFullMatrix (matrix.go:144-165): float32 little-endian row-major"""
    return np.frombuffer(payload, "<f4", rows * cols).reshape(rows, cols).astype(f32)


def dropout_uniform(seed: int, rows: np.ndarray, cols: np.ndarray) -> np.ndarray:
    """counter-based uniform in [0,1) per element (row, col): the integer hash the fused dropout epilogue uses
    (kaldi_fp16_b200/csrc/gemm_sm100.cuh::dropout_uniform), restated with uint32 wrap-around arithmetic.  The reference
    draws rand.Float64() per element (go/gotorch/layers.go:378); a counter-based generator gives the same distribution
    and makes the mask a pure function of (seed, row, col)."""
    with np.errstate(over="ignore"):
        r = np.asarray(rows, np.uint32)[:, None]
        c = np.asarray(cols, np.uint32)[None, :]
        x = np.uint32(seed & 0xFFFFFFFF) ^ (r * np.uint32(0x9E3779B1)) ^ (c * np.uint32(0x85EBCA77))
        x = x ^ (x >> np.uint32(16))
        x = x * np.uint32(0x7FEB352D)
        x = x ^ (x >> np.uint32(15))
        x = x * np.uint32(0x846CA68B)
        x = x ^ (x >> np.uint32(16))
    return (x >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)


def dropout_forward(x: np.ndarray, keep: np.ndarray, p: float) -> np.ndarray:
    """DropoutLayer.Forward (go/gotorch/layers.go:365-383): inverted dropout, kept values scaled by 1/(1-p); FP16 store"""
    return h(np.where(keep, x.astype(np.float32) * np.float32(1.0 / (1.0 - p)), np.float32(0)))


def dropout_backward(grad: np.ndarray, keep: np.ndarray, p: float) -> np.ndarray:
    """DropoutLayer.Backward (go/gotorch/layers.go:385-399)"""
    return h(np.where(keep, grad.astype(np.float32) * np.float32(1.0 / (1.0 - p)), np.float32(0)))


def relu_backward(x, grad):
    """bw_relu_backward_kernel (backward_wrappers.cu:41-49): grad if x>0 else 0"""
    return np.where(np.asarray(x, f32) > 0, np.asarray(grad, f32), f32(0))


def sigmoid_backward(out, grad):
    """bw_sigmoid_backward_kernel (backward_wrappers.cu:51-61): half arithmetic, each op rounded"""
    o, g = np.asarray(out, f32), np.asarray(grad, f32)
    return h(h(g * o) * h(f32(1) - o))


def tanh_backward(out, grad):
    """bw_tanh_backward_kernel (backward_wrappers.cu:63-73): __hmul(g, __hsub(1, __hmul(o,o))); the
    compiled reference contracts 1 - o*o into one HFMA (single rounding) -- pinned bit-for-bit by
    tests/golden/ref_ops.npz"""
    o, g = np.asarray(out, f32), np.asarray(grad, f32)
    return h(g * h(f32(1) - o * o))


def transpose(x):
    """bw_transpose_kernel (backward_wrappers.cu:75-85)"""
    return np.ascontiguousarray(np.asarray(x, f32).T)


def batchnorm_backward(grad_out, gamma, var, eps):
    """bw_batchnorm_backward_kernel (backward_wrappers.cu:104-115)"""
    scale = np.asarray(gamma, f32) / np.sqrt(np.asarray(var, f32) + f32(eps))
    return h(np.asarray(grad_out, f32) * scale)


def affine_backward_data(grad_out, W):
    """gpu.AffineBackwardData (internal/gpu/backward_ops.go:162-192): dX = dY * W^T"""
    return gemm(grad_out, W, transB=True)


def affine_backward_weights(x, grad_out):
    """gpu.AffineBackwardWeights (backward_ops.go:195-225): dW = X^T * dY"""
    return gemm(x, grad_out, transA=True)


def affine_backward_bias(grad_out):
    """gpu.AffineBackwardBias (backward_ops.go:228-253): ones[1xT] * dY, FP32 accumulate"""
    return h(np.asarray(grad_out, f32).sum(axis=0, dtype=f32).reshape(1, -1))


def sgd_update(w32, grad16, vel, lr, momentum):
    """bw_sgd_update_kernel (backward_wrappers.cu:129-142):
    g=float(g16); v=m*v+g; w32-=lr*v; w16=half(w32).  Returns (w32, w16, v)."""
    g = np.asarray(grad16, f32)
    v = (f32(momentum) * np.asarray(vel, f32) + g).astype(f32)
    w = (np.asarray(w32, f32) - f32(lr) * v).astype(f32)
    return w, h(w), v


# ----------------------------------------------------------------------------- error metric
def max_rel_err(got, want, floor=1e-6):
    """cmd/backtest/main.go:452-462: max |got-want| / max(|want|, floor)"""
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    return float(np.max(np.abs(got - want) / np.maximum(np.abs(want), floor))) if got.size else 0.0


def max_err_vs_scale(got, want):
    """max |got-want| relative to the tensor's max-abs (the gradient tolerance of SURVEY 8c)."""
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    scale = max(float(np.max(np.abs(want))) if want.size else 0.0, 1e-12)
    return float(np.max(np.abs(got - want))) / scale if got.size else 0.0
