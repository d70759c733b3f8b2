"""Host-side mirror of the reference's Go package ``internal/gpu`` over the C ABI (ctypes).

The reference binds cpp/include/{ops,bridge}.h through cgo (internal/gpu/ops.go:3-10); there is
no Go toolchain in this image, so the same operator surface -- same names, argument meaning and
error behaviour -- is mirrored here in Python over the SAME shared library a cgo build would link
(INTEGRATION.md shows the cgo stubs).  numpy is only used for host buffers; all compute happens in
libkaldi_fp16.so.  Citations are to /root/reference/internal/gpu.

Where the reference builds an operation out of several launches (AddBias = K=1 GEMM,
AffineBackward* = transpose + GEMM, ...) the function keeps its name and contract but makes ONE
call into the fused tcgen05 GEMM (include/kaldi_fp16_fused.h).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import GemmDesc, K_MAJOR, MN_MAJOR, NativeError


class GPUError(NativeError):
    """Go: the ``error`` values returned by internal/gpu (opsErr / lastError, ops.go:40-47)."""


def _ops_err(what: str) -> GPUError:
    lib = _lib.load()
    msg = lib.ops_last_error()
    return GPUError(f"{what}: {msg.decode(errors='replace') if msg else 'unknown ops error'}")


def _bridge_err(what: str) -> GPUError:
    lib = _lib.load()
    msg = lib.bridge_last_error()
    return GPUError(f"{what}: {msg.decode(errors='replace') if msg else 'unknown bridge error'}")


# --------------------------------------------------------------------------- fp16 bit helpers
def float32_to_fp16_bits(f: np.ndarray) -> np.ndarray:
    """Truncating float32 -> fp16 bits used for weights / test inputs (tensor.go:158-174):
    mantissa >> 13 without rounding, |x| < 2^-14 flushed to +-0, exponent > 15 (incl. NaN) -> +-Inf."""
    bits = np.ascontiguousarray(f, dtype=np.float32).view(np.uint32)
    sign = ((bits >> 16) & 0x8000).astype(np.uint16)
    exp = ((bits >> 23) & 0xFF).astype(np.int32) - 127
    frac = bits & 0x7FFFFF
    normal = (sign | ((exp + 15).astype(np.uint16) << 10) | (frac >> 13).astype(np.uint16)).astype(np.uint16)
    out = np.where(exp > 15, sign | np.uint16(0x7C00), np.where(exp < -14, sign, normal))
    return out.astype(np.uint16)


def fp16_bits_to_float32(h: np.ndarray) -> np.ndarray:
    """tensor.go:176-203 -- exact widening, identical to IEEE half -> float."""
    return np.ascontiguousarray(h, dtype=np.uint16).view(np.float16).astype(np.float32)


# --------------------------------------------------------------------------- device init
def Init(device_id: int = 0) -> None:
    """bridge.go:45"""
    if _lib.load().bridge_gpu_init(device_id) != 0:
        raise _bridge_err("gpu init")


def MemoryInfo() -> tuple[int, int]:
    """bridge.go:50 -> (free, total) bytes"""
    f, t = C.c_size_t(), C.c_size_t()
    if _lib.load().bridge_gpu_get_free_memory(C.byref(f), C.byref(t)) != 0:
        raise _bridge_err("memory info")
    return f.value, t.value


def Sync() -> None:
    """bridge.go:59"""
    if _lib.load().bridge_gpu_sync() != 0:
        raise _bridge_err("sync")


# --------------------------------------------------------------------------- Tensor
class Tensor:
    """FP16 row-major device matrix {Ptr, Rows, Cols, Owned} (tensor.go:17-22)."""

    __slots__ = ("Ptr", "Rows", "Cols", "Owned")

    def __init__(self, ptr: int, rows: int, cols: int, owned: bool):
        self.Ptr, self.Rows, self.Cols, self.Owned = ptr, rows, cols, owned

    def Numel(self) -> int:
        return self.Rows * self.Cols

    def Bytes(self) -> int:
        return self.Numel() * 2

    def Free(self) -> None:
        if self.Owned and self.Ptr:
            _lib.load().bridge_gpu_free(self.Ptr)
        self.Ptr = None

    def View(self, row_offset: int, rows: int, cols: int) -> "Tensor":
        """Non-owning alias starting at ``row_offset`` (tensor.go:116-120)."""
        return Tensor(self.Ptr + row_offset * self.Cols * 2, rows, cols, False)

    def ToBits(self) -> np.ndarray:
        out = np.empty((self.Rows, self.Cols), dtype=np.uint16)
        if self.Numel() and _lib.load().bridge_read_fp16(out.ctypes.data, self.Ptr, self.Numel()) != 0:
            raise _bridge_err("read fp16")
        return out

    def ToFP32(self) -> np.ndarray:
        """tensor.go:94-112"""
        return fp16_bits_to_float32(self.ToBits())

    def __repr__(self) -> str:
        return f"Tensor[{self.Rows}x{self.Cols} fp16 @0x{(self.Ptr or 0):x}]"

    # zero-copy hand-over to torch (torch.as_tensor(t, device='cuda')) for torch.distributed
    @property
    def __cuda_array_interface__(self):
        return {"shape": (self.Rows, self.Cols), "typestr": "<f2", "data": (self.Ptr, False), "version": 3,
                "strides": None}


def NewTensor(rows: int, cols: int) -> Tensor:
    """tensor.go:39 (uninitialised)"""
    ptr = _lib.load().bridge_gpu_malloc(max(rows * cols * 2, 16))
    if not ptr:
        raise _bridge_err(f"alloc tensor [{rows}x{cols}]")
    return Tensor(ptr, rows, cols, True)


def ZeroTensor(rows: int, cols: int) -> Tensor:
    """tensor.go:50 (the reference uploads a host zero buffer; here a device memset)"""
    t = NewTensor(rows, cols)
    _lib.load().bridge_gpu_memset(t.Ptr, 0, max(rows * cols * 2, 1))
    return t


def TensorFromBits(bits: np.ndarray) -> Tensor:
    bits = np.ascontiguousarray(bits, dtype=np.uint16)
    rows, cols = bits.shape if bits.ndim == 2 else (1, bits.size)
    t = NewTensor(rows, cols)
    if bits.size and _lib.load().bridge_transfer_fp16(t.Ptr, bits.ctypes.data, bits.size) != 0:
        t.Free()
        raise _bridge_err("transfer fp16")
    return t


def TensorFromFP32(data: np.ndarray, rows: int, cols: int) -> Tensor:
    """tensor.go:67-91: truncating fp32 -> fp16 on the host, then H2D."""
    data = np.asarray(data, dtype=np.float32).reshape(-1)
    if data.size != rows * cols:
        raise GPUError(f"data length {data.size} != {rows}x{cols}")
    return TensorFromBits(float32_to_fp16_bits(data).reshape(rows, cols))


def TensorFromFP16(data: np.ndarray) -> Tensor:
    """Upload values that are already fp16 (np.float16), bit-exact."""
    return TensorFromBits(np.ascontiguousarray(data, dtype=np.float16).view(np.uint16))


class DeviceF32:
    """FP32 device vector (BN statistics, master weights, velocities)."""

    __slots__ = ("Ptr", "N")

    def __init__(self, host: Optional[np.ndarray] = None, n: Optional[int] = None):
        lib = _lib.load()
        self.N = int(n if host is None else np.asarray(host).size)
        self.Ptr = lib.bridge_gpu_malloc(max(self.N * 4, 16))
        if not self.Ptr:
            raise _bridge_err("alloc f32")
        if host is not None:
            h = np.ascontiguousarray(host, dtype=np.float32).reshape(-1)
            if h.size and lib.bridge_transfer_float32(self.Ptr, h.ctypes.data, h.size) != 0:
                raise _bridge_err("transfer f32")
        else:
            lib.bridge_gpu_memset(self.Ptr, 0, max(self.N * 4, 1))

    def ToHost(self) -> np.ndarray:
        out = np.empty(self.N, dtype=np.float32)
        if self.N and _lib.load().bridge_read_float32(out.ctypes.data, self.Ptr, self.N) != 0:
            raise _bridge_err("read f32")
        return out

    def Free(self) -> None:
        if self.Ptr:
            _lib.load().bridge_gpu_free(self.Ptr)
        self.Ptr = None


# --------------------------------------------------------------------------- Handle / GEMM
class Handle:
    """ops.go:21-37.  The opaque pointer is a kfp16 context (device, stream, workspace)."""

    def __init__(self):
        self.ptr = _lib.load().ops_cublas_create()
        if not self.ptr:
            raise _ops_err("cublas create")

    def Destroy(self) -> None:
        if self.ptr:
            _lib.load().ops_cublas_destroy(self.ptr)
        self.ptr = None


def NewHandle() -> Handle:
    return Handle()


def GEMM(h: Handle, M: int, N: int, K: int, alpha: float, A: Tensor, B: Tensor, beta: float, C_: Tensor) -> None:
    """C[MxN] = alpha*A[MxK]*B[KxN] + beta*C (ops.go:55-70)."""
    if _lib.load().ops_gemm(h.ptr, M, N, K, alpha, A.Ptr, K, B.Ptr, N, beta, C_.Ptr, N) != 0:
        raise _ops_err("gemm")


def GEMMSimple(h: Handle, A: Tensor, B: Tensor, C_: Tensor) -> None:
    """ops.go:72"""
    GEMM(h, A.Rows, B.Cols, A.Cols, 1.0, A, B, 0.0, C_)


def GEMMAcc(h: Handle, A: Tensor, B: Tensor, C_: Tensor) -> None:
    """ops.go:77"""
    GEMM(h, A.Rows, B.Cols, A.Cols, 1.0, A, B, 1.0, C_)


def _unary(fn_name: str, what: str, t: Tensor, *extra) -> None:
    if getattr(_lib.load(), fn_name)(t.Ptr, t.Numel(), *extra) != 0:
        raise _ops_err(what)


def ReLU(t: Tensor) -> None:
    _unary("ops_relu", "relu", t)


def Sigmoid(t: Tensor) -> None:
    _unary("ops_sigmoid", "sigmoid", t)


def Tanh(t: Tensor) -> None:
    _unary("ops_tanh_act", "tanh", t)


def ClippedReLU(t: Tensor, ceiling: float) -> None:
    _unary("ops_clipped_relu", "clipped_relu", t, ceiling)


def Softmax(t: Tensor) -> None:
    if _lib.load().ops_softmax(t.Ptr, t.Rows, t.Cols) != 0:
        raise _ops_err("softmax")


def LogSoftmax(t: Tensor) -> None:
    if _lib.load().ops_log_softmax(t.Ptr, t.Rows, t.Cols) != 0:
        raise _ops_err("log_softmax")


@dataclass
class BNParams:
    """ops.go:134-201: FP32 device vectors, eps default 0.001."""

    Mean: DeviceF32
    Var: DeviceF32
    Gamma: DeviceF32
    Beta: DeviceF32
    Dim: int
    Epsilon: float = 0.001

    def Free(self) -> None:
        for v in (self.Mean, self.Var, self.Gamma, self.Beta):
            v.Free()


def NewBNParams(mean: Sequence[float], var: Sequence[float], gamma: Sequence[float], beta: Sequence[float]) -> BNParams:
    d = len(mean)
    if not (len(var) == d and len(gamma) == d and len(beta) == d):
        raise GPUError("BN param dimension mismatch")
    return BNParams(DeviceF32(np.asarray(mean)), DeviceF32(np.asarray(var)), DeviceF32(np.asarray(gamma)),
                    DeviceF32(np.asarray(beta)), d)


def BatchNormForward(x: Tensor, bn: BNParams, eps: float) -> None:
    """ops.go:204"""
    if x.Cols != bn.Dim:
        raise GPUError(f"BatchNorm: x cols {x.Cols} != bn dim {bn.Dim}")
    if _lib.load().ops_batchnorm_forward(x.Ptr, x.Rows, x.Cols, bn.Mean.Ptr, bn.Var.Ptr, bn.Gamma.Ptr, bn.Beta.Ptr, eps) != 0:
        raise _ops_err("batchnorm")


def BatchNormForwardRMS(x: Tensor, mean: DeviceF32, var: DeviceF32, target_rms: float, eps: float) -> None:
    """ops.go:218"""
    if _lib.load().ops_batchnorm_forward_rms(x.Ptr, x.Rows, x.Cols, mean.Ptr, var.Ptr, target_rms, eps) != 0:
        raise _ops_err("batchnorm_rms")


def AddScaled(dst: Tensor, src: Tensor, alpha: float, beta: float) -> None:
    """dst = alpha*src + beta*dst (ops.go:235)"""
    if dst.Numel() != src.Numel():
        raise GPUError("AddScaled: size mismatch")
    if _lib.load().ops_add_scaled(dst.Ptr, src.Ptr, dst.Numel(), alpha, beta) != 0:
        raise _ops_err("add_scaled")


def Add(dst: Tensor, src: Tensor) -> None:
    if dst.Numel() != src.Numel():
        raise GPUError("Add: size mismatch")
    if _lib.load().ops_add(dst.Ptr, src.Ptr, dst.Numel()) != 0:
        raise _ops_err("add")


def Copy(dst: Tensor, src: Tensor) -> None:
    if dst.Numel() != src.Numel():
        raise GPUError("Copy: size mismatch")
    if _lib.load().ops_copy(dst.Ptr, src.Ptr, dst.Numel()) != 0:
        raise _ops_err("copy")


def Fill(t: Tensor, val: float) -> None:
    if _lib.load().ops_fill(t.Ptr, t.Numel(), val) != 0:
        raise _ops_err("fill")


def ConcatCols(dst: Tensor, src: Tensor, col_offset: int) -> None:
    """ops.go:280"""
    if dst.Rows != src.Rows or col_offset + src.Cols > dst.Cols:
        raise GPUError("ConcatCols: shape mismatch")
    if _lib.load().ops_concat_cols(dst.Ptr, dst.Rows, dst.Cols, src.Ptr, src.Cols, col_offset) != 0:
        raise _ops_err("concat_cols")


def SliceCols(src: Tensor, dst: Tensor, col_offset: int) -> None:
    """ops.go:296"""
    if dst.Rows != src.Rows or col_offset + dst.Cols > src.Cols:
        raise GPUError("SliceCols: shape mismatch")
    if _lib.load().ops_slice_cols(src.Ptr, src.Rows, src.Cols, dst.Ptr, dst.Cols, col_offset) != 0:
        raise _ops_err("slice_cols")


def CombineFeatureMaps(t: Tensor, height: int, nf1: int, nf2: int) -> None:
    """ops.go:312"""
    if _lib.load().ops_combine_feature_maps(t.Ptr, t.Rows, t.Cols, height, nf1, nf2) != 0:
        raise _ops_err("combine_feature_maps")


def AddBias(h: Handle, x: Tensor, bias: Tensor) -> None:
    """x[t,:] += bias (ops.go:335-351).  The reference allocates a ones[Tx1] column and runs a K=1
    GEMM with beta=1; the result (one fp16 rounding of x+bias) is the same as this single pass."""
    if bias.Numel() != x.Cols:
        raise GPUError(f"AddBias: bias {bias.Numel()} != cols {x.Cols}")
    lib = _lib.load()
    if lib.kfp16_add_bias(h.ptr, x.Ptr, x.Cols, bias.Ptr, x.Rows, x.Cols) != 0:
        raise _ops_err("add_bias")


def SubsampleRows(src: Tensor, stride: int, row_offset: int) -> Tensor:
    """ops.go:355"""
    out_rows = (src.Rows - row_offset + stride - 1) // stride
    dst = NewTensor(out_rows, src.Cols)
    _lib.load().ops_subsample_rows(dst.Ptr, src.Ptr, src.Rows, src.Cols, stride, row_offset)
    return dst


# --------------------------------------------------------------------------- backward ops
def ReLUBackward(x: Tensor, grad: Tensor) -> None:
    """backward_ops.go:42"""
    if x.Numel() != grad.Numel():
        raise GPUError("ReLUBackward: size mismatch")
    if _lib.load().ops_relu_backward(x.Ptr, grad.Ptr, x.Numel()) != 0:
        raise _ops_err("relu_backward")


def SigmoidBackward(out: Tensor, grad: Tensor) -> None:
    if _lib.load().ops_sigmoid_backward(out.Ptr, grad.Ptr, out.Numel()) != 0:
        raise _ops_err("sigmoid_backward")


def TanhBackward(out: Tensor, grad: Tensor) -> None:
    if _lib.load().ops_tanh_backward(out.Ptr, grad.Ptr, out.Numel()) != 0:
        raise _ops_err("tanh_backward")


def BatchNormBackward(grad_out: Tensor, grad_in: Tensor, bn: BNParams, eps: float) -> None:
    """backward_ops.go:78"""
    if _lib.load().ops_batchnorm_backward(grad_out.Ptr, grad_in.Ptr, bn.Gamma.Ptr, bn.Var.Ptr, eps, grad_out.Rows, grad_out.Cols) != 0:
        raise _ops_err("batchnorm_backward")


def Transpose(src: Tensor, dst: Tensor) -> None:
    """backward_ops.go:97"""
    if src.Rows != dst.Cols or src.Cols != dst.Rows:
        raise GPUError(f"Transpose: src [{src.Rows}x{src.Cols}] dst [{dst.Rows}x{dst.Cols}] mismatch")
    if _lib.load().ops_transpose(src.Ptr, dst.Ptr, src.Rows, src.Cols) != 0:
        raise _ops_err("transpose")


def AffineBackwardData(h: Handle, grad_output: Tensor, weight: Tensor) -> Tensor:
    """gradInput[TxM] = gradOutput[TxK] * W^T, W [MxK] (backward_ops.go:162-192).
    One NT GEMM: the UMMA descriptor reads W as the K-major B operand, no transpose kernel."""
    T, K, M = grad_output.Rows, grad_output.Cols, weight.Rows
    if weight.Cols != K:
        raise GPUError(f"AffineBackwardData: weight cols {weight.Cols} != gradOutput cols {K}")
    out = NewTensor(T, M)
    if _lib.load().kfp16_gemm(h.ptr, T, M, K, 1.0, grad_output.Ptr, 0, weight.Ptr, 1, 0.0, out.Ptr) != 0:
        out.Free()
        raise _ops_err("GEMM gradInput")
    return out


def AffineBackwardWeights(h: Handle, input: Tensor, grad_output: Tensor) -> Tensor:
    """gradW[MxK] = input^T[MxT] * gradOutput[TxK] (backward_ops.go:195-225).  One TN GEMM."""
    T, K, M = grad_output.Rows, grad_output.Cols, input.Cols
    if input.Rows != T:
        raise GPUError(f"AffineBackwardWeights: input rows {input.Rows} != gradOutput rows {T}")
    out = NewTensor(M, K)
    if _lib.load().kfp16_gemm(h.ptr, M, K, T, 1.0, input.Ptr, 1, grad_output.Ptr, 0, 0.0, out.Ptr) != 0:
        out.Free()
        raise _ops_err("GEMM gradWeights")
    return out


def AffineBackwardBias(h: Handle, grad_output: Tensor) -> Tensor:
    """gradBias[1xK] = sum_t gradOutput[t,:] (backward_ops.go:228-253), fp32 accumulate."""
    out = NewTensor(1, grad_output.Cols)
    tmp = DeviceF32(n=grad_output.Cols)
    try:
        if _lib.load().kfp16_colsum(h.ptr, grad_output.Ptr, grad_output.Cols, grad_output.Rows, grad_output.Cols, tmp.Ptr, out.Ptr) != 0:
            out.Free()
            raise _ops_err("bias grad")
        Sync()
    finally:
        tmp.Free()
    return out


# --------------------------------------------------------------------------- optimiser
class SGDOptimizer:
    """optimize.go:30-133: FP32 master weights + velocity per (layer, param)."""

    def __init__(self, lr: float, momentum: float):
        self.LR, self.Momentum = lr, momentum
        self.MasterWeights: dict[str, DeviceF32] = {}
        self.Velocities: dict[str, DeviceF32] = {}
        self.ParamSizes: dict[str, int] = {}

    def RegisterParam(self, layer: str, param: str, w: Tensor) -> None:
        key = f"{layer}.{param}"
        n = w.Numel()
        master = DeviceF32(n=n)
        if _lib.load().ops_fp16_to_fp32(w.Ptr, master.Ptr, n) != 0:
            raise _ops_err("register param")
        self.MasterWeights[key] = master
        self.Velocities[key] = DeviceF32(n=n)
        self.ParamSizes[key] = n

    def Update(self, layer: str, param: str, w: Tensor, grad: Tensor) -> None:
        key = f"{layer}.{param}"
        if key not in self.MasterWeights:
            raise GPUError(f"param {key} not registered")
        n = self.ParamSizes[key]
        if grad.Numel() != n:
            raise GPUError(f"grad size {grad.Numel()} != param size {n}")
        if _lib.load().ops_sgd_update(self.MasterWeights[key].Ptr, w.Ptr, grad.Ptr, self.Velocities[key].Ptr,
                                      self.LR, self.Momentum, n) != 0:
            raise _ops_err("sgd_update")

    def SetLR(self, lr: float) -> None:
        self.LR = lr

    def Free(self) -> None:
        for d in (self.MasterWeights, self.Velocities):
            for v in d.values():
                v.Free()
            d.clear()


def NewSGDOptimizer(lr: float, momentum: float) -> SGDOptimizer:
    return SGDOptimizer(lr, momentum)


# --------------------------------------------------------------------------- batch transfer (bridge.go:68-221)
@dataclass
class TrainingBatch:
    """The fields of loader.TrainingBatch that TransferBatch consumes (internal/loader/dataloader.go:15-38):
    dense features [total_frames x feat_dim], optional ivectors [batch_size x ivec_dim], merged numerator FST in CSR."""

    features: np.ndarray
    batch_size: int
    ivectors: Optional[np.ndarray] = None
    csr_row_ptr: Optional[np.ndarray] = None     # int32 [num_states + 1]
    csr_col_idx: Optional[np.ndarray] = None     # int32 [num_arcs]
    csr_labels: Optional[np.ndarray] = None      # int32 [num_arcs]
    csr_weights: Optional[np.ndarray] = None     # float32 [num_arcs]


def _align256(n: int) -> int:
    return (n + 255) & ~255


def pack_batch(tb: TrainingBatch) -> tuple[np.ndarray, dict]:
    """Host side of TransferBatch (bridge.go:123-221): features / ivectors go through the round-to-nearest-even
    converter (internal/fp16/fp16.go:13-70 via bridge.go:141), then everything is packed into ONE buffer
    [feat fp16 | ivec fp16 | row_ptr i32 | col_idx i32 | labels i32 | weights f32], each section 256-byte aligned."""
    with np.errstate(over="ignore"):
        feat = np.ascontiguousarray(tb.features, dtype=np.float32).astype(np.float16)
        ivec = (np.ascontiguousarray(tb.ivectors, dtype=np.float32).astype(np.float16) if tb.ivectors is not None
                else np.zeros((tb.batch_size, 0), np.float16))
    rp = np.ascontiguousarray(tb.csr_row_ptr if tb.csr_row_ptr is not None else [0], dtype=np.int32)
    ci = np.ascontiguousarray(tb.csr_col_idx if tb.csr_col_idx is not None else [], dtype=np.int32)
    lb = np.ascontiguousarray(tb.csr_labels if tb.csr_labels is not None else [], dtype=np.int32)
    wt = np.ascontiguousarray(tb.csr_weights if tb.csr_weights is not None else [], dtype=np.float32)
    if not (ci.size == lb.size == wt.size):
        raise ValueError("CSR col_idx / labels / weights must have one entry per arc")
    sections = [("features", feat), ("ivectors", ivec), ("csr_row_ptr", rp), ("csr_col_idx", ci), ("csr_labels", lb),
                ("csr_weights", wt)]
    offsets, off = {}, 0
    for name, a in sections:
        offsets[name] = off
        off += _align256(a.nbytes)
    buf = np.zeros(off, dtype=np.uint8)
    for name, a in sections:
        buf[offsets[name]: offsets[name] + a.nbytes] = a.view(np.uint8).reshape(-1)
    meta = dict(total_frames=feat.shape[0], feat_dim=feat.shape[1], batch_size=tb.batch_size, ivec_dim=ivec.shape[1],
                num_states=rp.size - 1, num_arcs=ci.size, offsets=offsets, total_bytes=off)
    return buf, meta


class GPUBatch:
    """bridge.go:68-112: one device allocation holding the whole minibatch"""

    def __init__(self, ptrs, meta):
        self.ptrs, self.meta = ptrs, meta
        self.TotalFrames, self.FeatDim = meta["total_frames"], meta["feat_dim"]
        self.BatchSize, self.IvecDim = meta["batch_size"], meta["ivec_dim"]
        self.NumStates, self.NumArcs = meta["num_states"], meta["num_arcs"]

    def Features(self) -> Tensor:
        return Tensor(self.ptrs.d_features, self.TotalFrames, self.FeatDim, False)

    def Ivectors(self) -> Tensor:
        return Tensor(self.ptrs.d_ivectors, self.BatchSize, self.IvecDim, False)

    def TotalBytes(self) -> int:
        return int(self.ptrs.total_bytes)

    def Free(self) -> None:
        _lib.load().bridge_batch_free(C.byref(self.ptrs))


def TransferBatch(tb: TrainingBatch) -> GPUBatch:
    """bridge.go:123-221: convert, pack, ONE cudaMalloc + ONE cudaMemcpy (bridge.cu:206-267)"""
    lib = _lib.load()
    buf, meta = pack_batch(tb)
    ptrs = _lib.GPUBatchPtrs()
    if lib.bridge_batch_alloc(meta["total_frames"], meta["feat_dim"], meta["batch_size"], meta["ivec_dim"], meta["num_states"],
                              meta["num_arcs"], C.byref(ptrs)) != 0:
        raise _bridge_err("GPU alloc failed")
    if int(ptrs.total_bytes) != meta["total_bytes"]:
        lib.bridge_batch_free(C.byref(ptrs))
        raise GPUError(f"batch layout mismatch: device {int(ptrs.total_bytes)} bytes, host {meta['total_bytes']}")
    if lib.bridge_batch_transfer(C.byref(ptrs), buf.ctypes.data, buf.nbytes) != 0:
        lib.bridge_batch_free(C.byref(ptrs))
        raise _bridge_err("GPU transfer failed")
    return GPUBatch(ptrs, meta)
