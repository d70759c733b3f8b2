"""Host-side mirror of the reference's Go package ``internal/nnet`` over the C ABI
(include/kaldi_fp16_nnet.h).  Same vocabulary as the reference: a model is xconfig text
(internal/nnet/xconfig.go:143, model.go:22-65), ``NewNetwork`` instantiates it with random weights
(forward.go:111,1008-1187), ``Forward`` / ``Backward`` run it (forward.go:148,
network_backward.go:94) and ``Trainer.Step`` is TrainStep (train_step.go:41-283) with the
``0.5*||out||^2`` objective the acceptance mains use (cmd/sgdtest/main.go:258-267).

All compute happens in libkaldi_fp16.so; numpy only carries host buffers.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _lib, gpu
from ._lib import NetOpts


class NNetError(gpu.GPUError):
    pass


def _err(what: str) -> NNetError:
    return NNetError(f"{what}: {_lib.last_error()}")


@dataclass
class Model:
    """model.go:22-65 BuildModelFromString: here just the validated xconfig text."""

    xconfig: str


def BuildModelFromString(xconfig: str) -> Model:
    return Model(xconfig)


def BuildModel(path: str) -> Model:
    with open(path) as f:
        return Model(f.read())


def rne_fp16_bits(x: np.ndarray) -> np.ndarray:
    """features go through round-to-nearest-even (internal/fp16/fp16.go:13-70 via bridge.go:141);
    numpy's float32->float16 cast is the same IEEE conversion."""
    with np.errstate(over="ignore"):
        return np.ascontiguousarray(x, dtype=np.float32).astype(np.float16).view(np.uint16)


class Network:
    """forward.go:111 NewNetwork + Forward / Backward; owns the native executor."""

    def __init__(self, model: Model, handle: gpu.Handle, n_seq: int, seq_len: int, train: bool = True,
                 lr: float = 1e-3, momentum: float = 0.9, ref_round: bool = False, seed: Optional[int] = 42,
                 grad_scale: float = 1.0, round_grad: bool = True):
        self.lib = _lib.load()
        self.handle = handle
        self.n_seq, self.seq_len = n_seq, seq_len
        opts = NetOpts(n_seq=n_seq, seq_len=seq_len, ref_round=int(ref_round), train=int(train), lr=lr,
                       momentum=momentum, conv_cartesian=1, grad_scale=grad_scale, round_grad=int(round_grad))
        self.ptr = self.lib.kfp16_net_create(handle.ptr, model.xconfig.encode(), C.byref(opts))
        if not self.ptr:
            raise _err("NewNetwork")
        if seed is not None and seed != 42:
            if self.lib.kfp16_net_init_random(self.ptr, seed) != 0:
                raise _err("init_random")
        self.layers = [(self.lib.kfp16_net_layer_name(self.ptr, i).decode(),
                        self.lib.kfp16_net_layer_type(self.ptr, i).decode(),
                        self.lib.kfp16_net_layer_dim(self.ptr, i)) for i in range(self.lib.kfp16_net_num_layers(self.ptr))]
        self.params = {}
        for i in range(self.lib.kfp16_net_num_params(self.ptr)):
            r, c = C.c_int(), C.c_int()
            self.lib.kfp16_net_param_shape(self.ptr, i, C.byref(r), C.byref(c))
            self.params[self.lib.kfp16_net_param_name(self.ptr, i).decode()] = (
                r.value, c.value, self.lib.kfp16_net_param_offset(self.ptr, i))

    # ---- lifetime
    def Free(self) -> None:
        if self.ptr:
            self.lib.kfp16_net_destroy(self.ptr)
        self.ptr = None

    @property
    def T(self) -> int:
        return self.n_seq * self.seq_len

    def layer_dim(self, name: str) -> int:
        for n, _, d in self.layers:
            if n == name:
                return d
        raise NNetError(f"no layer {name}")

    # ---- parameters
    def SetParam(self, name: str, w: np.ndarray) -> None:
        r, c, _ = self.params[name]
        w = np.ascontiguousarray(w, dtype=np.float32).reshape(r, c)
        if self.lib.kfp16_net_set_param(self.ptr, name.encode(), w.ctypes.data, r, c) != 0:
            raise _err("SetParam")

    def GetParam(self, name: str) -> np.ndarray:
        r, c, _ = self.params[name]
        out = np.empty((r, c), dtype=np.uint16)
        if self.lib.kfp16_net_get_param(self.ptr, name.encode(), out.ctypes.data, r, c) != 0:
            raise _err("GetParam")
        return out.view(np.float16).astype(np.float32)

    def SetBN(self, layer: str, which: str, mean, var, gamma=None, beta=None, eps: float = 1e-3) -> None:
        mean = np.ascontiguousarray(mean, dtype=np.float32)
        var = np.ascontiguousarray(var, dtype=np.float32)
        g = None if gamma is None else np.ascontiguousarray(gamma, dtype=np.float32)
        b = None if beta is None else np.ascontiguousarray(beta, dtype=np.float32)
        if self.lib.kfp16_net_set_bn(self.ptr, layer.encode(), which.encode(), mean.ctypes.data, var.ctypes.data,
                                     None if g is None else g.ctypes.data, None if b is None else b.ctypes.data,
                                     eps, mean.size) != 0:
            raise _err("SetBN")

    def SetIDCT(self, layer: str, m: np.ndarray) -> None:
        """idct-layer matrix [in x out] from a loaded model (weight_loader.go:766-776) instead of makeIDCTMatrix"""
        m = np.ascontiguousarray(m, dtype=np.float32)
        if m.ndim != 2 or m.shape[0] != m.shape[1] or self.lib.kfp16_net_set_idct(self.ptr, layer.encode(), m.ctypes.data, m.shape[0]) != 0:
            raise _err("SetIDCT")

    def GetBN(self, layer: str, which: str, dim: int):
        """running mean / variance of a batch-norm (what train-mode batch-norm updates)"""
        mean, var = np.empty(dim, np.float32), np.empty(dim, np.float32)
        if self.lib.kfp16_net_get_bn(self.ptr, layer.encode(), which.encode(), mean.ctypes.data, var.ctypes.data, dim) != 0:
            raise _err("GetBN")
        return mean, var

    def SetTrainBatchNorm(self, on: bool, momentum: float = 0.1) -> None:
        """batch statistics instead of running statistics in every batch-norm of a training network
        (cpp/cuda/cnn_kernels.cu:236-320 training branch, go/gotorch/layers.go:257-330)"""
        if self.lib.kfp16_net_set_train_batchnorm(self.ptr, 1 if on else 0, float(momentum)) != 0:
            raise _err("SetTrainBatchNorm")

    def SetBNStatsHook(self, fn, world: int) -> None:
        """data-parallel batch-norm: fn(stats_ptr: int, count: int, stream_ptr: int) must sum the fp32 device vector over the
        `world` ranks, stream-ordered; fn = None removes the hook (see dp.BNStatsAllReducer)"""
        if fn is None:
            self._bn_hook = None
            rc = self.lib.kfp16_net_set_bn_stats_hook(self.ptr, None, None, 1)
        else:
            proto = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p)

            def tramp(_user, stats, count, stream):
                try:
                    fn(int(stats or 0), int(count), int(stream or 0))
                    return 0
                except Exception:  # noqa: BLE001 - reported through the C error path
                    return -1

            self._bn_hook = proto(tramp)          # keep the callback alive
            rc = self.lib.kfp16_net_set_bn_stats_hook(self.ptr, C.cast(self._bn_hook, C.c_void_p), None, int(world))
        if rc != 0:
            raise _err("SetBNStatsHook")

    def _bucket_f32(self, getter) -> np.ndarray:
        n = self.lib.kfp16_net_bucket_size(self.ptr)
        out = np.empty(n, dtype=np.float32)
        gpu.Sync()
        if n and self.lib.bridge_read_float32(out.ctypes.data, getter(self.ptr), n) != 0:
            raise _err("read bucket")
        return out

    def WeightGrads(self) -> dict:
        """bwdState.WeightGrads (network_backward.go): name -> fp32 [rows x cols]"""
        flat = self._bucket_f32(self.lib.kfp16_net_grads_f32)
        return {k: flat[o:o + r * c].reshape(r, c).copy() for k, (r, c, o) in self.params.items()}

    def MasterWeights(self) -> dict:
        flat = self._bucket_f32(self.lib.kfp16_net_params_f32)
        return {k: flat[o:o + r * c].reshape(r, c).copy() for k, (r, c, o) in self.params.items()}

    # ---- data
    def SetInput(self, name: str, x: np.ndarray, already_fp16_bits: bool = False) -> None:
        bits = np.ascontiguousarray(x, dtype=np.uint16) if already_fp16_bits else rne_fp16_bits(x)
        if self.lib.kfp16_net_set_input(self.ptr, name.encode(), bits.ctypes.data, bits.shape[0], bits.shape[1]) != 0:
            raise _err("SetInput")

    def SetInputF32(self, name: str, x: np.ndarray) -> None:
        """FP32 rows uploaded as they are; the RNE FP32 -> FP16 conversion (bridge.go:141) runs on the device"""
        x = np.ascontiguousarray(x, dtype=np.float32)
        if self.lib.kfp16_net_set_input_f32(self.ptr, name.encode(), x.ctypes.data, x.shape[0], x.shape[1]) != 0:
            raise _err("SetInputF32")

    def SetInputCompressed(self, name: str, mats) -> None:
        """egs feature matrices as they sit in the archive, one per sequence: ``mats`` is a list of
        (format, payload bytes, global_min, global_range) with format in {"CM", "CM2", "CM3", "FM"} (parser.go:300-365);
        the payload is uploaded and decoded + converted to FP16 on the device"""
        fmt = {"CM": 1, "CM2": 2, "CM3": 3, "FM": 4}
        cols = self.layer_dim(name)
        descs = (_lib.CmDesc * len(mats))()
        blob = bytearray()
        for i, (f, payload, gmin, grange) in enumerate(mats):
            if len(blob) & 1:
                blob.append(0)
            d = descs[i]
            d.format, d.rows, d.cols, d.global_min, d.global_range = fmt[f], self.seq_len, cols, gmin, grange
            d.payload_offset, d.dst_row = len(blob), 0
            blob += payload
        buf = (C.c_ubyte * len(blob)).from_buffer(blob)
        if self.lib.kfp16_net_set_input_compressed(self.ptr, name.encode(), buf, len(blob), descs, len(mats)) != 0:
            raise _err("SetInputCompressed")

    def PrefetchInputF32(self, name: str, host_ptr: int, rows: int, cols: int) -> None:
        """async H2D of the next minibatch's FP32 rows from pinned host memory; CommitInput converts on the device"""
        if self.lib.kfp16_net_prefetch_input_f32(self.ptr, name.encode(), host_ptr, rows, cols) != 0:
            raise _err("PrefetchInputF32")

    def PrefetchInput(self, name: str, host_ptr: int, rows: int, cols: int) -> None:
        """async H2D of the NEXT minibatch from pinned host memory (gpu.TransferBatchPinned done asynchronously)"""
        if self.lib.kfp16_net_prefetch_input(self.ptr, name.encode(), host_ptr, rows, cols) != 0:
            raise _err("PrefetchInput")

    def CommitInput(self, name: str) -> None:
        if self.lib.kfp16_net_commit_input(self.ptr, name.encode()) != 0:
            raise _err("CommitInput")

    def Forward(self, features: np.ndarray, ivectors: Optional[np.ndarray] = None, input_name: str = "input",
                ivector_name: str = "ivector") -> np.ndarray:
        """forward.go:148: returns the activation of the layer named ``output`` (dense real rows, fp32)."""
        self.SetInput(input_name, features)
        if ivectors is not None:
            self.SetInput(ivector_name, ivectors)
        if self.lib.kfp16_net_forward(self.ptr) != 0:
            raise _err("Forward")
        return self.Output("")

    def Output(self, layer: str = "") -> np.ndarray:
        dim = self.layer_dim(layer) if layer else self.layers[self._out_index()][2]
        rows = self._rows_of(layer)
        out = np.empty((rows, dim), dtype=np.uint16)
        if self.lib.kfp16_net_get_output(self.ptr, layer.encode(), out.ctypes.data, rows, dim) != 0:
            raise _err("Output")
        return out.view(np.float16).astype(np.float32)

    def Grad(self, layer: str) -> np.ndarray:
        dim = self.layer_dim(layer)
        rows = self._rows_of(layer)
        out = np.empty((rows, dim), dtype=np.uint16)
        if self.lib.kfp16_net_get_grad(self.ptr, layer.encode(), out.ctypes.data, rows, dim) != 0:
            raise _err("Grad")
        return out.view(np.float16).astype(np.float32)

    def Mask(self, layer: str, dim: int) -> np.ndarray:
        """ReLU mask saved by the fused epilogue of a tdnnf / prefinal layer (bool [T x dim])"""
        rows = self._rows_of(layer)
        out = np.empty((rows, dim), dtype=np.uint8)
        if self.lib.kfp16_net_get_mask(self.ptr, layer.encode(), out.ctypes.data, rows, dim) != 0:
            raise _err("Mask")
        return out.astype(bool)

    def _out_index(self) -> int:
        for i, (n, _, _) in enumerate(self.layers):
            if n == "output":
                return i
        for i, (_, t, _) in enumerate(self.layers):
            if t == "output-layer":
                return i
        return len(self.layers) - 1

    def _rows_of(self, layer: str) -> int:
        # per-sequence layers (ivector branch) have n_seq rows; probe through the error-free path
        return self._per_seq_rows.get(layer, self.T) if hasattr(self, "_per_seq_rows") else self.T

    def MarkPerSequence(self, *layers: str) -> None:
        self._per_seq_rows = {l: self.n_seq for l in layers}

    # ---- training
    def ZeroGrads(self) -> None:
        if self.lib.kfp16_net_zero_grads(self.ptr) != 0:
            raise _err("ZeroGrads")

    def Backward(self, output_grad: Optional[np.ndarray] = None) -> None:
        """network_backward.go:94.  output_grad None = dOut = out (0.5*||out||^2 objective)."""
        if output_grad is None:
            if self.lib.kfp16_net_loss_half_sq(self.ptr, b"") != 0:
                raise _err("loss")
        else:
            bits = rne_fp16_bits(output_grad)
            if self.lib.kfp16_net_set_output_grad(self.ptr, b"", bits.ctypes.data, bits.shape[0], bits.shape[1]) != 0:
                raise _err("set_output_grad")
        if self.lib.kfp16_net_backward(self.ptr) != 0:
            raise _err("Backward")

    def ReadLoss(self) -> float:
        v = C.c_float(0)
        if self.lib.kfp16_net_read_loss(self.ptr, C.byref(v)) != 0:
            raise _err("ReadLoss")
        return v.value

    def CaptureSegments(self, nseg: int, cut_layers: Optional[str] = None, export_f16: bool = False,
                        tail_max_ctas: int = 0) -> int:
        """cut the step graph along the backward pass (bucketed gradient all-reduce); returns the segment count"""
        k = self.lib.kfp16_net_capture_segments_ex(self.ptr, nseg, cut_layers.encode() if cut_layers else None,
                                                   int(export_f16), tail_max_ctas)
        if k < 0:
            raise _err("CaptureSegments")
        return k

    def LaunchSegment(self, seg: int) -> None:
        if self.lib.kfp16_net_launch_segment(self.ptr, seg) != 0:
            raise _err("LaunchSegment")

    def SegmentGrads(self, seg: int) -> tuple[int, int]:
        """(first element, count) of the gradient bucket that is final once segment `seg` has run"""
        a, b = C.c_size_t(0), C.c_size_t(0)
        if self.lib.kfp16_net_segment_grads(self.ptr, seg, C.byref(a), C.byref(b)) != 0:
            raise _err("SegmentGrads")
        return a.value, b.value

    def ReadLossAsync(self, slot: int) -> None:
        """queue the loss download behind the work already in the stream (pinned slot 0/1)"""
        if self.lib.kfp16_net_read_loss_async(self.ptr, slot) != 0:
            raise _err("ReadLossAsync")

    def WaitLoss(self, slot: int) -> float:
        v = C.c_float(0)
        if self.lib.kfp16_net_wait_loss(self.ptr, slot, C.byref(v)) != 0:
            raise _err("WaitLoss")
        return v.value

    def SGDStep(self, grad_scale: float = 1.0, round_grad: bool = True) -> None:
        if self.lib.kfp16_net_sgd_step(self.ptr, grad_scale, int(round_grad)) != 0:
            raise _err("SGDStep")

    def GradsToF16(self) -> None:
        """g16 = half(g32 * grad_scale): the reference's FP16 gradient tensors as one flat bucket"""
        if self.lib.kfp16_net_grads_to_f16(self.ptr) != 0:
            raise _err("GradsToF16")

    def SGDStepF16(self) -> None:
        if self.lib.kfp16_net_sgd_step_f16(self.ptr) != 0:
            raise _err("SGDStepF16")

    def SetLR(self, lr: float) -> None:
        """SGDOptimizer.SetLR (optimize.go:123); also takes effect for captured graphs"""
        if self.lib.kfp16_net_set_lr(self.ptr, lr) != 0:
            raise _err("SetLR")

    def SetMomentum(self, momentum: float) -> None:
        if self.lib.kfp16_net_set_momentum(self.ptr, momentum) != 0:
            raise _err("SetMomentum")

    def SetSparseOutputGrad(self, on: bool) -> None:
        """chain objective with frame subsampling: back-propagate only the output frames' rows through the row-wise layers
        behind the output (default on; same results as the dense form)"""
        if self.lib.kfp16_net_set_sparse_output_grad(self.ptr, 1 if on else 0) != 0:
            raise _err("SetSparseOutputGrad")

    def SetSpecAugment(self, on: bool) -> None:
        """spec-augment-layer masks in training (go/gotorch/cnn_tdnn.go:612-668); off = the reference executor's pass-through"""
        if self.lib.kfp16_net_set_spec_augment(self.ptr, 1 if on else 0) != 0:
            raise _err("SetSpecAugment")

    def SetOverlapLoss(self, on: bool) -> None:
        if self.lib.kfp16_net_set_overlap_loss(self.ptr, 1 if on else 0) != 0:
            raise _err("SetOverlapLoss")

    def SetFuseConvBackward(self, on: bool) -> None:
        """conv layers with one consumer take dZ from that consumer's input-gradient epilogue (default on); Grad(conv layer)
        then returns dZ instead of dY"""
        if self.lib.kfp16_net_set_fuse_conv_backward(self.ptr, 1 if on else 0) != 0:
            raise _err("SetFuseConvBackward")

    def Capture(self, phases: int = 3) -> None:
        if self.lib.kfp16_net_capture(self.ptr, phases) != 0:
            raise _err("Capture")

    def Launch(self, phases: int = 3) -> None:
        if self.lib.kfp16_net_launch(self.ptr, phases) != 0:
            raise _err("Launch")

    def grads_as_cuda_array(self, f16: bool = False):
        """flat gradient bucket (FP32, or the FP16 export) as a __cuda_array_interface__ object (for torch.distributed)."""
        n = self.lib.kfp16_net_bucket_size(self.ptr)
        ptr = self.lib.kfp16_net_grads_f16(self.ptr) if f16 else self.lib.kfp16_net_grads_f32(self.ptr)

        class _Arr:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f2" if f16 else "<f4", "data": (ptr, False),
                                        "version": 3, "strides": None}
        return _Arr()

    def params_as_cuda_array(self):
        """flat FP32 master weights (cross-rank equality checks in the data-parallel loop)"""
        n = self.lib.kfp16_net_bucket_size(self.ptr)
        ptr = self.lib.kfp16_net_params_f32(self.ptr)

        class _Arr:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 3,
                                        "strides": None}
        return _Arr()


def NewNetwork(model: Model, handle: gpu.Handle, n_seq: int, seq_len: int, **kw) -> Network:
    return Network(model, handle, n_seq, seq_len, **kw)


class Trainer:
    """train_step.go:41-140 Trainer: network + SGD optimiser; Step = one minibatch."""

    def __init__(self, net: Network):
        self.net = net

    def Step(self, features: np.ndarray, ivectors: Optional[np.ndarray] = None) -> float:
        net = self.net
        net.ZeroGrads()
        net.Forward(features, ivectors)
        net.Backward(None)
        loss = net.ReadLoss()
        net.SGDStep()
        return loss

    def SetLR(self, lr: float) -> None:
        self.net.SetLR(lr)
