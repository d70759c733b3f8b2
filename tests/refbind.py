"""ctypes binding of oracle/_ref/libkaldi_fp16_ref.so -- the reference's native operator library
compiled UNMODIFIED from /root/reference/cpp by oracle/Makefile.  Test infrastructure: it gives the
exact cublasGemmEx / elementwise results of the reference on the GPU box and is used to pin the
numpy oracle and to cross-check the new kernels.  Never imported by the product package."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

REF_PATH = Path(__file__).resolve().parent.parent / "oracle" / "_ref" / "libkaldi_fp16_ref.so"
_ref = None


def load_ref():
    global _ref
    if _ref is not None:
        return _ref
    if not REF_PATH.exists():
        return None
    lib = C.CDLL(str(REF_PATH))
    vp, ci, cf = C.c_void_p, C.c_int, C.c_float
    sig = {
        "ops_cublas_create": (vp, []), "ops_cublas_destroy": (None, [vp]),
        "ops_gemm": (ci, [vp, ci, ci, ci, cf, vp, ci, vp, ci, cf, vp, ci]),
        "ops_relu": (ci, [vp, ci]), "ops_sigmoid": (ci, [vp, ci]), "ops_tanh_act": (ci, [vp, ci]),
        "ops_clipped_relu": (ci, [vp, ci, cf]), "ops_softmax": (ci, [vp, ci, ci]), "ops_log_softmax": (ci, [vp, ci, ci]),
        "ops_batchnorm_forward": (ci, [vp, ci, ci, vp, vp, vp, vp, cf]),
        "ops_batchnorm_forward_rms": (ci, [vp, ci, ci, vp, vp, cf, cf]),
        "ops_add_scaled": (ci, [vp, vp, ci, cf, cf]), "ops_add": (ci, [vp, vp, ci]), "ops_copy": (ci, [vp, vp, ci]),
        "ops_fill": (ci, [vp, ci, cf]), "ops_concat_cols": (ci, [vp, ci, ci, vp, ci, ci]),
        "ops_slice_cols": (ci, [vp, ci, ci, vp, ci, ci]), "ops_combine_feature_maps": (ci, [vp, ci, ci, ci, ci, ci]),
        "ops_subsample_rows": (None, [vp, vp, ci, ci, ci, ci]),
        "ops_relu_backward": (ci, [vp, vp, ci]), "ops_sigmoid_backward": (ci, [vp, vp, ci]),
        "ops_tanh_backward": (ci, [vp, vp, ci]), "ops_transpose": (ci, [vp, vp, ci, ci]),
        "ops_batchnorm_backward": (ci, [vp, vp, vp, vp, cf, ci, ci]), "ops_fp16_to_fp32": (ci, [vp, vp, ci]),
        "ops_sgd_update": (ci, [vp, vp, vp, vp, cf, cf, ci]),
        "bridge_gpu_malloc": (vp, [C.c_size_t]), "bridge_gpu_free": (None, [vp]), "bridge_gpu_sync": (ci, []),
        "bridge_transfer_fp16": (ci, [vp, vp, C.c_size_t]), "bridge_read_fp16": (ci, [vp, vp, C.c_size_t]),
        "bridge_transfer_float32": (ci, [vp, vp, C.c_size_t]),
        "kaldi_cublas_create": (vp, []), "kaldi_cublas_destroy": (None, [vp]),
        "kaldi_tensor_create": (vp, [ci, ci]), "kaldi_tensor_free": (None, [vp]),
        "kaldi_tensor_copy_from_host_fp32": (None, [vp, vp, C.c_size_t]),
        "kaldi_tensor_copy_to_host_fp32": (None, [vp, vp, C.c_size_t]),
        "kaldi_gemm": (None, [vp, vp, vp, vp, cf, cf, ci, ci]),
        "launch_conv1d_forward_fp16": (None, [vp, vp, vp, vp, ci, ci, ci, ci, ci, ci, ci, ci, vp]),
        "launch_conv1d_backward_fp16": (None, [vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, ci, ci, ci, vp]),
        "launch_pointwise_conv1d_fp16": (None, [vp, vp, vp, vp, ci, ci, ci, ci, vp]),
        "launch_batchnorm1d_forward_fp16": (None, [vp, vp, vp, vp, vp, vp, vp, vp, ci, ci, ci, cf, cf, C.c_bool, vp]),
    }
    sig.update({
        "bridge_transfer_int32": (ci, [vp, vp, C.c_size_t]),
        "chain_compute_loss": (ci, [vp, vp, vp, ci, ci, vp, vp]),
        "chain_forward_backward": (ci, [vp, vp, ci, ci, vp, vp, C.POINTER(C.c_float)]),
        "chain_compute_posteriors": (ci, [vp, vp, ci, ci, vp, vp, cf, vp]),
        "chain_workspace_bytes": (C.c_size_t, [ci, ci]),
        "chain_last_error": (C.c_char_p, []),
    })
    for name, (res, args) in sig.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:          # an older build of the reference library without the chain objects
            continue
        fn.restype, fn.argtypes = res, args
    _ref = lib
    return lib


class RefBuf:
    """Device buffer owned by the reference library's allocator."""

    def __init__(self, lib, host: np.ndarray):
        self.lib = lib
        self.host_dtype = host.dtype
        self.shape = host.shape
        self.nbytes = host.nbytes
        self.ptr = lib.bridge_gpu_malloc(max(host.nbytes, 16))
        assert self.ptr, "reference bridge_gpu_malloc failed"
        hc = np.ascontiguousarray(host)
        if host.dtype == np.uint16:
            assert lib.bridge_transfer_fp16(self.ptr, hc.ctypes.data, hc.size) == 0
        elif host.dtype == np.float32:
            assert lib.bridge_transfer_float32(self.ptr, hc.ctypes.data, hc.size) == 0
        else:
            raise TypeError(host.dtype)

    def bits(self) -> np.ndarray:
        out = np.empty(self.shape, dtype=np.uint16)
        assert self.lib.bridge_read_fp16(out.ctypes.data, self.ptr, out.size) == 0
        return out

    def f32(self) -> np.ndarray:
        return self.bits().view(np.float16).astype(np.float32)

    def free(self):
        if self.ptr:
            self.lib.bridge_gpu_free(self.ptr)
            self.ptr = None


def ref_half(lib, x_f32: np.ndarray) -> RefBuf:
    """upload fp16-representable float32 values as fp16"""
    return RefBuf(lib, np.ascontiguousarray(x_f32, dtype=np.float32).astype(np.float16).view(np.uint16))


class RefChainFst(C.Structure):
    """struct ChainFstGPU (reference cpp/include/chain.h:24-36): CSR with DEVICE pointers"""
    _fields_ = [("row_ptr", C.c_void_p), ("col_idx", C.c_void_p), ("labels", C.c_void_p), ("weights", C.c_void_p),
                ("final_states", C.c_void_p), ("final_weights", C.c_void_p),
                ("num_states", C.c_int), ("num_arcs", C.c_int), ("num_final", C.c_int), ("start_state", C.c_int)]


class RefChainResult(C.Structure):
    _fields_ = [("num_logprob", C.c_float), ("den_logprob", C.c_float), ("loss", C.c_float)]


def ref_chain_fst(lib, fst) -> tuple:
    """upload an oracle.chain_oracle.Fst through the reference's own bridge; returns (struct, buffers to free)"""
    bufs = []

    def up_i32(a):
        a = np.ascontiguousarray(a, np.int32)
        p = lib.bridge_gpu_malloc(max(a.nbytes, 16))
        assert p and (a.size == 0 or lib.bridge_transfer_int32(p, a.ctypes.data, a.size) == 0)
        bufs.append(p)
        return p

    def up_f32(a):
        a = np.ascontiguousarray(a, np.float32)
        p = lib.bridge_gpu_malloc(max(a.nbytes, 16))
        assert p and (a.size == 0 or lib.bridge_transfer_float32(p, a.ctypes.data, a.size) == 0)
        bufs.append(p)
        return p

    f = RefChainFst()
    f.row_ptr, f.col_idx, f.labels, f.weights = up_i32(fst.row_ptr), up_i32(fst.col_idx), up_i32(fst.labels), up_f32(fst.weights)
    f.final_states, f.final_weights = up_i32(fst.final_states), up_f32(fst.final_weights)
    f.num_states, f.num_arcs, f.num_final, f.start_state = fst.num_states, fst.num_arcs, len(fst.final_states), fst.start_state
    return f, bufs


def read_f32(ptr: int, shape) -> np.ndarray:
    """device fp32 buffer -> host (plain cudaMemcpy: the pointer may come from either library's allocator)"""
    from kaldi_fp16_b200 import cudart
    out = np.empty(shape, np.float32)
    cudart.synchronize()
    cudart.check(cudart.rt().cudaMemcpy(C.c_void_p(out.ctypes.data), C.c_void_p(ptr), C.c_size_t(out.nbytes), C.c_int(2)), "cudaMemcpy D2H")
    return out
