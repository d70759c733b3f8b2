"""Generate tests/golden/ref_ops.npz: seeded inputs and the outputs of the REFERENCE's own operator
library (oracle/_ref/libkaldi_fp16_ref.so = /root/reference/cpp/{cuda,src} compiled unmodified by
oracle/Makefile, cuBLAS-backed) for every operator on the hot path.  Runs on the GPU box:

    gpurun -- python scripts/gen_ref_golden.py gpurun_out/ref_ops.npz      # then copy to tests/golden/

The fixtures pin the numpy oracle (tests/test_oracle_cpu.py) and, through it, the CUDA kernels.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from tests.refbind import RefBuf, load_ref, ref_half  # noqa: E402


def f16(rng, shape, scale=1.0):
    return (rng.standard_normal(shape).astype(np.float32) * np.float32(scale)).astype(np.float16)


def main(out_path):
    lib = load_ref()
    assert lib is not None, "oracle/_ref/libkaldi_fp16_ref.so missing (make -C oracle ref)"
    h = lib.ops_cublas_create()
    assert h
    rng = np.random.default_rng(20260210)
    out = {}

    def up(x16):
        return RefBuf(lib, np.ascontiguousarray(x16).view(np.uint16))

    def upf(x32):
        return RefBuf(lib, np.ascontiguousarray(x32, dtype=np.float32))

    # ---- ops_gemm (cublasGemmEx fp16 in/out, fp32 accumulate), TDNN-F / CNN shaped and ragged
    for i, (M, N, K, alpha, beta) in enumerate([(96, 160, 192, 1.0, 0.0), (130, 72, 88, 1.0, 0.0), (64, 256, 320, 0.5, 1.0),
                                                (33, 40, 40, 1.0, 0.0), (150, 136, 256, 1.0, 0.0), (200, 64, 54, 1.0, 0.0)]):
        A, B, C0 = f16(rng, (M, K)), f16(rng, (K, N), 0.1), f16(rng, (M, N))
        dA, dB, dC = up(A), up(B), up(C0)
        assert lib.ops_gemm(h, M, N, K, alpha, dA.ptr, K, dB.ptr, N, beta, dC.ptr, N) == 0
        lib.bridge_gpu_sync()
        out[f"gemm{i}_A"], out[f"gemm{i}_B"], out[f"gemm{i}_C0"] = A.view(np.uint16), B.view(np.uint16), C0.view(np.uint16)
        out[f"gemm{i}_ab"] = np.array([alpha, beta], np.float32)
        out[f"gemm{i}_out"] = dC.bits()
        for b in (dA, dB, dC):
            b.free()

    # ---- elementwise forward
    T, D = 37, 48
    x = f16(rng, (T, D), 2.0)
    x[0, :4] = [np.float16(-0.0), np.float16(0.0), np.float16("nan"), np.float16(-1e-4)]
    out["x"] = x.view(np.uint16)
    for name, fn, args in [("relu", lib.ops_relu, ()), ("sigmoid", lib.ops_sigmoid, ()), ("tanh", lib.ops_tanh_act, ()),
                           ("clipped_relu", lib.ops_clipped_relu, (1.5,))]:
        d = up(x)
        assert fn(d.ptr, x.size, *args) == 0
        lib.bridge_gpu_sync()
        out[name] = d.bits()
        d.free()
    xs = f16(rng, (T, D), 2.0)
    xs += np.float16(3.0)           # positive row maxima: the reference's atomicMax trick is only right there
    out["xs"] = xs.view(np.uint16)
    for name, fn in [("softmax", lib.ops_softmax), ("log_softmax", lib.ops_log_softmax)]:
        d = up(xs)
        assert fn(d.ptr, T, D) == 0
        lib.bridge_gpu_sync()
        out[name] = d.bits()
        d.free()
    mean, var = (rng.standard_normal(D) * 0.3).astype(np.float32), (rng.random(D) + 0.5).astype(np.float32)
    gamma, beta = (rng.random(D) + 0.5).astype(np.float32), (rng.standard_normal(D) * 0.2).astype(np.float32)
    out["bn_mean"], out["bn_var"], out["bn_gamma"], out["bn_beta"] = mean, var, gamma, beta
    dm, dv, dg, dbt = upf(mean), upf(var), upf(gamma), upf(beta)
    d = up(x)
    assert lib.ops_batchnorm_forward(d.ptr, T, D, dm.ptr, dv.ptr, dg.ptr, dbt.ptr, 1e-3) == 0
    lib.bridge_gpu_sync()
    out["bn_fwd"] = d.bits()
    d.free()
    d = up(x)
    assert lib.ops_batchnorm_forward_rms(d.ptr, T, D, dm.ptr, dv.ptr, 0.025, 1e-3) == 0
    lib.bridge_gpu_sync()
    out["bn_rms"] = d.bits()
    d.free()
    y = f16(rng, (T, D))
    out["y"] = y.view(np.uint16)
    d, s = up(x), up(y)
    assert lib.ops_add_scaled(d.ptr, s.ptr, x.size, 0.66, 1.0) == 0
    lib.bridge_gpu_sync()
    out["add_scaled"] = d.bits()
    d.free()
    d = up(x)
    assert lib.ops_add(d.ptr, s.ptr, x.size) == 0
    lib.bridge_gpu_sync()
    out["add"] = d.bits()
    d.free(); s.free()

    # ---- data movement
    H, F1, F2 = 8, 1, 5
    xc = f16(rng, (T, H * (F1 + F2)))
    out["combine_x"] = xc.view(np.uint16)
    d = up(xc)
    assert lib.ops_combine_feature_maps(d.ptr, T, H * (F1 + F2), H, F1, F2) == 0
    lib.bridge_gpu_sync()
    out["combine"] = d.bits()
    d.free()
    src, dst = up(x), RefBuf(lib, np.zeros((13, D), np.uint16))
    lib.ops_subsample_rows(dst.ptr, src.ptr, T, D, 3, 1)
    lib.bridge_gpu_sync()
    out["subsample_3_1"] = dst.bits()[: len(range(1, T, 3))]
    src.free(); dst.free()
    src, dst = up(x), RefBuf(lib, np.zeros((D, T), np.uint16))
    assert lib.ops_transpose(src.ptr, dst.ptr, T, D) == 0
    lib.bridge_gpu_sync()
    out["transpose"] = dst.bits()
    src.free(); dst.free()

    # ---- backward ops
    g = f16(rng, (T, D))
    out["g"] = g.view(np.uint16)
    for name, fn in [("relu_backward", lib.ops_relu_backward), ("sigmoid_backward", lib.ops_sigmoid_backward),
                     ("tanh_backward", lib.ops_tanh_backward)]:
        act = x if name == "relu_backward" else np.tanh(x.astype(np.float32) * 0.5).astype(np.float16)
        if name == "sigmoid_backward":
            act = (1 / (1 + np.exp(-x.astype(np.float32)))).astype(np.float16)
        out[name + "_act"] = np.ascontiguousarray(act).view(np.uint16)
        da, dg2 = up(act), up(g)
        assert fn(da.ptr, dg2.ptr, g.size) == 0
        lib.bridge_gpu_sync()
        out[name] = dg2.bits()
        da.free(); dg2.free()
    dgo, dgi = up(g), RefBuf(lib, np.zeros((T, D), np.uint16))
    assert lib.ops_batchnorm_backward(dgo.ptr, dgi.ptr, dg.ptr, dv.ptr, 1e-5, T, D) == 0
    lib.bridge_gpu_sync()
    out["bn_bwd"] = dgi.bits()
    dgo.free(); dgi.free()

    # ---- SGD with FP32 masters (cmd/sgdtest tests 1-3): 3 steps
    n = 1000
    w16 = f16(rng, (n,))
    w32 = w16.astype(np.float32)
    vel = np.zeros(n, np.float32)
    out["sgd_w0"] = w16.view(np.uint16)
    dw32, dw16, dvel = upf(w32), up(w16), upf(vel)
    grads = []
    for step in range(3):
        gr = f16(rng, (n,), 0.5)
        grads.append(gr.view(np.uint16))
        dgr = up(gr)
        assert lib.ops_sgd_update(dw32.ptr, dw16.ptr, dgr.ptr, dvel.ptr, 0.01, 0.9, n) == 0
        lib.bridge_gpu_sync()
        dgr.free()
    out["sgd_grads"] = np.stack(grads)
    out["sgd_w16"] = dw16.bits()
    tmp = np.empty(n, np.float32)
    import ctypes as C
    cudart = C.CDLL("libcudart.so")
    cudart.cudaMemcpy(C.c_void_p(tmp.ctypes.data), C.c_void_p(dw32.ptr), C.c_size_t(n * 4), C.c_int(2))
    out["sgd_w32"] = tmp.copy()
    cudart.cudaMemcpy(C.c_void_p(tmp.ctypes.data), C.c_void_p(dvel.ptr), C.c_size_t(n * 4), C.c_int(2))
    out["sgd_vel"] = tmp.copy()

    # ---- kaldi_gemm (cublasHgemm: fp16 accumulate type) second operator surface
    kh = lib.kaldi_cublas_create()
    M, N, K = 48, 40, 64
    A32, B32 = f16(rng, (M, K)).astype(np.float32), f16(rng, (K, N), 0.1).astype(np.float32)
    tA, tB, tC = lib.kaldi_tensor_create(M, K), lib.kaldi_tensor_create(K, N), lib.kaldi_tensor_create(M, N)
    lib.kaldi_tensor_copy_from_host_fp32(tA, A32.ctypes.data, A32.size)
    lib.kaldi_tensor_copy_from_host_fp32(tB, B32.ctypes.data, B32.size)
    lib.kaldi_gemm(kh, tA, tB, tC, 1.0, 0.0, 0, 0)
    lib.bridge_gpu_sync()
    Cout = np.empty((M, N), np.float32)
    lib.kaldi_tensor_copy_to_host_fp32(tC, Cout.ctypes.data, Cout.size)
    out["kgemm_A"], out["kgemm_B"], out["kgemm_out"] = A32, B32, Cout

    np.savez_compressed(out_path, **out)
    print("wrote", out_path, "with", len(out), "arrays")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else str(ROOT / "gpurun_out" / "ref_ops.npz"))
