"""FP16 GEMM shape sweep (BASELINE.json configs[4]): new tcgen05 kernel vs the
reference's ops_gemm (cuBLAS, oracle/_ref) on the same B200, CUDA-event timed, L2 flushed between
iterations.  Writes gpurun_out/gemm_sweep.json.   usage: python scripts/gemm_sweep.py [--iters N]"""
import argparse
import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from kaldi_fp16_b200 import _lib, cudart, gpu  # noqa: E402

PEAKS = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
PEAK = PEAKS.get("bf16_tflops", 1590.0)

# name, M, K, N, variants (NN fwd / NT dgrad / TN wgrad computed from the same layer)
LAYERS = [
    ("cnn2", 384000, 576, 64), ("cnn3", 192000, 576, 128), ("cnn4", 192000, 1152, 128), ("cnn5", 96000, 1152, 256),
    ("cnn6", 96000, 2304, 256), ("tdnnf7.linear", 9600, 2560, 256), ("tdnnf.linear", 9600, 3072, 160),
    ("tdnnf.affine", 9600, 320, 1536), ("prefinal.affine", 9600, 256, 1536), ("prefinal.linear", 9600, 1536, 256),
    ("output", 9600, 256, 6016), ("sq4096", 4096, 4096, 4096), ("sq8192", 8192, 8192, 8192),
]


def time_fn(fn, iters, flush):
    e0, e1 = cudart.Event(), cudart.Event()
    for _ in range(3):
        fn()
    cudart.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            cudart.memset(flush.Ptr, 0, flush.N * 4)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_ms(e1))
    return float(np.median(ts)), float(np.min(ts))


def fill_random(lib, t, noise):
    """tile a block of N(0, 0.05) fp16 noise over the tensor (constant data would under-state power)"""
    n, off = t.Numel(), 0
    while off < n:
        c = min(noise.Numel(), n - off)
        assert lib.ops_copy(t.Ptr + off * 2, noise.Ptr, c) == 0
        off += c


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--no-ref", action="store_true")
    args = ap.parse_args()
    lib = _lib.load()
    gpu.Init(0)
    h = gpu.NewHandle()
    ref = None
    if not args.no_ref:
        sys.path.insert(0, str(ROOT))
        from tests.refbind import load_ref
        ref = load_ref()
    rh = ref.ops_cublas_create() if ref else None
    flush = gpu.DeviceF32(n=64 * 1024 * 1024)   # 256 MB > 126 MB L2
    rng = np.random.default_rng(0)
    noise = gpu.TensorFromFP16((rng.standard_normal((4096, 2048)) * 0.05).astype(np.float16))
    rows = []
    for name, M, K, N in LAYERS:
        variants = [("fwd NN", M, N, K, 0, 0), ("dgrad NT", M, K, N, 0, 1), ("wgrad TN", K, N, M, 1, 0)]
        for vname, m, n, k, ta, tb in variants:
            a_shape = (k, m) if ta else (m, k)
            b_shape = (n, k) if tb else (k, n)
            A = gpu.NewTensor(*a_shape)
            B = gpu.NewTensor(*b_shape)
            Cc = gpu.NewTensor(m, n)
            fill_random(lib, A, noise)
            fill_random(lib, B, noise)
            flops = 2.0 * m * n * k

            def ours():
                rc = lib.kfp16_gemm(h.ptr, m, n, k, 1.0, A.Ptr, ta, B.Ptr, tb, 0.0, Cc.Ptr)
                assert rc == 0, _lib.last_error()

            med, best = time_fn(ours, args.iters, flush)
            row = {"layer": name, "variant": vname, "M": m, "N": n, "K": k, "ms": med, "ms_best": best,
                   "tflops": flops / med / 1e9, "frac_of_measured_peak": flops / med / 1e9 / PEAK}
            if ref is not None and not ta and not tb:
                def theirs():
                    assert ref.ops_gemm(rh, m, n, k, 1.0, A.Ptr, k, B.Ptr, n, 0.0, Cc.Ptr, n) == 0
                rmed, rbest = time_fn(theirs, args.iters, flush)
                row.update({"ref_cublas_ms": rmed, "ref_cublas_tflops": flops / rmed / 1e9, "speedup_vs_ref": rmed / med})
            rows.append(row)
            print(json.dumps(row), flush=True)
            for t in (A, B, Cc):
                t.Free()
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    (out / "gemm_sweep.json").write_text(json.dumps({"peak_tflops": PEAK, "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
