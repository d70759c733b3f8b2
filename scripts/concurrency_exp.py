"""Do a 78-CTA dgrad GEMM and a 144-CTA wgrad GEMM overlap when launched on two streams?"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from kaldi_fp16_b200 import _lib, cudart, gpu  # noqa: E402
from kaldi_fp16_b200._lib import GemmDesc, K_MAJOR, MN_MAJOR  # noqa: E402

lib = _lib.load()
gpu.Init(0)
hA, hB = gpu.NewHandle(), gpu.NewHandle()
sA, sB = cudart.Stream(), cudart.Stream()
lib.kfp16_ctx_set_stream(hA.ptr, sA.ptr)
lib.kfp16_ctx_set_stream(hB.ptr, sB.ptr)
T, H, Bt = 9984, 1536, 160
X, dZ, Bo = gpu.NewTensor(T, H), gpu.NewTensor(T, H), gpu.NewTensor(T, Bt)
Wl = gpu.NewTensor(2 * H, Bt)
ws = gpu.DeviceF32(n=2 * H * Bt)
d1 = GemmDesc()   # narrow dgrad-like GEMM: [T x 3072] * [3072 x 160] (78 CTAs)
d1.M, d1.N, d1.K, d1.groups, d1.kslabs, d1.kslab_len, d1.alpha = T, Bt, H, 1, 1, H, 1.0
d1.a_major, d1.b_major = K_MAJOR, MN_MAJOR
d1.A.ptr, d1.A.rows, d1.A.cols, d1.A.ld = X.Ptr, T, H, H
d1.B.ptr, d1.B.rows, d1.B.cols, d1.B.ld = Wl.Ptr, H, Bt, Bt
d1.D[0], d1.ldd = Bo.Ptr, Bt
d2 = GemmDesc()   # wgrad: [1536 x T] * [T x 160], split-K
d2.M, d2.N, d2.K, d2.groups, d2.kslabs, d2.kslab_len, d2.alpha = H, Bt, T, 1, 1, T, 1.0
d2.a_major, d2.b_major = MN_MAJOR, MN_MAJOR
d2.A.ptr, d2.A.rows, d2.A.cols, d2.A.ld = dZ.Ptr, T, H, H
d2.B.ptr, d2.B.rows, d2.B.cols, d2.B.ld = Bo.Ptr, T, Bt, Bt
d2.split_k, d2.ws_ld = 12, Bt
d2.ws[0] = ws.Ptr


def run(n, both, serial=False):
    cudart.synchronize()
    e0, e1 = cudart.Event(), cudart.Event()
    e0.record(sA.ptr)
    for _ in range(n):
        assert lib.kfp16_gemm_ex(hA.ptr, C.byref(d1)) == 0
        if both:
            assert lib.kfp16_gemm_ex(hA.ptr if serial else hB.ptr, C.byref(d2)) == 0
    cudart.synchronize()
    e1.record(sA.ptr)
    e1.synchronize()
    return e0.elapsed_ms(e1) * 1e3 / n


for _ in range(2):
    print("narrow only      %.1f us/iter" % run(200, False))
    print("narrow + wgrad, same stream   %.1f us/iter" % run(200, True, True))
    print("narrow + wgrad, two streams   %.1f us/iter" % run(200, True, False))
