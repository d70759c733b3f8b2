"""Micro-timing of the HBM-bound helper kernels on the TDNN-F shapes (back-to-back launches, CUDA events).
usage: python scripts/elt_exp.py [rot=1|4]   rot = number of distinct buffer sets cycled through (4 > L2)"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from kaldi_fp16_b200 import _lib, cudart, gpu  # noqa: E402

kv = dict(a.split("=") for a in sys.argv[1:])
rot = int(kv.get("rot", 1))
lib = _lib.load()
gpu.Init(0)
h = gpu.NewHandle()
n_seq, L, halo, D = 64, 150, 3, 1536
rows = n_seq * (L + 2 * halo)
mask_ld = D // 32
sets = []
for _ in range(rot):
    sets.append((gpu.ZeroTensor(rows, D), gpu.ZeroTensor(rows, D), gpu.DeviceF32(n=rows * mask_ld)))
scale, db = gpu.DeviceF32(np.ones(D, np.float32)), gpu.DeviceF32(n=D)


def timeit(name, fn, bytes_moved, iters=40):
    for i in range(5):
        fn(sets[i % rot])
    e0, e1 = cudart.Event(), cudart.Event()
    e0.record()
    for i in range(iters):
        fn(sets[i % rot])
    e1.record()
    e1.synchronize()
    us = e0.elapsed_ms(e1) * 1e3 / iters
    print(f"{name:34s} rot={rot}  {us:7.2f} us  {bytes_moved / us / 1e6:7.2f} TB/s")


def bn_fold(s):
    assert lib.kfp16_bn_relu_backward_bias_fold(h.ptr, s[0].Ptr, D, scale.Ptr, s[2].Ptr, mask_ld, s[1].Ptr, D, n_seq, L, halo, D, db.Ptr) == 0


def bn_plain(s):
    assert lib.kfp16_bn_relu_backward_bias(h.ptr, s[0].Ptr, D, scale.Ptr, s[2].Ptr, mask_ld, s[1].Ptr, D, rows, D, db.Ptr) == 0


def bn_nodb(s):
    assert lib.kfp16_bn_relu_backward_bias(h.ptr, s[0].Ptr, D, scale.Ptr, s[2].Ptr, mask_ld, s[1].Ptr, D, rows, D, None) == 0


def bn_simple(s):
    assert lib.kfp16_bn_relu_backward(h.ptr, s[0].Ptr, D, scale.Ptr, s[2].Ptr, mask_ld, s[1].Ptr, D, rows, D) == 0


def copy(s):
    assert lib.ops_copy(s[1].Ptr, s[0].Ptr, rows * D) == 0


def fold(s):
    assert lib.kfp16_fold_edges(h.ptr, s[0].Ptr, D, n_seq, L, D, halo) == 0


def pad(s):
    assert lib.kfp16_pad_edges(h.ptr, s[0].Ptr, D, n_seq, L, D, halo) == 0


B = rows * D * 2
timeit("ops_copy", copy, 2 * B)
timeit("bn_relu_backward (no colsum)", bn_simple, 2 * B + rows * mask_ld * 4)
timeit("bn_relu_backward_bias db=NULL", bn_nodb, 2 * B + rows * mask_ld * 4)
timeit("bn_relu_backward_bias", bn_plain, 2 * B + rows * mask_ld * 4)
timeit("bn_relu_backward_bias_fold", bn_fold, 2 * B + rows * mask_ld * 4)
timeit("fold_edges", fold, n_seq * 2 * (halo + 1) * D * 2 * 2)
timeit("pad_edges", pad, n_seq * 2 * (halo + 1) * D * 2)
