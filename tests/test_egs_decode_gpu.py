"""egs feature decode on the device (kfp16_decode_matrices / kfp16_net_set_input_compressed): Kaldi compressed matrices
(CM / CM2 / CM3) and full matrices as they sit in the archive -> FP16 rows, BIT-IDENTICAL to the reference's CPU path:
parser.ReadCompressedMatrix* (internal/parser/matrix.go:11-165) followed by the RNE FP32 -> FP16 conversion of the
features (internal/gpu/bridge.go:141, internal/fp16/fp16.go:13-70)."""
import ctypes as C

import numpy as np
import pytest

from kaldi_fp16_b200 import _lib, gpu, nnet
from oracle import kaldi_oracle as O

pytestmark = pytest.mark.gpu


def make_matrix(rng, fmt, rows, cols):
    """random payload in the archive's layout + the oracle's decode of it"""
    gmin, grange = np.float32(rng.uniform(-40, 5)), np.float32(rng.uniform(20, 140))
    if fmt == "CM":
        hdr = np.sort(rng.integers(0, 65536, size=(cols, 4)).astype(np.uint16), axis=1)
        data = rng.integers(0, 256, size=(cols, rows)).astype(np.uint8)
        data.reshape(-1)[:6] = [0, 64, 65, 192, 193, 255]
        payload = hdr.tobytes() + data.tobytes()
        want = O.decode_cm(payload, rows, cols, gmin, grange)
    elif fmt == "CM2":
        payload = rng.integers(0, 65536, size=rows * cols).astype(np.uint16).tobytes()
        want = O.decode_cm2(payload, rows, cols, gmin, grange)
    elif fmt == "CM3":
        payload = rng.integers(0, 256, size=rows * cols).astype(np.uint8).tobytes()
        want = O.decode_cm3(payload, rows, cols, gmin, grange)
    else:
        x = (rng.standard_normal((rows, cols)) * 30).astype("<f4")
        x.reshape(-1)[:3] = [65504.0, 65520.0, 1e-8]
        payload = x.tobytes()
        want = O.decode_fm(payload, rows, cols)
    return payload, float(gmin), float(grange), want


@pytest.mark.parametrize("fmts", [["CM"] * 3, ["CM2", "CM3", "FM", "CM"], ["CM"] * 70])
def test_decode_matrices_bit_exact(lib, handle, fmts):
    rng = np.random.default_rng(len(fmts))
    rows, cols, ld = 37, 40, 48
    code = {"CM": 1, "CM2": 2, "CM3": 3, "FM": 4}
    blob, descs, wants = bytearray(), (_lib.CmDesc * len(fmts))(), []
    for i, f in enumerate(fmts):
        payload, gmin, grange, want = make_matrix(rng, f, rows, cols)
        if len(blob) & 1:
            blob.append(0)
        d = descs[i]
        d.format, d.rows, d.cols, d.global_min, d.global_range = code[f], rows, cols, gmin, grange
        d.payload_offset, d.dst_row = len(blob), i * (rows + 2) + 1
        assert lib.kfp16_cm_payload_bytes(C.byref(d)) == len(payload)
        blob += payload
        wants.append(want)
    if len(blob) & 1:
        blob.append(0)
    dev = lib.bridge_gpu_malloc(len(blob))
    host = (C.c_ubyte * len(blob)).from_buffer(blob)
    assert lib.bridge_transfer_fp16(dev, host, len(blob) // 2) == 0          # (a byte copy: 2 bytes per "fp16" element)
    total_rows = len(fmts) * (rows + 2)
    dst = gpu.TensorFromFP16(np.full((total_rows, ld), 7.0, np.float32))
    assert lib.kfp16_decode_matrices(handle.ptr, dev, len(blob), descs, len(fmts), dst.Ptr, ld, total_rows) == 0, _lib.last_error()
    gpu.Sync()
    got = dst.ToBits().reshape(total_rows, ld)
    for i, want in enumerate(wants):
        r0 = i * (rows + 2) + 1
        assert np.array_equal(got[r0:r0 + rows, :cols], O.fp16_from_float32_rne(want)), (i, fmts[i])
        assert np.all(got[r0 - 1] == 0x4700) and np.all(got[r0:r0 + rows, cols:] == 0x4700)     # untouched (7.0)
    # malformed descriptors fail loudly
    descs[0].payload_offset = len(blob)
    assert lib.kfp16_decode_matrices(handle.ptr, dev, len(blob), descs, 1, dst.Ptr, ld, total_rows) == -1
    lib.bridge_gpu_free(dev)
    dst.Free()


def test_network_input_from_compressed_egs(lib, handle):
    """kfp16_net_set_input_compressed == decoding on the host (oracle) + kfp16_net_set_input: same padded input, same output"""
    xconfig = """
input name=input dim=40
linear-component name=lin0 dim=64
tdnnf-layer name=tdnnf1 dim=64 bottleneck-dim=32 time-stride=3 bypass-scale=0.66
output-layer name=output include-log-softmax=false dim=24
"""
    n_seq, L = 5, 33
    rng = np.random.default_rng(8)
    net = nnet.NewNetwork(nnet.BuildModelFromString(xconfig), handle, n_seq, L, train=False)
    mats, dense = [], []
    for q, f in enumerate(["CM", "CM2", "CM", "CM3", "FM"]):
        payload, gmin, grange, want = make_matrix(rng, f, L, 40)
        if f == "FM":
            want = np.clip(want, -60000, 60000)
            payload = want.astype("<f4").tobytes()
        mats.append((f, payload, gmin, grange))
        dense.append(want)
    x = np.concatenate(dense, 0)
    want_out = net.Forward(x)                       # host decode (oracle) + RNE on the host
    want_in = net.Output("input")
    net.SetInput("input", np.zeros_like(x))
    net.SetInputCompressed("input", mats)           # payload bytes -> device decode
    assert lib.kfp16_net_forward(net.ptr) == 0
    assert np.array_equal(net.Output("input"), want_in)
    assert np.array_equal(net.Output(""), want_out)
    net.Free()
