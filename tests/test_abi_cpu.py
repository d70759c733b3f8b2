"""CPU tests of the drop-in boundary: the C-ABI shared library builds for sm_100a, loads without a
GPU, exports every symbol include/*.h declares (and every reference symbol SURVEY 8b lists for the
path), the ctypes binding table matches the headers, and entry points fail loudly -- never fall
back to a CPU path -- when there is no CUDA device."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
INC = ROOT / "include"


def declared_functions() -> dict:
    """name -> header, for every function prototype in include/*.h"""
    out = {}
    for hdr in sorted(INC.glob("*.h")):
        text = re.sub(r"/\*.*?\*/", "", hdr.read_text(), flags=re.S)
        text = re.sub(r"//[^\n]*", "", text)
        text = re.sub(r"typedef\s+struct\s*\{.*?\}\s*\w+\s*;", "", text, flags=re.S)
        text = re.sub(r"enum\s*\{.*?\}\s*;", "", text, flags=re.S)
        for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b((?:kfp16|ops|bridge|kaldi|launch|chain)_\w+)\s*\(", text):
            out[m.group(2)] = hdr.name
    return out


@pytest.fixture(scope="module")
def built_lib():
    from kaldi_fp16_b200 import build

    return build.build(verbose=False)


def test_headers_declare_the_reference_surface():
    fns = declared_functions()
    # the symbols internal/gpu binds (cpp/include/ops.h:16-188, bridge.h:12-60)
    for name in ["ops_cublas_create", "ops_cublas_destroy", "ops_gemm", "ops_gemm_strided", "ops_relu", "ops_sigmoid",
                 "ops_tanh_act", "ops_clipped_relu", "ops_softmax", "ops_log_softmax", "ops_batchnorm_forward",
                 "ops_batchnorm_forward_rms", "ops_add_scaled", "ops_add", "ops_copy", "ops_fill", "ops_concat_cols",
                 "ops_slice_cols", "ops_combine_feature_maps", "ops_subsample_rows", "ops_relu_backward",
                 "ops_sigmoid_backward", "ops_tanh_backward", "ops_transpose", "ops_batchnorm_backward", "ops_fp16_to_fp32",
                 "ops_sgd_update", "ops_last_error", "ops_clear_error", "bridge_gpu_init", "bridge_gpu_get_free_memory",
                 "bridge_gpu_sync", "bridge_gpu_malloc", "bridge_gpu_free", "bridge_host_alloc", "bridge_host_free",
                 "bridge_transfer_fp16", "bridge_transfer_int32", "bridge_transfer_float32", "bridge_read_fp16",
                 "bridge_batch_alloc", "bridge_batch_transfer", "bridge_batch_free", "bridge_gpu_memset",
                 "bridge_fp16_to_fp32_gpu", "bridge_fp32_to_fp16_gpu", "bridge_last_error", "bridge_clear_error"]:
        assert name in fns, f"{name} (reference C ABI) is not declared in include/*.h"


def test_library_exports_every_declared_symbol(built_lib):
    fns = declared_functions()
    assert len(fns) > 100
    nm = subprocess.run(["nm", "-D", "--defined-only", str(built_lib)], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in nm.splitlines() if " T " in line}
    missing = sorted(n for n in fns if n not in exported)
    assert not missing, f"declared in include/*.h but not exported by {built_lib.name}: {missing}"


def test_binding_table_matches_headers(built_lib):
    from kaldi_fp16_b200 import _lib

    fns = declared_functions()
    assert sorted(set(fns) - set(_lib.SIGNATURES)) == [], "header functions without a ctypes signature"
    assert sorted(set(_lib.SIGNATURES) - set(fns)) == [], "ctypes signatures without a header declaration"
    _lib.load()     # resolves every symbol; raises ImportError otherwise


def test_struct_layouts_match_the_headers(built_lib):
    """sizeof() of the ctypes mirrors == sizeof() the C compiler sees (a probe compiled with gcc)"""
    from kaldi_fp16_b200 import _lib

    src = ('#include <stdio.h>\n#include "kaldi_fp16_nnet.h"\n#include "kaldi_fp16_bridge.h"\n'
           'int main(){printf("%zu %zu %zu %zu\\n", sizeof(kfp16_gemm_desc), sizeof(kfp16_mat), sizeof(kfp16_net_opts), sizeof(GPUBatchPtrs));return 0;}\n')
    exe = Path("/tmp/kfp16_sizeof_probe")
    subprocess.run(["gcc", "-x", "c", "-", f"-I{INC}", "-o", str(exe)], input=src, text=True, check=True)
    sizes = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [C.sizeof(_lib.GemmDesc), C.sizeof(_lib.KMat), C.sizeof(_lib.NetOpts), C.sizeof(_lib.GPUBatchPtrs)]


def test_compiled_for_sm100a_with_tcgen05_and_tma(built_lib):
    """the product library carries sm_100a SASS with UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld) and UTMALDG/UTMASTG (TMA)"""
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", "gemm_f16_sm100", str(built_lib)], capture_output=True, text=True)
    text = sass.stdout
    if "UTCHMMA" not in text:      # -fun needs the mangled name on some toolkits: fall back to a full dump
        text = subprocess.run(["cuobjdump", "-sass", str(built_lib)], capture_output=True, text=True).stdout
    assert "sm_100a" in text
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG"):
        assert mnemonic in text, f"{mnemonic} not found in the SASS"
    assert "HMMA.16816" not in text    # no legacy mma.sync path


def test_fails_loudly_without_a_gpu(built_lib):
    from kaldi_fp16_b200 import _lib
    from tests.conftest import _have_gpu

    if _have_gpu():
        pytest.skip("a GPU is present")
    lib = _lib.load()
    lib.ops_clear_error()
    assert not lib.ops_cublas_create()                       # NULL, with a message -- no CPU fallback
    assert lib.kfp16_last_error()
    assert lib.bridge_gpu_init(0) != 0
    assert not lib.kfp16_ctx_create(0)
    opts = _lib.NetOpts(n_seq=1, seq_len=8, train=0)
    assert not lib.kfp16_net_create(None, b"input name=input dim=8\n", C.byref(opts))
    x = np.zeros(8, np.uint16)
    assert lib.ops_gemm(None, 1, 8, 8, 1.0, x.ctypes.data, 8, x.ctypes.data, 8, 0.0, x.ctypes.data, 8) == -1
    # the peer-memory gradient exchange: no context, no communicator -- and no silent no-op
    assert not lib.kfp16_peer_comm_create(None, 0, 2, x.ctypes.data, 8)
    assert b"kfp16_peer_comm_create" in lib.kfp16_last_error()
    assert lib.kfp16_peer_allreduce_f16(None) == -1 and lib.kfp16_peer_comm_status(None) == -1


def test_product_package_never_imports_the_oracle():
    for py in (ROOT / "kaldi_fp16_b200").rglob("*.py"):
        text = py.read_text()
        assert "import oracle" not in text and "from oracle" not in text, f"{py} imports the oracle"
    for src in (ROOT / "kaldi_fp16_b200" / "csrc").iterdir():
        assert "#include \"../../oracle" not in src.read_text() and "gotorch_port" not in src.read_text(), src


def test_loss_scaler_host_logic(built_lib):
    """kaldi_loss_scaler_* (cgo_interface.cu:405-449): x0.5 on overflow, x2 after 2000 clean steps, clamped to [1, 65536]"""
    from kaldi_fp16_b200 import _lib

    lib = _lib.load()
    s = lib.kaldi_loss_scaler_create(1024.0)
    assert lib.kaldi_loss_scaler_get_scale(s) == 1024.0
    lib.kaldi_loss_scaler_update(s, 1)
    assert lib.kaldi_loss_scaler_get_scale(s) == 512.0
    for _ in range(1999):
        lib.kaldi_loss_scaler_update(s, 0)
    assert lib.kaldi_loss_scaler_get_scale(s) == 512.0
    lib.kaldi_loss_scaler_update(s, 0)
    assert lib.kaldi_loss_scaler_get_scale(s) == 1024.0
    for _ in range(12):
        lib.kaldi_loss_scaler_update(s, 1)
    assert lib.kaldi_loss_scaler_get_scale(s) == 1.0
    lib.kaldi_loss_scaler_free(s)
    big = lib.kaldi_loss_scaler_create(65536.0)
    for _ in range(2000):
        lib.kaldi_loss_scaler_update(big, 0)
    assert lib.kaldi_loss_scaler_get_scale(big) == 65536.0
    lib.kaldi_loss_scaler_free(big)
    assert lib.kaldi_loss_scaler_get_scale(None) == 1.0


def test_host_converters_and_batch_packing():
    """gpu.float32_to_fp16_bits == the oracle's truncating converter (tensor.go:158-174); pack_batch lays the minibatch
    out as bridge_batch_alloc does (bridge.cu:206-246): 256-byte aligned sections, RNE features"""
    from kaldi_fp16_b200 import gpu
    from oracle import kaldi_oracle as O

    rng = np.random.default_rng(0)
    x = np.concatenate([rng.standard_normal(5000).astype(np.float32) * s for s in (1e-6, 1e-3, 1.0, 300.0, 1e5)])
    assert np.array_equal(gpu.float32_to_fp16_bits(x), O.float32_to_fp16_bits_trunc(x))
    assert np.array_equal(gpu.fp16_bits_to_float32(np.arange(0x7C00, dtype=np.uint16)),
                          O.fp16_bits_to_float32(np.arange(0x7C00, dtype=np.uint16)))
    feats = rng.standard_normal((30, 40)).astype(np.float32) * 20
    ivec = rng.standard_normal((2, 100)).astype(np.float32)
    tb = gpu.TrainingBatch(features=feats, batch_size=2, ivectors=ivec, csr_row_ptr=[0, 2, 3, 3], csr_col_idx=[1, 2, 2],
                           csr_labels=[5, 6, 7], csr_weights=[-0.5, -0.25, 0.0])
    buf, meta = gpu.pack_batch(tb)
    off = meta["offsets"]
    assert all(v % 256 == 0 for v in off.values()) and meta["total_bytes"] % 256 == 0
    assert off["ivectors"] == 30 * 40 * 2 + (256 - (30 * 40 * 2) % 256) % 256
    got_feat = buf[: 30 * 40 * 2].view(np.uint16)
    assert np.array_equal(got_feat, O.fp16_from_float32_rne(feats).reshape(-1))          # RNE, not truncation
    assert np.array_equal(buf[off["csr_row_ptr"]: off["csr_row_ptr"] + 16].view(np.int32), [0, 2, 3, 3])
    assert np.array_equal(buf[off["csr_weights"]: off["csr_weights"] + 12].view(np.float32), np.float32([-0.5, -0.25, 0.0]))
    assert (meta["num_states"], meta["num_arcs"]) == (3, 3)
    with pytest.raises(ValueError):
        gpu.pack_batch(gpu.TrainingBatch(features=feats, batch_size=2, csr_row_ptr=[0, 1], csr_col_idx=[0], csr_labels=[], csr_weights=[0.0]))
