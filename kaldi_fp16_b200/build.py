"""In-tree build of the native library: nvcc -> kaldi_fp16_b200/libkaldi_fp16.so (sm_100a only).

No JIT cache and no torch.utils.cpp_extension: the .so is a plain C-ABI shared library (the same
file a Go / cgo caller would link) and has to travel with the repo snapshot
to the GPU box.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OBJ = PKG / "_obj"
LIB = PKG / "libkaldi_fp16.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17", *os.environ.get("KFP16_NVCC_EXTRA", "").split(),
    "-Xcompiler", "-fPIC,-fvisibility=default",
]


def _sources() -> list[Path]:
    return sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cpp")))


def _deps_hash(src: Path) -> str:
    h = hashlib.sha1()
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(src.read_bytes())
    for hdr in sorted(list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list((PKG.parent / "include").glob("*.h"))):
        h.update(hdr.read_bytes())
    return h.hexdigest()


def _compile(src: Path) -> Path:
    obj = OBJ / (src.name + ".o")
    stamp = OBJ / (src.name + ".sha1")
    want = _deps_hash(src)
    if obj.exists() and stamp.exists() and stamp.read_text() == want:
        return obj
    cmd = ["nvcc", *NVCC_FLAGS, "-x", "cu", "-c", str(src), "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
    stamp.write_text(want)
    return obj


def build(verbose: bool = True) -> Path:
    """Compile every CUDA source for sm_100a and link the shared library (incremental)."""
    OBJ.mkdir(exist_ok=True)
    srcs = _sources()
    if not srcs:
        raise RuntimeError("no CUDA sources found")
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(_compile, srcs))
    newest = max(o.stat().st_mtime for o in objs)
    if (not LIB.exists()) or LIB.stat().st_mtime < newest:
        cmd = ["nvcc", "-shared", "-Xlinker", "-Bsymbolic", "-o", str(LIB), *map(str, objs), "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(f"[kaldi_fp16_b200] built {LIB} from {len(srcs)} sources", file=sys.stderr)
    return LIB


if __name__ == "__main__":
    build()
