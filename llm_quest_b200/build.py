"""Build recipe for libvfuse.so (the C-ABI CUDA library behind the drop-in modules).

Plain ``nvcc`` for sm_100a only, in-tree output (``llm_quest_b200/libvfuse.so``) so the binary
travels with a snapshot of the repository; objects go to ``llm_quest_b200/csrc/build/``.
nvcc cross-compiles without a GPU, so this also runs in CPU-only CI.

    python -m llm_quest_b200.build [--force] [--verbose]
"""

from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
BUILD = CSRC / "build"
LIB = PKG_DIR / "libvfuse.so"
INCLUDE = PKG_DIR.parent / "include"

SOURCES = ["vf_api.cu", "vf_gemm.cu", "vf_attention.cu", "vf_attention_gqa.cu", "vf_attention_small.cu", "vf_elementwise.cu", "vf_norm.cu", "vf_rope.cu", "vf_fuse.cu", "vf_preprocess.cu"]
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(ARCH_FLAGS + NVCC_FLAGS).encode())
    return h.hexdigest()


def _rpaths() -> list[str]:
    out = ["/usr/local/cuda/lib64"]
    try:  # the libcudart PyTorch ships (already mapped when torch is imported first)
        import nvidia.cuda_runtime  # type: ignore

        out.insert(0, str(Path(nvidia.cuda_runtime.__path__[0]) / "lib"))
    except Exception:
        pass
    return out


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu for sm_100a and link libvfuse.so; no-op when sources are unchanged."""
    BUILD.mkdir(parents=True, exist_ok=True)
    deps = [CSRC / s for s in SOURCES] + list(CSRC.glob("*.cuh")) + list(INCLUDE.glob("*.h"))
    stamp = BUILD / "stamp.sha256"
    digest = _digest(deps)
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == digest:
        return LIB
    nvcc = _nvcc()

    def compile_one(src: str) -> Path:
        obj = BUILD / (Path(src).stem + ".o")
        cmd = [nvcc, *ARCH_FLAGS, *NVCC_FLAGS, "-I", str(INCLUDE), "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose and r.stderr:
            print(r.stderr, flush=True)
        return obj

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    link = [nvcc, *ARCH_FLAGS, "-shared", "-cudart", "shared", "-o", str(LIB), *map(str, objs)]
    for rp in _rpaths():
        link += ["-Xlinker", f"-rpath={rp}"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    stamp.write_text(digest)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
