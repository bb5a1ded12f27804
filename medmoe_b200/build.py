"""In-tree build of libmedmoe_b200.so (sm_100a only).

`python -m medmoe_b200.build` or `__graft_entry__.build()`.  nvcc cross-compiles without a
GPU; the .so stays in the tree (git-ignored) so it travels with the snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB_DIR = PKG / "lib"
LIB_PATH = LIB_DIR / "libmedmoe_b200.so"
OBJ_DIR = PKG / "build"

SOURCES = ["api_core.cu", "b2b.cu", "router.cu", "dispatch.cu", "combine.cu", "loss.cu", "infonce_fused.cu", "local_loss.cu", "p2p.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
] + os.environ.get("MEDMOE_NVCC_EXTRA", "").split()   # e.g. -DMM_COMBINE_MINBLOCKS=1 for tuning experiments


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found: medmoe_b200 needs the CUDA 12.9 toolkit to build")
    return cand


def _source_digest() -> str:
    h = hashlib.sha256()
    for p in sorted(CSRC.iterdir()):
        if p.suffix in (".cu", ".cuh", ".h"):
            h.update(p.name.encode())
            h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu for sm_100a and link the shared library. Returns its path."""
    LIB_DIR.mkdir(exist_ok=True)
    OBJ_DIR.mkdir(exist_ok=True)
    stamp = LIB_DIR / "build.stamp"
    digest = _source_digest()
    if not force and LIB_PATH.exists() and stamp.exists() and stamp.read_text().strip() == digest:
        return LIB_PATH
    nvcc = _nvcc()

    variant = bool(os.environ.get("MEDMOE_LIB_OUT"))

    def compile_one(src: str) -> Path:
        obj = OBJ_DIR / (src + (".variant.o" if variant else ".o"))      # a variant never overwrites the product's objects
        cmd = [nvcc, *NVCC_FLAGS, "-I", str(CSRC), "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        return obj

    # MEDMOE_BUILD_ONLY="b2b.cu,..." (tuning experiments): recompile just these, reuse the other objects of the last build
    only = [x for x in os.environ.get("MEDMOE_BUILD_ONLY", "").split(",") if x]

    def compile_or_reuse(src: str) -> Path:
        obj = OBJ_DIR / (src + ".o")
        return obj if only and src not in only and obj.exists() else compile_one(src)

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_or_reuse, SOURCES))
    # static cudart: the library must load (and export its symbols) on a box without a driver
    # MEDMOE_LIB_OUT=path (tuning experiments): link a variant library next to the product one; load it with MEDMOE_LIB=path
    out = Path(os.environ["MEDMOE_LIB_OUT"]).resolve() if os.environ.get("MEDMOE_LIB_OUT") else LIB_PATH
    out.parent.mkdir(parents=True, exist_ok=True)
    link = [nvcc, "-shared", "-o", str(out), *map(str, objs), "-cudart", "static",
            "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if out == LIB_PATH:
        stamp.write_text(digest)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
