"""Build ``libcrb200.so`` (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

Equivalent of the reference's ``cellular_raza-core/build.rs:9-19`` (``cc`` + nvcc for the ``cara``
backend), with the architecture flags that file lacks.  ``python -m cellular_raza_b200.build``.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB_DIR = PKG / "lib"
LIB = LIB_DIR / "libcrb200.so"
SOURCES = [CSRC / "crb_api.cu", CSRC / "crb_tile.cu", CSRC / "crb_tile2.cu", CSRC / "crb_tile3.cu"]
OBJ_DIR = PKG / "build"
HEADERS = [p for p in CSRC.glob("*.cuh")] + [PKG.parent / "include" / "crb.h"]

NVCC_FLAGS = [
    "-O3",
    "-std=c++17",
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "--shared",
    "-Xcompiler",
    "-fPIC",
]


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found; cannot build libcrb200.so")
    return p


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any(f.stat().st_mtime > t for f in SOURCES + HEADERS + [Path(__file__)])


def build(force: bool = False, verbose: bool = False, extra: list[str] | None = None) -> Path:
    """one object per translation unit, compiled in parallel (the tile kernel families dominate the compile time),
    then one link step"""
    if not force and not needs_build():
        return LIB
    from concurrent.futures import ThreadPoolExecutor

    LIB_DIR.mkdir(exist_ok=True)
    OBJ_DIR.mkdir(exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != "--shared"]

    def compile_one(src: Path):
        obj = OBJ_DIR / (src.stem + ".o")
        cmd = [nvcc_path(), *flags, *(extra or []), "-c", "-o", str(obj), str(src)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        res = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, res

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    for src, obj, res in results:
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError(f"nvcc failed compiling {src.name}")
        if verbose:
            print(f"==== {src.name}\n" + res.stdout + res.stderr)
    cmd = [nvcc_path(), "--shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB),
           *[str(o) for _, o, _ in results], "-lpthread"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libcrb200.so")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
