"""Builds libmsda_b200.so (sm_100a only) from csrc/*.cu with nvcc -- in-tree, no JIT cache, no torch headers.

    python msda-triton_b200/build.py [--force] [--verbose]

The output (msda-triton_b200/lib/libmsda_b200.so) is git-ignored but travels to the GPU box with the tree.
"""
from __future__ import annotations

import argparse
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB_DIR = HERE / "lib"
LIB = LIB_DIR / "libmsda_b200.so"
OBJ_DIR = LIB_DIR / "obj"
INCLUDE = HERE.parent / "include"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-I", str(INCLUDE), "-I", str(CSRC),
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (set NVCC or add /usr/local/cuda/bin to PATH)")


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> Path:
    nvcc = _nvcc()
    sources = sorted(CSRC.glob("*.cu"))
    headers = sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + sorted(INCLUDE.glob("*.h"))
    OBJ_DIR.mkdir(parents=True, exist_ok=True)

    def compile_one(src: Path) -> Path:
        obj = OBJ_DIR / (src.stem + ".o")
        if force or _stale(obj, [src, *headers]):
            extra = os.environ.get("MSDA_B200_NVCC_EXTRA", "").split()   # e.g. -DMSDA_TMEM_PROF (debug builds only)
            cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", str(src), "-o", str(obj)]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd), flush=True)
            r = subprocess.run(cmd, capture_output=True, text=True)
            if verbose or r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src.name}")
        return obj

    # the translation units that instantiate the tuned kernel templates take 20-70 s each, the rest a few seconds: start
    # the heavy ones first and use the cores there are
    heavy = ("msda_fwd_tiled", "msda_fwd_module", "msda_bwd_tiled", "msda_bwd_dense", "msda_bwd_split", "msda_fwd_points",
             "msda_bwd_module", "msda_bwd_shapes")
    order = sorted(sources, key=lambda s: heavy.index(s.stem) if s.stem in heavy else len(heavy))
    workers = max(1, min(os.cpu_count() or 8, 16, len(sources)))
    with ThreadPoolExecutor(max_workers=workers) as ex:
        done = dict(zip(order, ex.map(compile_one, order)))
    objs = [done[s] for s in sources]
    if force or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB), *map(str, objs)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link of libmsda_b200.so failed")
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ns = ap.parse_args()
    print(build_library(ns.force, ns.verbose))
