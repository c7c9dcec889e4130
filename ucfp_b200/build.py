"""Builds libucfp_cuda.so in-tree with nvcc for sm_100a (no torch involved).

    python -m ucfp_b200.build [--force] [--verbose]

The .so lands next to this file so that it travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
SO = os.path.join(HERE, "libucfp_cuda.so")
SOURCES = ["api.cu", "corpus.cu", "batcher.cu", "group.cu", "multihash.cu", "jpeg.cu", "hamming.cu", "jaccard.cu", "cosine.cu", "image.cu", "merge.cu"]
# Per-file extra flags.  image.cu must not contract a*b+c into FMA: the hash spec fixes
# separately rounded mul and add (docs/HASH_SPEC.md section 2).
EXTRA = {"image.cu": ["-fmad=false"], "cosine.cu": ["-fmad=false"]}
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "-Xcompiler", "-fPIC,-fvisibility=hidden,-O2", "--expt-relaxed-constexpr"]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _deps_mtime() -> float:
    m = 0.0
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in os.listdir(root):
            if f.endswith((".cu", ".cuh", ".h")):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return max(m, os.path.getmtime(__file__))


def build(force: bool = False, verbose: bool = False, experiments: bool = False) -> str:
    """experiments=True also compiles csrc/hamming_experiments.cuh (round-2 tensor-scan schedules that were measured and not
    adopted; selected with UCFP_HAMMING_MMA_V / UCFP_HAMMING_EPI_WARPS)."""
    extra_defs = os.environ.get("UCFP_BUILD_DEFINES", "").split()   # developer: -D switches for A/B builds on the GPU box
    if not force and not experiments and not extra_defs and os.path.exists(SO) and os.path.getmtime(SO) >= _deps_mtime():
        return SO
    os.makedirs(OBJ, exist_ok=True)
    cc = nvcc()

    def compile_one(src: str) -> str:
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        cmd = [cc, *NVCC_FLAGS, *EXTRA.get(src, []), *(["-DUCFP_HAMMING_EXPERIMENTS"] if experiments and src == "hamming.cu" else []), *extra_defs,
               "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    link = [cc, "-shared", "-o", SO, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart", "-lcuda", "-ldl", "-lpthread"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, experiments="--experiments" in sys.argv))
