"""Builds ``libmfac.so`` in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m meanflow_audio_codec_b200.build [--force]

One translation unit per ``csrc/*.cu``; objects are cached under ``csrc/_build/`` keyed
by source mtime so rebuilding after a single-file edit takes seconds.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OUT = PKG / "libmfac.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17", "--use_fast_math=false",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
]


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _headers_mtime() -> float:
    hs = list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "mfac.h"]
    return max(h.stat().st_mtime for h in hs)


def build(force: bool = False, verbose: bool = False) -> Path:
    bdir = CSRC / "_build"
    bdir.mkdir(exist_ok=True)
    hm = _headers_mtime()
    jobs = []
    for src in _sources():
        obj = bdir / (src.stem + ".o")
        if force or not obj.exists() or obj.stat().st_mtime < max(src.stat().st_mtime, hm):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [NVCC, *[f for f in FLAGS if f != "--use_fast_math=false"], "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, r

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for src, r in ex.map(compile_one, jobs):
            if verbose and r.stderr:
                print(r.stderr, file=sys.stderr)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
    objs = [str(bdir / (s.stem + ".o")) for s in _sources()]
    if jobs or not OUT.exists():
        cmd = [NVCC, "-shared", "-o", str(OUT), *objs, "-gencode", "arch=compute_100a,code=sm_100a",
               "-Xcompiler", "-fPIC", "-cudart", "shared", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return OUT


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
