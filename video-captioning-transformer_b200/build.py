"""Build libvct_b200.so (sm_100a only) in-tree with nvcc.  No torch headers, no JIT cache:
the built .so sits next to the sources so it travels to the GPU box with the repo snapshot.

    python video-captioning-transformer_b200/build.py [--force] [--verbose]
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
TRACE = os.environ.get("VCT_TRACE_BUILD") == "1"      # per-k-block pipeline timestamps in the tcgen05 GEMM (tools/gemm_trace.py)
OUT = os.path.join(HERE, "libvct_b200_trace.so" if TRACE else "libvct_b200.so")
OBJ = os.path.join(HERE, "build_trace" if TRACE else "build")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--use_fast_math=false",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
FLAGS = [f for f in FLAGS if f != "--use_fast_math=false"] + (["-DVCT_GEMM_TRACE_KB"] if TRACE else [])


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_hash(src: str) -> str:
    h = hashlib.sha1()
    for f in [src] + sorted(os.path.join(CSRC, x) for x in os.listdir(CSRC) if x.endswith((".cuh", ".h"))) + [
            os.path.join(os.path.dirname(HERE), "include", "vct.h"), os.path.abspath(__file__)]:
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    srcs = _sources()
    objs, jobs = [], []
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s)[:-3] + ".o")
        stamp = o + ".sha1"
        hsh = _deps_hash(s)
        objs.append(o)
        if not force and os.path.isfile(o) and os.path.isfile(stamp) and open(stamp).read() == hsh:
            continue
        jobs.append((s, o, stamp, hsh))

    def compile_one(job):
        s, o, stamp, hsh = job
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {s}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        with open(stamp, "w") as f:
            f.write(hsh)
        return s

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(compile_one, jobs))
    if jobs or not os.path.isfile(OUT) or force:
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT] + objs + ["-lcuda"] 
        cmd = [c for c in cmd if c != "-lcuda"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
