"""Builds hoomd_blue_b200/libhoomd_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch
extension machinery: the product is a plain C-ABI shared library, include/hoomd_b200.h).

    python -m hoomd_blue_b200.build [--force]
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# HB2_VARIANT / HB2_EXTRA_FLAGS build an experiment library beside the product one (kernel tuning:
# `HB2_VARIANT=mb10 HB2_EXTRA_FLAGS=-DPAIR_MIN_BLOCKS=10 python -m hoomd_blue_b200.build`, loaded with
# HB2_LIB=libhoomd_b200_mb10.so)
_VAR = os.environ.get("HB2_VARIANT", "")
OBJ = os.path.join(HERE, "build" + ("_" + _VAR if _VAR else ""))
LIB = os.path.join(HERE, "libhoomd_b200" + ("_" + _VAR if _VAR else "") + ".so")
SOURCES = ["engine.cu", "capi.cu", "binning.cu", "nlist.cu", "nlist_tile.cu", "integrate.cu", "pair_lj.cu",
           "pair_dpd.cu", "pair_example.cu", "pair_more_a.cu", "pair_more_b.cu", "pair_more_c.cu", "pair_more_d.cu", "pair_more_e.cu", "pair_more_f.cu", "pair_more_g.cu", "comm.cu", "bond.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
         "-Xptxas", "-v"] + os.environ.get("HB2_EXTRA_FLAGS", "").split()


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(HERE, "..", "include", "hoomd_b200.h"))
    return hs


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src):
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    path = os.path.join(CSRC, src)
    if not _stale(obj, [path] + _headers()):
        return obj, ""
    cmd = [NVCC] + FLAGS + ["-c", path, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj, r.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    if force:
        for s in srcs:
            o = os.path.join(OBJ, s.replace(".cu", ".o"))
            if os.path.exists(o):
                os.remove(o)
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(_compile, srcs))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            if log:
                sys.stderr.write(log)
    if _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                     "-lcudart", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
