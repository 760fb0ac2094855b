"""Builds libb200convlstm.so (hand-written sm_100a CUDA kernels + the C ABI of include/b200_convlstm.h).

In-tree build with plain nvcc: every csrc/*.cu is compiled to an object (in parallel) and linked into
one shared library next to this file.  nvcc cross-compiles without a GPU, so this runs in the build
container; the resulting .so travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libb200convlstm.so")
# the torch custom-op face of the C ABI: csrc/torch_ops.cpp (generated from the header by tools/gen_torch_ops.py),
# compiled with the host compiler against torch's headers and linked to LIB
TORCH_LIB = os.path.join(HERE, "libb200convlstm_torch.so")
TORCH_SRC = os.path.join(CSRC, "torch_ops.cpp")
CXX = os.environ.get("CXX", "g++")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
# no --use_fast_math: the fp32 check mode needs IEEE tanhf/expf/division; the bf16 kernels ask for
# tanh.approx explicitly where they want it
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def build(force: bool = False, verbose: bool = False) -> str:
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "b200_convlstm.h"))
    os.makedirs(OBJ, exist_ok=True)
    hdr_time = _newest(hdrs)
    jobs = []
    objs = []
    for s in srcs:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_time):
            jobs.append([NVCC, *FLAGS, "-c", src, "-o", obj] + (["-Xptxas", "-v"] if verbose else []))

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, r

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for cmd, r in ex.map(run, jobs):
                if verbose or r.returncode != 0:
                    sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
                if r.returncode != 0:
                    raise RuntimeError(f"nvcc failed for {cmd[-3]}")
    if jobs or not os.path.exists(LIB) or os.path.getmtime(LIB) < _newest(objs):
        cmd = [NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    build_torch_ops(force)
    return LIB


def build_torch_ops(force: bool = False) -> str:
    """g++ -shared csrc/torch_ops.cpp -> libb200convlstm_torch.so (TORCH_LIBRARY(b200convlstm, ...) registrations)."""
    deps = [TORCH_SRC, os.path.join(os.path.dirname(HERE), "include", "b200_convlstm.h"), LIB]
    if not force and os.path.exists(TORCH_LIB) and os.path.getmtime(TORCH_LIB) >= _newest(deps):
        return TORCH_LIB
    import torch
    from torch.utils import cpp_extension as ce
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    cmd = [CXX, "-O2", "-std=c++17", "-fPIC", "-shared", "-w", f"-D_GLIBCXX_USE_CXX11_ABI={int(torch.compiled_with_cxx11_abi())}",
           *[f"-I{d}" for d in ce.include_paths()], f"-I{cuda_inc}", TORCH_SRC, "-o", TORCH_LIB,
           f"-L{tlib}", "-ltorch", "-ltorch_cpu", "-lc10", "-ltorch_cuda", "-lc10_cuda",
           f"-L{HERE}", "-lb200convlstm", "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{tlib}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("building libb200convlstm_torch.so failed")
    return TORCH_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
