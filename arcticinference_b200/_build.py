"""In-tree build of libarctic_b200.so: plain nvcc, sm_100a only, no torch headers (seconds, not
minutes).  The .so is git-ignored but travels to the GPU box with the gpurun snapshot.

Two flavours from the same sources:
  libarctic_b200.so      the product: compiled-in launch parameters, no ab200_debug_* symbol
  libarctic_b200_dbg.so  -DAB200_DEBUG: adds the exhaustive-search / comparator hooks and
                         ab200_debug_set_tunable; loaded by tests/ and tools/ only (_capi.debug_lib)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libarctic_b200.so")
LIB_DEBUG = os.path.join(HERE, "libarctic_b200_dbg.so")
SOURCES = ["kv_write.cu", "kv_rope_write.cu", "lstm_cell.cu", "mlp_cell.cu", "gemm_sm100.cu", "quant_fp8.cu", "speculator.cu"]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build(lib: str = LIB) -> bool:
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "arctic_b200.h"))
    deps.append(os.path.abspath(__file__))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, debug: bool = False) -> str:
    lib = LIB_DEBUG if debug else LIB
    if not force and not needs_build(lib):
        return lib
    nvcc = nvcc_path()
    objdir = os.path.join(HERE, "build", "dbg" if debug else "rel")
    os.makedirs(objdir, exist_ok=True)
    flags = NVCC_FLAGS + (["-DAB200_DEBUG"] if debug else [])
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *flags, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for src, obj, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
        objs.append(obj)
    tmp = lib + ".tmp"
    subprocess.check_call([nvcc, "-shared", "-o", tmp, *objs, "-Xcompiler", "-fPIC",
                           "-gencode", "arch=compute_100a,code=sm_100a"])
    os.replace(tmp, lib)
    return lib


def build_all(force: bool = False, verbose: bool = False):
    return build(force, verbose, debug=False), build(force, verbose, debug=True)


if __name__ == "__main__":
    print(*build_all(force="--force" in sys.argv, verbose="-v" in sys.argv), sep="\n")
