"""Build libspt_b200.so in-tree with nvcc for sm_100a (B200) only.

    python -m superpoint_transformer_b200.csrc.build [--force] [--verbose]

The library is a plain C-ABI shared object (include/spt_b200.h): no torch
headers, no pybind.  nvcc cross-compiles without a GPU, so this runs on the CPU
build box; the resulting .so travels to the GPU box with the snapshot.
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
LIB_NAME = "libspt_b200.so"
LIB_PATH = os.path.join(PKG, LIB_NAME)
STAMP_PATH = os.path.join(PKG, ".libspt_b200.stamp")

SOURCES = [
    "index.cu",
    "segment.cu",
    "norm.cu",
    "attention.cu",
    "gemm.cu",
    "gemm_umma.cu",
    "edge_features.cu",
    "vrpe.cu",
    "batch.cu",
    "select.cu",
    "sample.cu",
]
HEADERS = sorted(f for f in os.listdir(HERE) if f.endswith((".cuh", ".h"))) + \
    [os.path.join(ROOT, "include", "spt_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xptxas=-v",
    "-Xcompiler", "-fPIC",
    "-shared",
] + (["-DSPT_WATCHDOG"] if os.environ.get("SPT_WATCHDOG") else []) \
  + [f"-D{d}" for d in os.environ.get("SPT_NVCC_DEFINES", "").split() if d]   # tuning experiments


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; cannot build libspt_b200.so")
    return nvcc


def _fingerprint():
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        p = f if os.path.isabs(f) else os.path.join(HERE, f)
        with open(p, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def needs_build():
    if not os.path.exists(LIB_PATH) or not os.path.exists(STAMP_PATH):
        return True
    with open(STAMP_PATH) as fh:
        return fh.read().strip() != _fingerprint()


def build(force=False, verbose=False):
    """Compile every .cu into one shared library. Returns the library path."""
    if not force and not needs_build():
        return LIB_PATH
    objs = []
    obj_dir = os.path.join(PKG, "build")
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = _nvcc()
    compile_flags = [f for f in NVCC_FLAGS if f != "-shared"]
    procs = []
    for src in SOURCES:
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        cmd = [nvcc, *compile_flags, "-c", os.path.join(HERE, src), "-o", obj]
        procs.append((src, cmd, subprocess.Popen(
            cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + out + "\n")
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
    link = [nvcc, "-shared", "-Xcompiler", "-fPIC",
            "-gencode", "arch=compute_100a,code=sm_100a", *objs, "-o", LIB_PATH]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("nvcc link failed")
    with open(STAMP_PATH, "w") as fh:
        fh.write(_fingerprint())
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
