"""Builds innfer_b200/lib/libinnfer_b200.so (sm_100a only) with nvcc, in-tree.

The library is a plain C-ABI shared object (include/innfer_b200.h); nothing here depends on
torch headers.  ``python -m innfer_b200.build`` or ``__graft_entry__.build()`` run this.
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIBNAME = "libinnfer_b200.so"
SOURCES = ["sync_ops.cu", "conv_tc.cu", "conv_rows.cu", "conv_up.cu", "tmap.cu", "layers.cu", "pixel_ops.cu", "conv_direct.cu", "color_fix.cu", "pan_ops.cu",
           "i2i.cu", "engine.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


if os.environ.get("INNFER_EXPERIMENTS_BUILD"):  # timing experiments that produce wrong results (INNFER_I2I_DEBUG, INNFER_ROWS_DX0)
    NVCC_FLAGS.append("-DINNFER_EXPERIMENTS")
if os.environ.get("INNFER_TRACE_BUILD"):  # debugging build: clock64 tracing inside conv_rows (tests/gpu_bringup.py --stage trace)
    NVCC_FLAGS.append("-DINNFER_ROWS_TRACE")


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; the CUDA extension cannot be built")
    return exe


def _digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for name in sorted(os.listdir(root)):
            if name.endswith((".cu", ".cuh", ".h")):
                with open(os.path.join(root, name), "rb") as f:
                    h.update(name.encode())
                    h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def lib_path():
    return os.path.join(LIBDIR, LIBNAME)


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a and link the shared library. Returns its path."""
    os.makedirs(LIBDIR, exist_ok=True)
    stamp = os.path.join(LIBDIR, "build.stamp")
    digest = _digest()
    if not force and os.path.exists(lib_path()) and os.path.exists(stamp):
        with open(stamp) as f:
            if f.read().strip() == digest:
                return lib_path()
    nvcc = _nvcc()
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd))
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    objs = []
    for src, obj, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, out.decode(errors="replace")))
        objs.append(obj)
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib_path()] + objs
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if out.returncode != 0:
        raise RuntimeError("link failed:\n%s" % out.stdout.decode(errors="replace"))
    with open(stamp, "w") as f:
        f.write(digest)
    return lib_path()


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
