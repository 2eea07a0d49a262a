"""In-tree build of libipx.so (hand-written CUDA for sm_100a behind the C ABI of include/ipx.h).

``python -m interpax_b200._build`` or ``build()`` compiles ``csrc/*.cu`` with nvcc into
``interpax_b200/libipx.so``.  The .so is git-ignored but travels with the tree to the GPU
box; nothing is JIT-compiled into a user cache.
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libipx.so")
SOURCES = ["ipx_eval.cu", "ipx_stencil.cu", "ipx_grad.cu", "ipx_grad_tables.cu", "ipx_slopes.cu", "ipx_prep.cu", "ipx_sort.cu", "ipx_vjp.cu", "ipx_ppoly.cu"]
HEADERS = [os.path.join(CSRC, "ipx_device.cuh"), os.path.join(CSRC, "ipx_internal.cuh"), os.path.join(CSRC, "ipx_halo.cuh"), os.path.join(CSRC, "ipx_stencil_col.cuh"), os.path.join(CSRC, "ipx_eval_tile.cuh"), os.path.join(CSRC, "ipx_contract.cuh"), os.path.join(CSRC, "ipx_eval_rows.cuh"),
           os.path.join(HERE, "..", "include", "ipx.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
]
OBJDIR = os.path.join(HERE, "build")


def find_nvcc() -> str | None:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def _obj(src: str) -> str:
    return os.path.join(OBJDIR, os.path.splitext(src)[0] + ".o")


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libipx.so if it is missing or older than its sources. Returns its path.
    Every csrc/*.cu is compiled to its own object (in parallel; only the ones older than the
    sources) and the objects are linked into the shared library."""
    if not force and not is_stale():
        return LIB
    nvcc = find_nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build interpax_b200/libipx.so")
    os.makedirs(OBJDIR, exist_ok=True)
    hdr_time = max(os.path.getmtime(h) for h in HEADERS if os.path.exists(h))
    procs = []
    for src in SOURCES:
        path, obj = os.path.join(CSRC, src), _obj(src)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(path), hdr_time):
            continue
        cmd = [nvcc, *NVCC_FLAGS, "-c", path, "-o", obj]
        if verbose:
            cmd += ["-Xptxas", "-v"]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
    for src, proc in procs:
        out, err = proc.communicate()
        if proc.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n" + out + err)
        if verbose:
            sys.stderr.write(err)
    link = subprocess.run([nvcc, "-shared", "-o", LIB + ".tmp"] + [_obj(s) for s in SOURCES],
                          capture_output=True, text=True)
    if link.returncode != 0:
        raise RuntimeError("link failed:\n" + link.stdout + link.stderr)
    os.replace(LIB + ".tmp", LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
