"""In-tree build of libfigdraw_cuda.so for sm_100a (nvcc cross-compiles without a GPU).

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.  cudart is linked statically so
the library loads in any process (python/torch, a Nim executable) without libcudart on the loader path.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(CSRC, "libfigdraw_cuda.so")
SOURCES = ["fdc_context.cu", "fdc_bin.cu", "fdc_shade.cu", "fdc_blur.cu", "fdc_flatten.cu", "fdc_glyph.cu"]
HEADERS = ["fdc_types.h", "fdc_kernels.h", "fdc_flatten.h", os.path.join("..", "..", "include", "figdraw_cuda.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--cudart", "static",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall",
    "-Xptxas", "-v",
]


def _stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        objs.append(obj)
        extra = os.environ.get("FDC_NVCC_EXTRA", "").split()
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-ccbin", "/usr/bin/g++", "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"== {src}\n{out}")
        failed = failed or p.returncode != 0
    text = "\n".join(log)
    with open(os.path.join(CSRC, "build.log"), "w") as fh:
        fh.write(text)
    if failed:
        sys.stderr.write(text)
        raise RuntimeError("nvcc failed; see figdraw_b200/csrc/build.log")
    link = [nvcc, "-shared", "--cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "/usr/bin/g++",
            "-Xcompiler", "-fPIC", "-o", OUT, *objs, "-Xlinker", "--exclude-libs,ALL"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link failed")
    if verbose:
        print(text)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
