"""Builds libvti.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("VTI_LIB", os.path.join(HERE, "libvti.so"))      # (tuning sweeps build side-by-side variants)
SOURCES = ["api.cu", "k0_ingest.cu", "k1_preprocess.cu", "k2_decode.cu", "k3_nms.cu", "k4_masks.cu", "k5_measure.cu", "k6_overlay.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
COMMON = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
          "-Xcompiler", "-ffp-contract=off"]
PER_FILE = {"k5_measure.cu": ["--fmad=false"]}     # numpy / OpenCV do not fuse multiply-adds (fp64 parity)


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "vti.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_lib(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [NVCC, *COMMON, *PER_FILE.get(src, []), *os.environ.get("VTI_NVCC_FLAGS", "").split(), "-c",
               os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- nvcc {src}\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    # nvJPEG (toolkit library, same image on the GPU box) for vti_encode_jpeg; rpath so that dlopen finds it
    cuda_lib = os.path.join(os.path.dirname(os.path.dirname(NVCC)), "lib64")
    subprocess.check_call([NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
                           f"-L{cuda_lib}", "-lnvjpeg", "-Xlinker", f"-rpath={cuda_lib}"])
    return LIB


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
