"""Build libfd_b200.so (sm_100a only) in-tree with nvcc.  No GPU needed: nvcc cross-compiles.

    python pytorch-face-detection-from-scratch_b200/csrc/build.py [--force]
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
LIB = os.path.join(PKG, "libfd_b200.so")
SOURCES = ["fd_host.cu", "conv3x3_tc.cu", "conv3x3_wide.cu", "resblock_chain.cu", "wgrad3x3_tc.cu", "wgrad3x3_wide.cu", "layers.cu", "yolo_kernels.cu", "ssd_kernels.cu", "stem_tc.cu", "comm.cu", "sepblock.cu", "stem_s2_tc.cu", "head_tc.cu", "pw_gemm_tc.cu", "mbv3_kernels.cu", "mbv3_stem_tc.cu", "misc_kernels.cu", "ssd_head.cu"]
HEADERS = ["fd_host.h", "fd_ptx.cuh", os.path.join("..", "..", "include", "fd_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _digest():
    h = hashlib.sha256()
    for f in SOURCES + HEADERS + ["build.py"]:
        with open(os.path.join(HERE, f), "rb") as fp:
            h.update(fp.read())
    return h.hexdigest()


def build(force=False, verbose=False):
    stamp = os.path.join(HERE, "build", "stamp")
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)

    def cc(src):
        obj = os.path.join(HERE, "build", src.replace(".cu", ".o"))
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(HERE, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(cc, SOURCES))
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as fp:
        fp.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
