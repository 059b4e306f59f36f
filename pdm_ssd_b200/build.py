"""Build libpdmops.so (hand-written sm_100a CUDA + the C ABI of include/pdm_ops.h) in-tree.

    python -m pdm_ssd_b200.build [--force] [-v]

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels to the
GPU box with gpurun.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libpdmops.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "--fmad=false",          # no implicit contraction: every fma in the kernels is spelled out
    "-Xptxas", "-v",
]
# nms.cu restates arithmetic whose reference build uses nvcc's default contraction (iou3d_nms has
# no -fmad flag, setup.py:63-72); keep decisions sit on `iou > thresh`, so it is compiled the same way
DEFAULT_FMAD = {"nms.cu"}


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        [os.path.join(HERE, "..", "include", "pdm_ops.h"), os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not stale():
        return SO
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for src in sources():
        obj = os.path.join(HERE, "build", os.path.basename(src)[:-3] + ".o")
        drop = {"-shared"} | ({"--fmad=false"} if os.path.basename(src) in DEFAULT_FMAD else set())
        cmd = [NVCC] + [f for f in FLAGS if f not in drop] + ["-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    log = []
    for src, pr in procs:
        out, _ = pr.communicate()
        log.append(out)
        if pr.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("nvcc failed on %s" % src)
    link = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC",
            "-o", SO] + objs
    subprocess.check_call(link)
    with open(os.path.join(HERE, "build", "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
