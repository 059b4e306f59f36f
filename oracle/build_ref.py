"""Build the UNMODIFIED reference pointnet2_batch CUDA extension into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Compiles the reference's own sources *where they lie*
under /root/reference (pcdet/ops/pointnet2/pointnet2_batch/src/*.{cpp,cu}; the
same file list as the reference setup.py:104-119) with nvcc for sm_100 and
writes ONLY build products (build.ninja, *.o, *.so) into oracle/_ref/.  No
reference source is copied into this repo.  oracle/_ref/ is git-ignored but
travels to the GPU box with gpurun, where /root/reference does not exist.

The resulting module `pointnet2_batch_cuda_ref` exports the 9 pybind functions
of pointnet2_api.cpp:10-24 and is used by tests/ and bench.py as (i) the
bit-exact GPU oracle for indices and (ii) the "reference CUDA-op pipeline"
timing arm.  It is never imported by the product package.
"""
import glob
import os
import sys

REF = os.environ.get("PDM_REFERENCE_ROOT", "/root/reference")
SRC = os.path.join(REF, "pcdet/ops/pointnet2/pointnet2_batch/src")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
NAME = "pointnet2_batch_cuda_ref"


def so_path():
    return os.path.join(OUT, NAME + ".so")


def build(verbose=False):
    """Compile if the reference tree is present; otherwise keep the prebuilt .so."""
    if not os.path.isdir(SRC):
        return so_path() if os.path.exists(so_path()) else None
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    os.environ.setdefault("MAX_JOBS", "8")
    from torch.utils.cpp_extension import load
    sources = sorted(glob.glob(os.path.join(SRC, "*.cpp")) + glob.glob(os.path.join(SRC, "*.cu")))
    load(name=NAME, sources=sources, build_directory=OUT, verbose=verbose,
         is_python_module=False,
         extra_cflags=["-w"], extra_cuda_cflags=["-w", "-lineinfo"])
    return so_path()


def load_ref():
    """Import the prebuilt reference extension (GPU box or here). Returns module or None."""
    p = so_path()
    if not os.path.exists(p):
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch symbols must be loaded first)
    spec = importlib.util.spec_from_file_location(NAME, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
