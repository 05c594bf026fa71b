"""Builds the sm_100a CUDA library in-tree (zigzag_b200/libzigzag_gpu.so) with nvcc.  No GPU is needed to build."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = [os.path.join(HERE, "csrc", "zz_gpu.cu")]
DEPS = [os.path.join(HERE, "csrc", f) for f in sorted(os.listdir(os.path.join(HERE, "csrc")))] + \
       [os.path.join(HERE, "..", "include", "zigzag_gpu.h")]
OUT = os.environ.get("ZZ_BUILD_OUT", os.path.join(HERE, "libzigzag_gpu.so"))   # kernel-tuning builds go elsewhere

# -fmad=false: the kernels keep the reference's operation order (a*b+c is not contracted), see DESIGN.md "Numerics"
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-fmad=false", "-std=c++17",
              "--shared", "-Xcompiler", "-fPIC,-pthread", "-cudart", "static"]


def nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in DEPS)


DBG_OUT = os.path.join(HERE, "libzigzag_gpu_dbg.so")   # the same sources with -DZZ_BOUNDS_CHECK (device-side index asserts), tests only


def build_debug(force=False):
    """The bounds-checked build the GPU suite runs once (tests/test_gpu_debug_build.py)."""
    if not force and os.path.exists(DBG_OUT) and all(os.path.getmtime(d) <= os.path.getmtime(DBG_OUT) for d in DEPS):
        return DBG_OUT
    return build(force=True, out=DBG_OUT, extra=["-DZZ_BOUNDS_CHECK"])


def build(force=False, verbose=False, out=None, extra=None):
    out = out or OUT
    if out == OUT and not force and not needs_build():
        return OUT
    extra = (extra or []) + os.environ.get("ZZ_NVCC_EXTRA", "").split()     # e.g. -DZZ_ILP=4 for kernel-tuning experiments
    cmd = [nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + SRC
    env = dict(os.environ)
    env.pop("CC", None); env.pop("CXX", None)  # the image's $CC points at a gcc without the host libs nvcc needs
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if verbose:
        sys.stderr.write(r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
