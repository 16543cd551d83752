"""Build libtcdiff_sm100a.so in-tree with nvcc (sm_100a only; cross-compiles without a GPU).

    python -m tcdiff_b200.build [--force] [--verbose]

The .so lands in tcdiff_b200/lib/ (git-ignored, travels to the GPU box with the snapshot).
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, "libtcdiff_sm100a.so")
SOURCES = ["api.cu", "step.cu", "fk.cu", "norm.cu", "simt.cu", "gemm_tc.cu", "gemm_tc2.cu", "gemm_frn.cu", "gemm_wgrad.cu", "attention_tc.cu", "train_ops.cu", "train_ops16.cu", "attention_bwd.cu", "attention_bwd_tc.cu", "optim.cu", "traj.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr"]


def _digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".h")):
                h.update(f.encode())
                h.update(open(os.path.join(root, f), "rb").read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False, defines=(), out=None):
    """Compile every source for sm_100a and link the shared library.  The default build is skipped only when a stamp of the
    sources, headers and flags matches the library on disk (the returned path's `.built` attribute is not needed: callers that
    want to know read build_mode()).  defines / out: an A/B build with other csrc/tuning.cuh choices, e.g.
    build(defines=["TCD_TUNE_GELU_RAT=1"], out="libtcdiff_ab_gelu.so") — a development aid; the product loads LIB only."""
    os.makedirs(LIBDIR, exist_ok=True)
    lib = LIB if out is None else os.path.join(LIBDIR, out)
    objdir = OBJDIR if out is None else os.path.join(OBJDIR, os.path.splitext(out)[0])
    os.makedirs(objdir, exist_ok=True)
    stamp = lib + ".sha256" if out is not None else os.path.join(LIBDIR, "build.sha256")
    dflags = ["-D" + d for d in defines]
    dig = _digest() + "|" + " ".join(dflags)
    if not force and os.path.exists(lib) and os.path.exists(stamp) and open(stamp).read() == dig:
        return lib
    extra = (["-Xptxas", "-v"] if verbose else []) + dflags

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [NVCC, *FLAGS, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [NVCC, "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static",
           "-Xcompiler", "-fPIC"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    open(stamp, "w").write(dig)
    return lib


if __name__ == "__main__":
    argv = sys.argv[1:]
    defs = [argv[i + 1] for i, a in enumerate(argv) if a == "--define"]
    outs = [argv[i + 1] for i, a in enumerate(argv) if a == "--out"]
    print(build(force="--force" in argv, verbose="--verbose" in argv, defines=defs, out=outs[0] if outs else None))
