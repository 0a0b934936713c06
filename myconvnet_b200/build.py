"""Build libmcn.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

The .so is git-ignored but travels to the GPU box with the working tree, so it is built here
(nvcc cross-compiles without a GPU) and only rebuilt when a source is newer than the library.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# MCN_ROLE_TIMING=1: the instrumented build (per-role stall counters in the conv kernels) as a
# separate library, never loaded by default (MCN_LIB=.../libmcn_timing.so selects it)
TIMING = os.environ.get("MCN_ROLE_TIMING", "0") == "1"
LIB = os.path.join(HERE, "libmcn_timing.so" if TIMING else "libmcn.so")
SOURCES = ["runtime.cu", "conv_tc.cu", "conv_direct.cu", "bn.cu", "pool.cu", "eltwise.cu",
           "loss.cu", "opt.cu", "comm.cu", "dropout.cu", "dwconv.cu", "norm_extra.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]
# Approximate division / sqrt and flush-to-zero only where they cannot reach the fp32 parity of the
# optimiser, loss and normalisation maths: the tensor-core conv kernels (epilogue is adds and
# casts), the CUDA-core convs, pooling and the element-wise kernels (which call __expf explicitly).
# bn.cu keeps it too: its cancellation-prone maths (mean / variance / invstd) is explicit fp64, and
# without fast-math the runtime-selected activation (tanhf, precise division) doubles the register
# count of the streaming kernels (measured: bn_apply 142 -> 293 us on a 411 MB tensor).
FAST_MATH = {"conv_tc.cu", "conv_direct.cu", "pool.cu", "eltwise.cu", "dwconv.cu", "bn.cu", "dropout.cu"}


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, "..", "include", "mcn.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every .cu into objects (in parallel) and link libmcn.so. Returns the path."""
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(HERE, "..", "build", "obj_timing" if TIMING else "obj")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    objs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, *(["-DMCN_ROLE_TIMING"] if TIMING else []),
               *(["--use_fast_math"] if src in FAST_MATH else []), "-c",
               os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(out.decode())
        if p.returncode != 0:
            failed = True
            sys.stderr.write("nvcc failed for %s\n" % src)
    if failed:
        raise RuntimeError("libmcn build failed")
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB, *objs,
                           "-lcudart"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
