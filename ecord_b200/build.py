"""Build libecord_b200.so in-tree with nvcc for sm_100a (B200).  No JIT cache, no torch extension:
the library is a plain C-ABI shared object (include/ecord_b200.h) loaded with ctypes.

    python -m ecord_b200.build [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["env.cu", "gnn.cu", "policy.cu", "gemm_tc.cu", "tail_tc.cu", "qselect.cu", "qtc.cu", "rollout.cu"]
LIB = os.path.join(HERE, "libecord_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--use_fast_math=false", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "ecord_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), out=None):
    """defines / out: diagnostic variants (e.g. -DECORD_GRU_STAGES=6) built next to the product library and selected at run
    time with ECORD_B200_LIB=<path> (see _lib.py)."""
    if not force and not needs_build() and not defines and out is None:
        return LIB
    nvcc = _nvcc()
    objs = []
    os.makedirs(os.path.join(HERE, "_build"), exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"] + [f"-D{d}" for d in defines]
    tag = ("_" + "_".join(d.replace("=", "") for d in defines)) if defines else ""
    for src in SOURCES:
        obj = os.path.join(HERE, "_build", src.replace(".cu", tag + ".o"))
        cmd = [nvcc] + flags + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
        with open(obj + ".ptxas.txt", "w") as f:
            f.write(r.stderr)
        objs.append(obj)
    lib = out or LIB
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs
    subprocess.check_call(cmd)
    return lib


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs, out=outs[0] if outs else None))
