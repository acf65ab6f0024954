"""Build librome_b200.so (sm_100a only) in-tree with nvcc.  Used by __graft_entry__.build()."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "librome_b200.so")
SOURCES = ["factor_kernels.cu", "rome_b200_api.cu"]
HEADERS = ["device_utils.cuh", "tables.h", os.path.join("..", "..", "include", "rome_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--shared",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-ccbin", "/usr/bin/g++", "-cudart", "static"]


def stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if force or stale():
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
              [os.path.join(CSRC, s) for s in SOURCES] + ["-o", SO]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + r.stdout)
        if verbose:
            print(r.stdout)
    return SO


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
