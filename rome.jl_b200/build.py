"""Build librome_b200.so (sm_100a only) in-tree with nvcc.  Used by __graft_entry__.build().
One translation unit per factor-family group, compiled in parallel, then linked into one shared library."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
SO = os.path.join(HERE, "librome_b200.so")
SOURCES = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + \
    [os.path.join("..", "..", "include", "rome_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CFLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
          "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-ccbin", "/usr/bin/g++"]
LFLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++",
          "-cudart", "static"]


def _newest_header() -> float:
    return max(os.path.getmtime(os.path.join(CSRC, h)) for h in HEADERS + [os.path.abspath(__file__)])


def stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return _newest_header() > t or any(os.path.getmtime(os.path.join(CSRC, s)) > t for s in SOURCES)


def _compile(src: str, force: bool, verbose: bool) -> str:
    obj = os.path.join(OBJ, src[:-3] + ".o")
    path = os.path.join(CSRC, src)
    if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(path), _newest_header()):
        return ""
    cmd = [NVCC] + CFLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n" + r.stdout)
    return r.stdout


def build(force: bool = False, verbose: bool = False) -> str:
    if force or stale():
        os.makedirs(OBJ, exist_ok=True)
        with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
            logs = list(ex.map(lambda s: _compile(s, force, verbose), SOURCES))
        if verbose:
            print("\n".join(logs))
        cmd = [NVCC] + LFLAGS + [os.path.join(OBJ, s[:-3] + ".o") for s in SOURCES] + ["-o", SO]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout)
    return SO


if __name__ == "__main__":
    print(build(force="-f" in sys.argv, verbose="-v" in sys.argv))
