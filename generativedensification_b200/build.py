"""Build libgdr.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m generativedensification_b200.build [--force] [--verbose]

Each .cu under csrc/ is compiled to an object file in parallel
(-gencode arch=compute_100a,code=sm_100a -lineinfo, default FMA contraction, no
fast-math: the numerical contract with the reference depends on it) and linked
into generativedensification_b200/libgdr.so with the static CUDA runtime.  The
library has no torch or Python dependency.
"""
from __future__ import annotations

import contextlib
import fcntl
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libgdr.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps():
    d = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    d.append(os.path.join(HERE, "..", "include", "gdr.h"))
    d.append(os.path.abspath(__file__))
    return d


def is_fresh() -> bool:
    if not os.path.isfile(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(p) <= t for p in _deps() if os.path.exists(p))


@contextlib.contextmanager
def _build_lock():
    """Exclusive inter-process lock around a build: every rank of a multi-process run may find the library stale at the
    same time; one of them builds, the others wait here and then find it fresh."""
    os.makedirs(OBJ, exist_ok=True)
    with open(os.path.join(OBJ, ".lock"), "w") as f:
        fcntl.flock(f, fcntl.LOCK_EX)
        try:
            yield
        finally:
            fcntl.flock(f, fcntl.LOCK_UN)


def build(force: bool = False, verbose: bool = False, extra_flags=()) -> str:
    if not force and is_fresh():
        return LIB
    if not os.path.isfile(NVCC):
        raise RuntimeError(f"nvcc not found at {NVCC}; cannot build libgdr.so")
    with _build_lock():
        if not force and is_fresh():  # another process built it while we waited for the lock
            return LIB
        return _build_locked(verbose, extra_flags)


def _build_locked(verbose: bool, extra_flags) -> str:
    srcs = _sources()
    extra_flags = tuple(extra_flags) + tuple(os.environ.get("GDR_NVCC_FLAGS", "").split())  # experiments (-D...)

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        cmd = [NVCC, *NVCC_FLAGS, *extra_flags, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(compile_one, srcs))
    if verbose:
        for _, err in results:
            sys.stderr.write(err)
    objs = [o for o, _ in results]
    # objects of sources that no longer exist must not be linked; link into a temporary file and rename it over the
    # library, so a process that is loading libgdr.so never sees a half-written file
    tmp = f"{LIB}.tmp.{os.getpid()}"
    cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp, *objs, "-Xcompiler", "-fPIC"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        with contextlib.suppress(OSError):
            os.unlink(tmp)
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
