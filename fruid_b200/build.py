"""Build recipe for libfruid_b200.so (hand-written sm_100a CUDA behind a C ABI).

nvcc cross-compiles without a GPU.  Flags that matter for parity with the reference
(SURVEY.md 8c): -fmad=false (Rust never contracts a*b+c), no --use_fast_math, default
-prec-div=true -prec-sqrt=true -ftz=false.  -lineinfo keeps ncu's source page usable.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libfruid_b200.so")

DBG_LIB = os.path.join(LIBDIR, "libfruid_b200_dbg.so")  # -DFRUID_DEBUG: tail delays + poisoned scratch (ordering tests)
OBJDIR = os.path.join(HERE, "build")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-O2,-Wall",
]


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]


def _src_hash() -> str:
    import hashlib
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
    files += [os.path.join(HERE, "..", "include", "fruid_b200.h"), os.path.abspath(__file__)]
    for f in files:
        with open(f, "rb") as fh:
            h.update(os.path.basename(f).encode() + b"\0" + fh.read())
    return h.hexdigest()


def _stale(lib) -> bool:
    """A library is current when the hash of its sources, recorded next to it at build time, still matches
    (content, not mtimes: the copy of the tree that travels to a GPU box does not keep them)."""
    try:
        with open(lib + ".srchash") as fh:
            return not os.path.exists(lib) or fh.read().strip() != _src_hash()
    except OSError:
        return True


def _compile(lib, extra, tag, verbose):
    """One nvcc -c per translation unit, in parallel, then one link."""
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    objs, cmds = [], []
    for src in sources():
        obj = os.path.join(OBJDIR, os.path.basename(src)[:-3] + tag + ".o")
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", "-o", obj, src]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        cmds.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r.stderr

    with ThreadPoolExecutor(max_workers=min(len(cmds), os.cpu_count() or 4)) as ex:
        logs = list(ex.map(run, cmds))
    if verbose:
        for cmd, log in zip(cmds, logs):
            print(" ".join(cmd), file=sys.stderr)
            print(log, file=sys.stderr)
    link = [nvcc, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib, *objs, "-lcuda"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    with open(lib + ".srchash", "w") as fh:
        fh.write(_src_hash())
    return lib


FRAME_E2E = os.path.join(LIBDIR, "frame_e2e")  # host program: the reference's frame loop through include/fruid.hpp


def _build_frame_e2e():
    src = os.path.join(HERE, "host", "frame_e2e.cpp")
    hpp = os.path.join(HERE, "..", "include", "fruid.hpp")
    if os.path.exists(FRAME_E2E) and all(os.path.getmtime(FRAME_E2E) >= os.path.getmtime(f) for f in (src, hpp, LIB)):
        return
    r = subprocess.run(["g++", "-std=c++17", "-O2", "-o", FRAME_E2E, src, "-L" + LIBDIR, "-lfruid_b200",
                        "-Wl,-rpath,$ORIGIN"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed:\n" + r.stdout + r.stderr)


def build(force: bool = False, verbose: bool = False, debug: bool = True) -> str:
    """Builds libfruid_b200.so, (debug=True) libfruid_b200_dbg.so -- the FRUID_DEBUG twin the ordering tests load -- and
    the host-side frame driver of the end-to-end bench."""
    if force or _stale(LIB):
        _compile(LIB, [], "", verbose)
    if debug and (force or _stale(DBG_LIB)):
        _compile(DBG_LIB, ["-DFRUID_DEBUG"], ".dbg", False)
    _build_frame_e2e()
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, debug="--no-debug" not in sys.argv))
