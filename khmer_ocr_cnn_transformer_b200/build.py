"""Build libkocr_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libkocr_b200.so"
SOURCES = ["kocr_api.cu", "gemm_tc.cu", "dec_fused.cu", "preprocess.cu", "cnn_misc.cu", "seq.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "kocr.h"]
    return any(d.exists() and d.stat().st_mtime > t for d in deps)     # (an installed copy may lack ../include)


def build(force: bool = False, verbose: bool = False) -> Path:
    """Serialised across processes (N ranks of a fresh checkout all call this): an exclusive flock on build/.lock, the
    library is linked under a temporary name and renamed into place, so nobody dlopens a half-written file."""
    if not force and not needs_build():
        return LIB
    import fcntl
    build_dir = PKG / "build"
    build_dir.mkdir(exist_ok=True)
    with open(build_dir / ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():         # another process built it while we waited
                return LIB
            return _build_locked(build_dir, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(build_dir: Path, verbose: bool) -> Path:
    objs = []
    procs = []
    for src in SOURCES:
        obj = build_dir / (src + ".o")
        cmd = [_nvcc(), *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(obj))
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"==== {src}\n{out}")
        if p.returncode != 0:
            sys.stderr.write("\n".join(log))
            raise RuntimeError(f"nvcc failed on {src}")
    (build_dir / "ptxas.log").write_text("\n".join(log))
    tmp = LIB.with_name(LIB.name + f".tmp{os.getpid()}")
    cmd = [_nvcc(), "-shared", "-o", str(tmp), *objs, "-lcudart"]
    subprocess.run(cmd, check=True)
    os.replace(tmp, LIB)
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
