"""Builds libmgb_b200.so in-tree with nvcc for sm_100a (no torch involved; plain C ABI)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("MGB_B200_LIB") or os.path.join(HERE, "libmgb_b200.so")
SOURCES = ["mgb_b200.cu", "plan_host.cpp", "launch.cu", "inst_1d.cu", "inst_2d.cu"]
HEADERS = ["kernels.cuh", "kernels_dense.cuh", "kernels_csr.cuh", "kernels_dist.cuh", "plan_host.h", "launch.h", "inst_common.cuh", os.path.join("..", "..", "include", "mgb_b200.h")]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, out: str = LIB, extra=()) -> str:
    """``out``/``extra`` build tuning variants (e.g. -DMGB_ELEM_MINBLOCKS=6) next to the default library.
    Serialised across processes by a lock file: the ranks of one job do not compile the same library at once."""
    if not force and out == LIB and not needs_build():
        return LIB
    import fcntl
    with open(out + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and out == LIB and not needs_build():   # another process built it while we waited
                return LIB
            return _build_locked(verbose, out, extra)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose: bool, out: str, extra) -> str:
    import concurrent.futures as cf
    import tempfile
    objdir = tempfile.mkdtemp(prefix="mgb_b200_obj_")
    common = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-O3", *extra]
    if verbose:
        common[1:1] = ["-Xptxas", "-v"]

    def compile_one(src):
        obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
        res = subprocess.run(common + ["-c", os.path.join(CSRC, src), "-o", obj], capture_output=True, text=True)
        return src, obj, res

    objs, logs = [], []
    with cf.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        for src, obj, res in ex.map(compile_one, SOURCES):
            if res.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}:\n" + res.stdout + res.stderr)
            objs.append(obj)
            logs.append(res.stderr)
    # link next to the target and rename into place: a process that dlopens `out` concurrently (every rank of a
    # torchrun job calls load()) sees either the old or the new complete file, never a half-written one
    tmp_out = f"{out}.tmp.{os.getpid()}"
    res = subprocess.run([_nvcc(), "--shared", "-o", tmp_out] + objs, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    os.replace(tmp_out, out)
    shutil.rmtree(objdir, ignore_errors=True)
    if verbose:
        sys.stderr.write("".join(logs))
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
