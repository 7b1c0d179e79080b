"""In-tree build of liblsk.so (hand-written sm_100a kernels + C ABI + C++ host layer).

nvcc cross-compiles for sm_100a without a GPU, so this runs in the authoring container and the
built library travels to the B200 box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
HOST = PKG / "host"
INCLUDE = PKG.parent / "include"
LIB_DIR = PKG / "lib"
LIB_PATH = LIB_DIR / "liblsk.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: liblsk.so cannot be built (there is no CPU fallback)")


def sources() -> list[Path]:
    srcs = sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cpp"))
    if HOST.exists():
        srcs += sorted(HOST.glob("*.cu")) + sorted(HOST.glob("*.cpp"))
    return srcs


def _deps() -> list[Path]:
    deps = sources() + sorted(CSRC.glob("*.cuh")) + sorted(INCLUDE.glob("*.h"))
    if HOST.exists():
        deps += sorted(HOST.glob("*.hpp")) + sorted(HOST.glob("*.h"))
    return deps


def is_stale() -> bool:
    if not LIB_PATH.exists():
        return True
    t = LIB_PATH.stat().st_mtime
    return any(d.stat().st_mtime > t for d in _deps())


def build_library(force: bool = False, verbose: bool = False, defines: list[str] | None = None, out: Path | None = None) -> Path:
    """defines / out: developer experiments (extra -D flags into an alternative .so)."""
    if out is None and not force and not is_stale():
        return LIB_PATH
    LIB_DIR.mkdir(exist_ok=True)
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    host_cxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"
    cmd = [_nvcc(), *NVCC_FLAGS, "-ccbin", host_cxx, "-I", str(INCLUDE), "-I", str(CSRC), "-I", str(HOST),
           "-shared", "-o", str(out or LIB_PATH), *[f"-D{d}" for d in (defines or [])], *map(str, sources()), "-lnccl", "-lcudart"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    proc = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stderr)
    return out or LIB_PATH


if __name__ == "__main__":
    import sys

    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
