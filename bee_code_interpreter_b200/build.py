"""In-tree build of libbci.so (sm_100a only) and of the oracle's C restatement.

`python -m bee_code_interpreter_b200.build` or `__graft_entry__.build()` runs it.  nvcc
cross-compiles without a GPU; the resulting .so files are git-ignored but travel to the
GPU box with the working tree.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
REPO_ROOT = PKG_DIR.parent
CSRC = PKG_DIR / "csrc"
BUILD_DIR = REPO_ROOT / "build" / "bci"
LIB_PATH = PKG_DIR / "libbci.so"

ORACLE_DIR = REPO_ROOT / "oracle"
ORACLE_LIB = ORACLE_DIR / "liboracle.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-Wno-format-truncation",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: libbci.so cannot be built")


def _newer(target: Path, deps: list[Path]) -> bool:
    if not target.exists():
        return False
    t = target.stat().st_mtime
    return all(d.stat().st_mtime <= t for d in deps)


def build_lib(force: bool = False, verbose: bool = False) -> Path:
    """Compiles every csrc/*.cu for sm_100a and links libbci.so next to the package."""
    sources = sorted(CSRC.glob("*.cu"))
    headers = sorted(CSRC.glob("*.cuh")) + [REPO_ROOT / "include" / "bci.h"]
    if not sources:
        raise RuntimeError(f"no CUDA sources under {CSRC}")
    if not force and _newer(LIB_PATH, sources + headers):
        return LIB_PATH
    BUILD_DIR.mkdir(parents=True, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src: Path) -> Path:
        obj = BUILD_DIR / (src.stem + ".o")
        if not force and _newer(obj, [src] + headers):
            return obj
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        objs = list(pool.map(compile_one, sources))
    tmp = LIB_PATH.with_suffix(".so.tmp")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(tmp), *map(str, objs), "-cudart", "static", "-Xlinker", "--no-undefined", "-ldl", "-lpthread", "-lrt"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link of libbci.so failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


def build_oracle(force: bool = False) -> Path:
    """Compiles oracle/*.c (CPU restatement used only by tests/bench baselines) with gcc."""
    sources = sorted(ORACLE_DIR.glob("*.c"))
    if not sources:
        raise RuntimeError(f"no C sources under {ORACLE_DIR}")
    if not force and _newer(ORACLE_LIB, sources):
        return ORACLE_LIB
    gcc = shutil.which("gcc") or "gcc"
    tmp = ORACLE_LIB.with_suffix(".so.tmp")
    cmd = [gcc, "-O2", "-fPIC", "-shared", "-std=c11", "-o", str(tmp), *map(str, sources), "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"gcc failed for the oracle:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, ORACLE_LIB)
    return ORACLE_LIB


SERVER_SRC = CSRC / "bci_executor_server.cpp"
SERVER_BIN = PKG_DIR / "bci-executor-server"


def build_server(force: bool = False) -> Path:
    """Compiles the native executor server (the server.rs stand-in; plain C++17, no CUDA)."""
    if not force and _newer(SERVER_BIN, [SERVER_SRC]):
        return SERVER_BIN
    gxx = shutil.which("g++") or "g++"
    tmp = SERVER_BIN.with_suffix(".tmp")
    cmd = [gxx, "-O2", "-std=c++17", "-Wall", "-Wextra", "-o", str(tmp), str(SERVER_SRC), "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"g++ failed for the executor server:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, SERVER_BIN)
    return SERVER_BIN


def main() -> None:
    force = "--force" in sys.argv
    verbose = "-v" in sys.argv
    print(build_lib(force=force, verbose=verbose))
    print(build_server(force=force))
    print(build_oracle(force=force))


if __name__ == "__main__":
    main()
