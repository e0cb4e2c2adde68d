"""In-tree build of the C-ABI CUDA library (sm_100a only).

    python -m imagine360_b200.build [--force] [--verbose]

Each ``csrc/*.cu`` is compiled to an object in ``csrc/_obj`` (skipped when up to date) and linked
into ``imagine360_b200/libimagine360_b200.so``.  nvcc cross-compiles without a GPU, so this runs
on the CPU-only build box; the resulting ``.so`` travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OBJ = CSRC / "_obj"
LIB = HERE / "libimagine360_b200.so"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v" if os.environ.get("I360_PTXAS_V") else "-O3",
]


def _newer(src: Path, dst: Path, deps) -> bool:
    if not dst.exists():
        return True
    t = dst.stat().st_mtime
    return any(p.stat().st_mtime > t for p in [src, *deps])


def build(force: bool = False, verbose: bool = False) -> Path:
    OBJ.mkdir(exist_ok=True)
    headers = sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + sorted((HERE.parent / "include").glob("*.h"))
    sources = sorted(CSRC.glob("*.cu"))
    jobs = []
    for src in sources:
        obj = OBJ / (src.stem + ".o")
        if force or _newer(src, obj, headers):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [NVCC, *FLAGS, "-I", str(HERE.parent / "include"), "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, r

    failed = False
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for src, r in ex.map(compile_one, jobs):
            if verbose or r.returncode != 0:
                sys.stderr.write(f"[build] {src.name}\n{r.stdout}{r.stderr}\n")
            if r.returncode != 0:
                failed = True
    if failed:
        raise RuntimeError("nvcc failed; see messages above")
    objs = [OBJ / (s.stem + ".o") for s in sources]
    if force or jobs or not LIB.exists():
        cmd = [NVCC, "-shared", "-o", str(LIB), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}{r.stderr}")
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
