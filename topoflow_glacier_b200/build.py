"""In-tree nvcc build of libtfglacier.so (sm_100a only).

``python -m topoflow_glacier_b200.build`` or ``__graft_entry__.build()``.  The shared library is
written next to the sources (``topoflow_glacier_b200/lib/``) so that it travels with a snapshot of
the repository; nothing is JIT-compiled at import time and nothing is cached outside the tree.
"""

from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "lib"
LIB = LIBDIR / "libtfglacier.so"
INCLUDE = PKG.parent / "include"

SOURCES = ["tfg_abi.cu", "tfg_run_strict.cu", "tfg_run_fast.cu", "tfg_run_f32.cu"]
# float32 kernel: flush-to-zero, so that MUFU.EX2 / LG2 / RCP need no denormal pre- and post-scaling (three extra
# instructions around each of the ~25 special-function calls of a step); float64 code is unaffected by the flag
# fast float64 kernel: no implicit contraction -- every fused multiply-add is written out (fmadd / fma), so all template
# instantiations (recording, aggregates, TMA staging) evaluate exactly the same arithmetic and agree bit for bit
EXTRA_FLAGS = {"tfg_run_f32.cu": ["-ftz=true", "-prec-sqrt=false", "-prec-div=false"], "tfg_run_fast.cu": ["-fmad=false"]}
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def source_digest() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(INCLUDE.glob("*.h"))):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update((" ".join(NVCC_FLAGS) + repr(sorted(EXTRA_FLAGS.items()))).encode())
    return h.hexdigest()


def is_current() -> bool:
    stamp = LIBDIR / "build.sha256"
    return LIB.exists() and stamp.exists() and stamp.read_text().strip() == source_digest()


def build(force: bool = False, verbose: bool = False, defines=(), out: Path | None = None, extra=()) -> Path:
    """Compile every CUDA source for sm_100a and link libtfglacier.so; returns its path.

    ``defines`` / ``out`` build an experimental variant next to the stock library (tuning sweeps).
    """
    variant = bool(defines) or out is not None
    if not force and not variant and is_current():
        return LIB
    nvcc = _nvcc()
    LIBDIR.mkdir(exist_ok=True)
    target = Path(out) if out else LIB
    objdir = LIBDIR / ("obj_" + target.stem if variant else "obj")
    objdir.mkdir(exist_ok=True)

    def compile_one(src: str) -> Path:
        obj = objdir / (Path(src).stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *EXTRA_FLAGS.get(src, []), *[f"-D{d}" for d in defines], *extra, "-I", str(INCLUDE), "-c",
               str(CSRC / src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(target), *map(str, objs)]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if not variant:
        (LIBDIR / "build.sha256").write_text(source_digest() + "\n")
    return target


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--out=")]
    extra = [a for a in sys.argv[1:] if a.startswith("-X")]   # e.g. -Xptxas=--register-usage-level=7 (tuning sweeps)
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs, out=outs[0] if outs else None, extra=extra))
