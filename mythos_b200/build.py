"""In-tree build of libmythos_b200.so (nvcc, sm_100a only).

``python -m mythos_b200.build`` or ``__graft_entry__.build()``.  The shared object lands next to this file
(mythos_b200/libmythos_b200.so); it is git-ignored but travels to the GPU box with the tree.
"""

from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libmythos_b200.so"
SOURCES = ["abi_common.cu", "energy_kernels.cu", "frame_kernels.cu", "list_kernels.cu", "neighbors.cu", "langevin.cu", "difftre.cu", "peaks.cu", "theta_tape.cu", "observables.cu", "trajectory.cu"]
NVCC_FLAGS = [
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-O3",
    "-std=c++17",
    "-lineinfo",
    "-Xcompiler",
    "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (Path(cand).exists() or cand == "nvcc"):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "mythos_b200.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False, out: Path | None = None, defines: tuple[str, ...] = ()) -> Path:
    """Build the library.  ``out`` / ``defines`` build a kernel-variant copy (experiments: load it through the
    MYTHOS_B200_LIB environment variable); the default build is the product."""
    lib = Path(out) if out else LIB
    if out is None and not force and not needs_build():
        return LIB
    objs = []
    obj_dir = PKG / "csrc" / ("_obj" if out is None else "_obj_" + lib.stem)
    obj_dir.mkdir(exist_ok=True)
    procs = []
    for src in SOURCES:
        path = CSRC / src
        if not path.exists():
            continue
        obj = obj_dir / (path.stem + ".o")
        cmd = [_nvcc(), *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-c", str(path), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(obj))
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            print(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
    cmd = [_nvcc(), "-shared", "-o", str(lib), *objs, "-lcudart"]
    subprocess.run(cmd, check=True)
    return lib


if __name__ == "__main__":
    _out = next((a.split("=", 1)[1] for a in sys.argv if a.startswith("--out=")), None)
    _defs = tuple(a[2:] for a in sys.argv if a.startswith("-D"))
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, out=_out, defines=_defs))
