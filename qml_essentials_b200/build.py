"""In-tree build of libqmlb200.so for sm_100a (nvcc cross-compiles without a GPU).

    python -m qml_essentials_b200.build [--force]

Each translation unit is compiled in its own nvcc process (the per-precision
template instantiations dominate compile time), then linked into one shared
library next to this file so it travels with the repo snapshot.
"""

from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "_obj")
LIB = os.path.join(HERE, "libqmlb200.so")
UNITS = ["qmlb_api.cu", "qmlb_reg_f32.cu", "qmlb_reg_f64.cu", "qmlb_tile_f32.cu",
         "qmlb_tile_f64.cu", "qmlb_stream_f32_lean.cu", "qmlb_stream_f32_heavy.cu", "qmlb_stream_f64_lean.cu",
         "qmlb_stream_f64_heavy.cu", "qmlb_frame_plan.cu", "qmlb_frame_f32.cu", "qmlb_frame_f64.cu",
         "qmlb_fstream_f32.cu", "qmlb_fstream_f64.cu", "qmlb_frame_ptm_f32.cu",
         "qmlb_frame_ptm_f64.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v",
]


def _sources():
    out = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    out.append(os.path.join(HERE, "..", "include", "qmlb200.h"))
    return out


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in _sources())


def _compile(unit: str) -> str:
    obj = os.path.join(OBJ, unit.replace(".cu", ".o"))
    src = os.path.join(CSRC, unit)
    deps = [s for s in _sources() if not s.endswith(".cu")] + [src]
    if os.path.exists(obj) and all(os.path.getmtime(d) <= os.path.getmtime(obj) for d in deps):
        return obj
    cmd = ["nvcc", *NVCC_FLAGS, "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(obj + ".log", "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {unit}:\n{r.stdout}\n{r.stderr}")
    return obj


def build(force: bool = False, verbose: bool = True) -> str:
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    with ThreadPoolExecutor(max_workers=min(len(UNITS), os.cpu_count() or 4)) as pool:
        objs = list(pool.map(_compile, UNITS))
    cmd = ["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(f"built {LIB}")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
