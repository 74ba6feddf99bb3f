"""Build recipes: the C-ABI host library and one sm_100a cubin per fractal module.

Role in the reference: src/main/cuda/compile.sh:10-15 / compile_all.sh (one PTX per fractal,
``-arch=sm_30``) plus Maven for the host.  Here: ``nvcc -cubin`` for sm_100a only, ``g++`` for
the driver-API host library.  Everything is built in-tree so the artefacts travel with the
repository snapshot to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
KERNELS_DIR = PKG_DIR / "cudaKernels"
LIB_DIR = PKG_DIR / "lib"
LIB_PATH = LIB_DIR / "libchaos_ultra.so"

ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-lineinfo", "-std=c++17", "-Xptxas", "-v"]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; the fractal modules cannot be built")
    return exe


def _cuda_include() -> str:
    return str(Path(_nvcc()).resolve().parent.parent / "include")


BUILD_LOG = []      # (action, artefact): what the last build_all() rebuilt and what it found up to date


def _note(action: str, out: Path):
    BUILD_LOG.append((action, str(out.relative_to(PKG_DIR.parent)) if str(out).startswith(str(PKG_DIR.parent)) else str(out)))
    if os.environ.get("CHAOS_BUILD_QUIET") != "1":
        sys.stderr.write("[build] %-7s %s\n" % (action, BUILD_LOG[-1][1]))


def _newer(target: Path, sources) -> bool:
    if not target.exists():
        return False
    t = target.stat().st_mtime
    return all(Path(s).stat().st_mtime <= t for s in sources)


def _run(cmd, log=None):
    res = subprocess.run(cmd, capture_output=True, text=True)
    if log is not None:
        Path(log).write_text(res.stdout + res.stderr)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("build step failed: " + " ".join(map(str, cmd)))
    return res.stdout + res.stderr


def module_sources():
    return sorted((CSRC / "fractals").glob("*.cu"))


def build_module(src: Path, force: bool = False) -> Path:
    """fractals/<name>.cu (+ render_generic.cuh) -> cudaKernels/<name>.cubin"""
    KERNELS_DIR.mkdir(exist_ok=True)
    out = KERNELS_DIR / (src.stem + ".cubin")
    deps = [src] + sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + [Path(__file__)]
    if not force and _newer(out, deps):
        _note("reused", out)
        return out
    _note("rebuilt", out)
    extra = os.environ.get("CHAOS_NVCC_EXTRA", "").split()    # experiments only, e.g. -DCHAOS_REFILL_SLOTS=6
    cmd = [_nvcc(), "-cubin", *ARCH_FLAGS, *NVCC_FLAGS, *extra, "-I", str(CSRC), str(src), "-o", str(out)]
    _run(cmd, log=KERNELS_DIR / (src.stem + ".ptxas.log"))
    return out


COMPAT_DIR = PKG_DIR / "cudaKernels_compat"
REFERENCE_FRACTALS = Path("/root/reference/src/main/cuda/fractals")
# reference modules whose build keeps c.y's multiply and subtract apart (see compat/chaos_compat_post.cuh)
UNFUSED_PLANE_Y = {"mandelbrot", "julia"}


def build_compat_module(src: Path, out_dir: Path = COMPAT_DIR, fused_plane_y: bool = True, force: bool = False) -> Path:
    """A module file written for the REFERENCE's contract (fractal.cuh:7-28: computeFractal / colorize / debugFractal),
    unmodified, -> <out_dir>/<name>.cubin for this backend.  The translation unit is the counterpart of the one
    compile.sh:10-15 writes: compat prologue, the author's file (by path, nothing is copied), compat epilogue."""
    src = Path(src).resolve()
    out_dir.mkdir(exist_ok=True)
    out = out_dir / (src.stem + ".cubin")
    deps = [src] + sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + sorted((CSRC / "compat").glob("*.cuh")) + [Path(__file__)]
    if not force and _newer(out, deps):
        _note("reused", out)
        return out
    _note("rebuilt", out)
    wrap = out_dir / ("tmp_compiling_%s.cu" % src.stem)
    wrap.write_text('#include "compat/chaos_compat_pre.cuh"\n#include "%s"\n#include "compat/chaos_compat_post.cuh"\n' % src)
    try:
        # the author's directory first (its own headers), then the compat headers (fractal.cuh / helpers.cuh)
        cmd = [_nvcc(), "-cubin", *ARCH_FLAGS, *NVCC_FLAGS, "-w", "-DCHAOS_COMPAT_FUSED_PLANE_Y=%d" % int(fused_plane_y),
               "-I", str(CSRC), "-I", str(CSRC / "compat"), "-I", str(src.parent.parent), str(wrap), "-o", str(out)]
        _run(cmd, log=out_dir / (src.stem + ".ptxas.log"))
    finally:
        wrap.unlink(missing_ok=True)
    return out


def build_compat_reference(force: bool = False):
    """every module of the reference, from its own unmodified source where it lies (only where /root/reference is
    mounted: this container; the GPU box gets the built files)"""
    if not REFERENCE_FRACTALS.is_dir():
        return []
    return [build_compat_module(s, fused_plane_y=s.stem not in UNFUSED_PLANE_Y, force=force) for s in sorted(REFERENCE_FRACTALS.glob("*.cu"))]


def build_library(force: bool = False) -> Path:
    LIB_DIR.mkdir(exist_ok=True)
    srcs = [CSRC / "chaos_abi.cpp", CSRC / "chaos_driver.cpp"]
    deps = srcs + sorted(CSRC.glob("*.h")) + [PKG_DIR.parent / "include" / "chaos_ultra.h", Path(__file__)]
    if not force and _newer(LIB_PATH, deps):
        _note("reused", LIB_PATH)
        return LIB_PATH
    _note("rebuilt", LIB_PATH)
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wextra", "-fvisibility=hidden",
           "-I", _cuda_include(), *map(str, srcs), "-o", str(LIB_PATH), "-ldl", "-lm"]
    _run(cmd)
    return LIB_PATH


def build_bench_kernels(force: bool = False) -> Path:
    """bench_kernels/peak.cu -> bench_kernels/peak.cubin (measurement only: FP64/FP32 FMA peak)"""
    src = PKG_DIR.parent / "bench_kernels" / "peak.cu"
    out = src.with_suffix(".cubin")
    if not force and _newer(out, [src]):
        _note("reused", out)
        return out
    _note("rebuilt", out)
    _run([_nvcc(), "-cubin", *ARCH_FLAGS, "-O3", "-lineinfo", "-std=c++17", str(src), "-o", str(out)])
    return out


def build_all(force: bool = False):
    """Everything, in-tree.  An artefact is reused only if it is newer than every source it depends on (all headers of csrc/
    and this script included); what was rebuilt and what was reused is logged line by line (stderr) and kept in BUILD_LOG.
    CHAOS_BUILD_FORCE=1 rebuilds regardless."""
    force = force or os.environ.get("CHAOS_BUILD_FORCE") == "1"
    del BUILD_LOG[:]
    lib = build_library(force)
    mods = [build_module(s, force) for s in module_sources()]
    build_bench_kernels(force)
    build_compat_reference(force)
    return lib, mods


if __name__ == "__main__":
    if sys.argv[1:2] == ["compat"]:      # python build.py compat <module.cu> [out_dir] [--unfused-plane-y]
        args = [a for a in sys.argv[2:] if not a.startswith("--")]
        print(build_compat_module(Path(args[0]), Path(args[1]) if len(args) > 1 else COMPAT_DIR,
                                  fused_plane_y="--unfused-plane-y" not in sys.argv, force=True))
        sys.exit(0)
    lib, mods = build_all(force="--force" in sys.argv)
    print(lib)
    for m in mods:
        print(m)
