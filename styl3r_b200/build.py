"""In-tree nvcc build of libstyl3r_b200.so (sm_100a only).

`python -m styl3r_b200.build [--verbose] [--force]` or `build()`; called by __graft_entry__.build().
The library is plain CUDA runtime code behind a C-ABI (include/styl3r_b200.h): no torch headers.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
LIBDIR = ROOT / "lib"
# development aid: S3R_BUILD_TAG=x builds lib/libstyl3r_b200_x.so (objects under build/obj_x) with S3R_NVCC_FLAGS, so that
# kernel variants can be A/B-timed in one GPU session (S3R_LIB_TAG=x selects it in _lib.py); the product is the untagged one
_TAG = os.environ.get("S3R_BUILD_TAG", "")
LIB = LIBDIR / (f"libstyl3r_b200_{_TAG}.so" if _TAG else "libstyl3r_b200.so")
OBJDIR = ROOT.parent / "build" / (f"obj_{_TAG}" if _TAG else "obj")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
          "-I", str(ROOT.parent / "include")]
# per-file extra flags: the preprocess stage must not contract a*b+c (bit-exact tile rects / depth keys)
EXTRA = {"raster_preprocess.cu": ["-fmad=false"]}


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def sources():
    return sorted(CSRC.glob("*.cu"))


def _stale(out: Path, deps) -> bool:
    if not out.exists():
        return True
    t = out.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    srcs = sources()
    hdrs = list(CSRC.glob("*.cuh")) + list((ROOT.parent / "include").glob("*.h"))
    OBJDIR.mkdir(parents=True, exist_ok=True)
    LIBDIR.mkdir(parents=True, exist_ok=True)
    cc = nvcc()
    env = dict(os.environ)
    # /opt/gcc wrappers in this image miss libgomp specs etc.; use the system host compiler
    ccbin = ["-ccbin", "/usr/bin/g++"] if Path("/usr/bin/g++").exists() else []

    def compile_one(src: Path):
        obj = OBJDIR / (src.stem + ".o")
        if not force and not _stale(obj, [src, *hdrs, Path(__file__)]):
            return obj, ""
        cmd = [cc, *ccbin, *ARCH, *COMMON, *EXTRA.get(src.name, []), *os.environ.get("S3R_NVCC_FLAGS", "").split(),
               "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd += ["-Xptxas", "-v"]
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(compile_one, srcs))
    objs = [o for o, _ in results]
    if verbose:
        for (_, log), s in zip(results, srcs):
            if log:
                print(f"--- {s.name}\n{log}")
    if force or _stale(LIB, objs):
        cmd = [cc, *ccbin, *ARCH, "-shared", "-o", str(LIB), *map(str, objs), "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(p)
