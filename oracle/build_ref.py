"""Compile the reference's own C++ sources for the hot path where they lie under /root/reference into
oracle/_ref/ (git-ignored, travels to the GPU box).  TEST INFRASTRUCTURE ONLY.

Only one native file of the reference is buildable here: curope.cpp (the CPU RoPE-2D, `rope_2d_cpu`); its CUDA
sibling kernels.cu does not compile against torch 2.11 (`tokens.type()` dispatch, kernels.cu:101) and is replaced
by a one-line stub so that the extension links.  The rasterizer is third-party and absent (see raster_oracle.c)."""
from __future__ import annotations

import importlib.util
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF_SRC = Path("/root/reference/src/model/encoder/backbone/croco/curope/curope.cpp")
OUT = HERE / "_ref"
NAME = "curope_ref"


def built_path():
    cands = sorted(OUT.glob(f"{NAME}*.so"))
    return cands[0] if cands else None


def build(force: bool = False):
    if built_path() is not None and not force:
        return built_path()
    if not REF_SRC.exists():
        raise RuntimeError("reference sources not present (only the prebuilt oracle/_ref travels to the GPU box)")
    from torch.utils import cpp_extension
    OUT.mkdir(exist_ok=True)
    stub = OUT / "rope_2d_cuda_stub.cpp"
    stub.write_text('#include <torch/extension.h>\n'
                    'void rope_2d_cuda(torch::Tensor, const torch::Tensor, const float, const float) {\n'
                    '  TORCH_CHECK(false, "reference CPU build: rope_2d_cuda is not available");\n}\n')
    cpp_extension.load(name=NAME, sources=[str(REF_SRC), str(stub)], build_directory=str(OUT),
                       extra_cflags=["-O2"], verbose=False)
    return built_path()


def load():
    """Import the compiled reference module (needs torch). Returns None if it was never built."""
    p = built_path()
    if p is None:
        return None
    import torch  # noqa: F401  (the extension links against libtorch)
    spec = importlib.util.spec_from_file_location(NAME, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build())
