"""Drop-in module shims: importing names the reference imports from third-party/native extensions.

`install()` registers `diff_gaussian_rasterization` and `curope` in sys.modules so that the unmodified
reference files src/model/decoder/cuda_splatting.py:5-8 and …/croco/curope/curope2d.py:6-9 import the B200
implementations.
"""
import sys


def install(rasterizer: bool = True, rope: bool = True) -> None:
    if rasterizer:
        from . import diff_gaussian_rasterization as dgr
        sys.modules.setdefault("diff_gaussian_rasterization", dgr)
    if rope:
        from .. import curope as _curope
        sys.modules.setdefault("curope", _curope)
