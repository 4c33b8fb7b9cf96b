"""Drop-in module shims: importing names the reference imports from third-party/native extensions.

`install()` registers `diff_gaussian_rasterization`, `curope` and `xformers` / `xformers.ops` in sys.modules so that
the unmodified reference files src/model/decoder/cuda_splatting.py:5-8, …/croco/curope/curope2d.py:6-9 and
…/croco/blocks.py:25 import the B200 implementations.
"""
import sys


def install(rasterizer: bool = True, rope: bool = True, attention: bool = True) -> None:
    if rasterizer:
        from . import diff_gaussian_rasterization as dgr
        sys.modules.setdefault("diff_gaussian_rasterization", dgr)
    if rope:
        from .. import curope as _curope
        sys.modules.setdefault("curope", _curope)
    if attention:
        from . import xformers as _xf
        sys.modules.setdefault("xformers", _xf)
        sys.modules.setdefault("xformers.ops", _xf.ops)
