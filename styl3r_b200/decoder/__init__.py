"""Renderer half of the hot path: drop-in for the reference's `src/model/decoder` package surface
(src/model/decoder/__init__.py:11-12, decoder_splatting_cuda.py:22-68, cuda_splatting.py:46-227)."""
from .cuda_splatting import (DepthRenderingMode, get_fov, get_projection_matrix, render_cuda,  # noqa: F401
                             render_cuda_orthographic)
from .decoder_splatting_cuda import (DecoderOutput, DecoderSplattingCUDA, DecoderSplattingCUDACfg,  # noqa: F401
                                     get_decoder)
from .session import RenderSession  # noqa: F401,E402
