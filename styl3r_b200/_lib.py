"""ctypes binding of libstyl3r_b200.so (the C-ABI declared in include/styl3r_b200.h).

There is no CPU fallback: if the shared library is missing this module raises, loudly.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import os as _os

_TAG = _os.environ.get("S3R_LIB_TAG", "")  # development aid: A/B-timing of kernel variants (styl3r_b200/build.py)
_LIB_PATH = Path(__file__).resolve().parent / "lib" / (f"libstyl3r_b200_{_TAG}.so" if _TAG else "libstyl3r_b200.so")
_lib = None

S3R_OK = 0
ABI_VERSION = 3


class RasterParams(C.Structure):
    _fields_ = [
        ("n_views", C.c_int32), ("n_sets", C.c_int32), ("P", C.c_int32), ("width", C.c_int32),
        ("height", C.c_int32), ("sh_degree", C.c_int32), ("sh_coeffs", C.c_int32), ("cov_stride", C.c_int32),
        ("means3D", C.c_void_p), ("cov3D", C.c_void_p), ("shs", C.c_void_p), ("colors_precomp", C.c_void_p),
        ("opacities", C.c_void_p), ("view_set", C.c_void_p), ("viewmatrix", C.c_void_p), ("projmatrix", C.c_void_p),
        ("projmatrix_raw", C.c_void_p), ("campos", C.c_void_p), ("tanfov", C.c_void_p), ("scales", C.c_void_p),
        ("background", C.c_void_p),
    ]


class RasterOutputs(C.Structure):
    _fields_ = [("color", C.c_void_p), ("depth", C.c_void_p), ("opacity", C.c_void_p), ("radii", C.c_void_p),
                ("n_touched", C.c_void_p)]


class RasterLayout(C.Structure):
    _fields_ = [(n, C.c_int64) for n in (
        "total_bytes", "status", "counters", "depths", "xy", "conic_opacity", "rgb", "rect", "chunk_hist",
        "chunk_base", "tile_count", "ranges", "keys_unsorted", "keys_tmp", "point_list", "point_keys", "records",
        "final_T", "n_contrib", "grecords", "work_order", "blists", "bcounts", "n_contrib_blk")] + [(n, C.c_int32) for n in ("tiles_x", "tiles_y", "tiles", "chunks")]


class RasterGrads(C.Structure):
    _fields_ = [
        ("dL_dcolor", C.c_void_p), ("dL_ddepth", C.c_void_p), ("dL_dmeans3D", C.c_void_p), ("dL_dcov3D", C.c_void_p),
        ("dL_dshs", C.c_void_p), ("dL_dcolors", C.c_void_p), ("dL_dopacities", C.c_void_p),
        ("dL_dmeans2D", C.c_void_p), ("dL_dtau", C.c_void_p), ("scratch", C.c_void_p), ("scratch_bytes", C.c_size_t),
    ]


class S3RError(RuntimeError):
    pass


def lib_path() -> Path:
    return _LIB_PATH


def lib():
    """Load the CUDA library once. Raises if it has not been built (python -m styl3r_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise S3RError(
            f"{_LIB_PATH} is missing: the sm_100a CUDA library has not been built. Run "
            "`python -m styl3r_b200.build` (or __graft_entry__.build()). styl3r_b200 has no CPU/PyTorch fallback.")
    L = C.CDLL(str(_LIB_PATH))
    L.s3r_abi_version.restype = C.c_int
    L.s3r_error_string.restype = C.c_char_p
    L.s3r_error_string.argtypes = [C.c_int]
    L.s3r_raster_layout_query.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64,
                                          C.POINTER(RasterLayout)]
    L.s3r_raster_forward.argtypes = [C.POINTER(RasterParams), C.POINTER(RasterOutputs), C.c_void_p, C.c_size_t,
                                     C.c_int64, C.c_void_p]
    L.s3r_raster_forward_stages.argtypes = [C.POINTER(RasterParams), C.POINTER(RasterOutputs), C.c_void_p, C.c_size_t,
                                            C.c_int64, C.c_uint32, C.c_void_p]
    L.s3r_raster_read_status.argtypes = [C.c_void_p, C.POINTER(C.c_int64 * 4), C.c_void_p]
    L.s3r_raster_backward_scratch_bytes.restype = C.c_size_t
    L.s3r_raster_backward_scratch_bytes.argtypes = [C.c_int32, C.c_int32]
    L.s3r_raster_backward.argtypes = [C.POINTER(RasterParams), C.c_void_p, C.c_size_t, C.c_int64,
                                      C.POINTER(RasterGrads), C.c_void_p]
    L.s3r_rope2d.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int64,
                             C.c_int64, C.c_float, C.c_float, C.c_int32, C.c_void_p]
    L.s3r_camera_setup.argtypes = [C.c_void_p] * 4 + [C.c_int32, C.c_int32, C.c_int32] + [C.c_void_p] * 7
    L.s3r_gaussian_adapter.argtypes = [C.c_void_p] * 4 + [C.c_int32] * 5 + [C.c_float] + [C.c_void_p] * 7
    L.s3r_gemm_bf16.argtypes = [C.c_void_p] * 5 + [C.c_int32] * 8 + [C.c_void_p]
    L.s3r_gemm_bf16_rope.argtypes = [C.c_void_p] * 5 + [C.c_int32] * 8 + [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                                                          C.c_void_p, C.c_size_t, C.c_void_p]
    L.s3r_gemm_bf16_majors.argtypes = [C.c_void_p] * 5 + [C.c_int32] * 10 + [C.c_void_p, C.c_void_p]
    L.s3r_gemm_bf16_batched.argtypes = ([C.c_void_p] * 3 + [C.c_int32] * 6 + [C.c_int64] * 3 + [C.c_int32] * 4 + [C.c_void_p])
    L.s3r_softmax_rows_bf16.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_void_p]
    L.s3r_attention_ds_bf16.argtypes = [C.c_void_p] * 5 + [C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_void_p]
    L.s3r_conv2d_bf16.argtypes = [C.c_void_p] * 5 + [C.c_int32] * 9 + [C.c_void_p]
    L.s3r_upsample2x_nhwc_bf16.argtypes = [C.c_void_p] * 3 + [C.c_int32] * 4 + [C.c_void_p]
    L.s3r_gaussian_adapter_nhwc.argtypes = ([C.c_void_p] * 3 + [C.c_int32] * 3 + [C.c_void_p] + [C.c_int32] * 5 + [C.c_float]
                                            + [C.c_void_p] * 7)
    L.s3r_ply_pack.argtypes = [C.c_void_p] * 6 + [C.c_int32] * 3 + [C.c_void_p, C.c_void_p]
    L.s3r_rescale_crop.argtypes = ([C.c_void_p] + [C.c_int32] * 6 + [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
                                   + [C.c_int32] * 5 + [C.c_void_p] * 5)
    L.s3r_layernorm_bf16.argtypes = [C.c_void_p] * 4 + [C.c_int32, C.c_int32, C.c_int64, C.c_float, C.c_void_p]
    L.s3r_layernorm_f32_bf16.argtypes = [C.c_void_p] * 4 + [C.c_int32, C.c_int32, C.c_int64, C.c_float, C.c_void_p]
    L.s3r_layernorm_bwd_bf16.argtypes = [C.c_void_p] * 6 + [C.c_int32, C.c_int32, C.c_int64, C.c_float, C.c_void_p]
    L.s3r_set_tunable.argtypes = [C.c_int32, C.c_int32]
    L.s3r_im2col7x7_bf16.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    L.s3r_gather_other_views_bf16.argtypes = [C.c_void_p] * 3 + [C.c_int32] * 4 + [C.c_void_p]
    L.s3r_rope_table.argtypes = [C.c_void_p, C.c_int32, C.c_float, C.c_void_p]
    L.s3r_attention_bf16.argtypes = [C.c_void_p] * 4 + [C.c_int32] * 5 + [C.POINTER(C.c_int64)] * 4 + [C.c_float, C.c_void_p]
    L.s3r_se3_update_w2c.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    if L.s3r_abi_version() != ABI_VERSION:
        raise S3RError(f"ABI mismatch: library {L.s3r_abi_version()} != binding {ABI_VERSION}; rebuild")
    _lib = L
    return L


def check(code: int, what: str = "") -> None:
    if code != S3R_OK:
        msg = lib().s3r_error_string(code).decode()
        raise S3RError(f"{what or 'styl3r_b200 call'} failed: {msg} (code {code})")


EXPORTED_SYMBOLS = (
    "s3r_abi_version", "s3r_error_string", "s3r_raster_layout_query", "s3r_raster_forward", "s3r_raster_forward_stages",
    "s3r_raster_read_status", "s3r_raster_backward_scratch_bytes", "s3r_raster_backward", "s3r_rope2d",
    "s3r_se3_update_w2c", "s3r_camera_setup", "s3r_gaussian_adapter", "s3r_gemm_bf16", "s3r_gemm_bf16_rope", "s3r_gemm_bf16_majors", "s3r_gemm_bf16_batched", "s3r_softmax_rows_bf16", "s3r_attention_ds_bf16", "s3r_rope_table", "s3r_attention_bf16",
    "s3r_conv2d_bf16", "s3r_upsample2x_nhwc_bf16", "s3r_gaussian_adapter_nhwc", "s3r_set_tunable", "s3r_ply_pack", "s3r_rescale_crop", "s3r_layernorm_bf16", "s3r_layernorm_f32_bf16", "s3r_layernorm_bwd_bf16",
    "s3r_im2col7x7_bf16", "s3r_gather_other_views_bf16",
)
