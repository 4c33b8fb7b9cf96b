"""`.ply` export of a Gaussian scene - mirror of src/model/ply_export.py:12-74:

    export_ply(means[G,3], scales[G,3], rotations[G,4] xyzw, harmonics[G,3,d_sh], opacities[G], path,
               shift_and_scale=False, save_sh_dc_only=True)

writes the standard 3DGS vertex layout (x y z, nx ny nz, f_dc_*, [f_rest_*], opacity, scale_* (log), rot_* (wxyz)) as
a binary little-endian PLY - byte-compatible with what the reference produces through `plyfile`.  The row packing
(incl. the scipy quaternion round trip of ply_export.py:46-49) runs in one CUDA kernel (`s3r_ply_pack`); the host only
computes the optional median / 0.95-quantile normalisation with torch ops on the device, copies the packed rows back
once and writes header + bytes.  CUDA only (no CPU fallback)."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import torch
from torch import Tensor

from . import _lib


def construct_list_of_attributes(num_rest: int) -> list:
    """ply_export.py:12-23."""
    attributes = ["x", "y", "z", "nx", "ny", "nz"]
    attributes += [f"f_dc_{i}" for i in range(3)]
    attributes += [f"f_rest_{i}" for i in range(num_rest)]
    attributes.append("opacity")
    attributes += [f"scale_{i}" for i in range(3)]
    attributes += [f"rot_{i}" for i in range(4)]
    return attributes


def pack_vertices(means: Tensor, scales: Tensor, rotations: Tensor, harmonics: Tensor, opacities: Tensor,
                  shift_and_scale: bool = False, save_sh_dc_only: bool = True) -> Tensor:
    """The [G, 17 (+ 3*(d_sh-1))] float32 vertex rows, on the device."""
    if means.device.type != "cuda":
        raise _lib.S3RError("styl3r_b200.ply_export needs CUDA tensors (no CPU fallback)")
    f32 = lambda t: t.detach().to(torch.float32).contiguous()
    means, scales, rotations, harmonics, opacities = map(f32, (means, scales, rotations, harmonics, opacities))
    n, d_sh = means.shape[0], harmonics.shape[-1]
    if scales.shape != (n, 3) or rotations.shape != (n, 4) or harmonics.shape != (n, 3, d_sh) or opacities.shape != (n,):
        raise _lib.S3RError("export_ply: inconsistent shapes")
    xform = None
    if shift_and_scale:  # ply_export.py:36-43
        median = means.median(dim=0).values
        factor = (means - median).abs().quantile(0.95, dim=0).max()
        xform = torch.cat((median, factor[None])).contiguous()
    n_rest = 0 if save_sh_dc_only else 3 * (d_sh - 1)
    out = torch.empty((n, 17 + n_rest), dtype=torch.float32, device=means.device)
    p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
    _lib.check(_lib.lib().s3r_ply_pack(p(means), p(scales), p(rotations), p(harmonics), p(opacities), p(xform), n, d_sh,
                                       0 if save_sh_dc_only else 1, p(out),
                                       C.c_void_p(torch.cuda.current_stream(means.device).cuda_stream)), "s3r_ply_pack")
    return out


def export_ply(means: Tensor, scales: Tensor, rotations: Tensor, harmonics: Tensor, opacities: Tensor, path: Path,
               shift_and_scale: bool = False, save_sh_dc_only: bool = True) -> None:
    rows = pack_vertices(means, scales, rotations, harmonics, opacities, shift_and_scale, save_sh_dc_only)
    names = construct_list_of_attributes(rows.shape[1] - 17)
    header = "ply\nformat binary_little_endian 1.0\n" + f"element vertex {rows.shape[0]}\n" + \
        "".join(f"property float {a}\n" for a in names) + "end_header\n"
    path = Path(path)
    path.parent.mkdir(exist_ok=True, parents=True)
    host = rows.cpu().numpy().astype("<f4", copy=False)
    with open(path, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(host.tobytes())


def read_ply(path: Path):
    """Minimal reader of the files written above (tests / round trips): (names, float32 rows [G, n_attr])."""
    import numpy as np
    data = Path(path).read_bytes()
    end = data.index(b"end_header\n") + len(b"end_header\n")
    lines = data[:end].decode("ascii").splitlines()
    if lines[0] != "ply" or lines[1] != "format binary_little_endian 1.0":
        raise ValueError("not a binary little-endian PLY")
    n = int(next(l for l in lines if l.startswith("element vertex")).split()[-1])
    names = [l.split()[-1] for l in lines if l.startswith("property float")]
    return names, np.frombuffer(data[end:], dtype="<f4").reshape(n, len(names))
