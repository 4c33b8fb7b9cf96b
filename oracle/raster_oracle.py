"""ctypes/numpy front-end of the CPU oracle (oracle/raster_oracle.c, oracle/rope_oracle.c).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / ``--impl reference`` legs.  Never imported by styl3r_b200/.

PARITY UNPINNED for the rasterizer (third-party, un-vendored, unpinned upstream — see the header
of raster_oracle.c).  The camera set-up below restates the reference call site
``src/model/decoder/cuda_splatting.py:16-43,65-88`` and ``src/geometry/projection.py:247-261``
in numpy fp32, one rounding per operation like the torch ops it mirrors.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "liboracle.so"
_lib = None

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int32)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
u8p = C.POINTER(C.c_uint8)
i64p = C.POINTER(C.c_int64)


def build(force: bool = False) -> Path:
    """Compile the C oracle with the system gcc (OpenMP if available)."""
    srcs = [_HERE / "raster_oracle.c", _HERE / "rope_oracle.c"]
    if not force and _LIB_PATH.exists() and all(_LIB_PATH.stat().st_mtime >= s.stat().st_mtime for s in srcs):
        return _LIB_PATH
    _LIB_PATH.parent.mkdir(exist_ok=True)
    base = ["-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math"]
    last = None
    for cc in ("/usr/bin/gcc", "gcc", "cc"):
        for omp in (["-fopenmp"], []):
            cmd = [cc, *base, *omp, "-o", str(_LIB_PATH), *map(str, srcs), "-lm"]
            try:
                subprocess.run(cmd, check=True, capture_output=True)
                return _LIB_PATH
            except (subprocess.CalledProcessError, FileNotFoundError) as e:  # try next variant
                last = e
    raise RuntimeError(f"could not build the oracle: {last}")


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_LIB_PATH))
        _lib.s3r_oracle_bin_sort.restype = C.c_int64
        _lib.s3r_oracle_num_threads.restype = C.c_int
    return _lib


def num_threads() -> int:
    return int(lib().s3r_oracle_num_threads())


def set_num_threads(n: int) -> None:
    """OpenMP team size of the oracle (torchrun exports OMP_NUM_THREADS=1; the reference arm wants every host core)."""
    lib().s3r_oracle_set_num_threads(int(n))


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


# --------------------------------------------------------------------------- rasterizer


def forward(means, cov6, opac, view, proj, campos, W, H, tanx, tany, bg, shs=None, colors=None, deg=0,
            render=True):
    """One view, inputs exactly as the reference hands them to GaussianRasterizer (already scaled).

    view/proj: 16 floats m[4*col+row].  shs: [P,M,3].  Returns a dict with every intermediate."""
    L = lib()
    means, cov6, opac = _f32(means), _f32(cov6), _f32(opac).reshape(-1)
    view, proj, campos, bg = _f32(view).reshape(16), _f32(proj).reshape(16), _f32(campos).reshape(3), _f32(bg)
    P = means.shape[0]
    M = 1
    if shs is not None:
        shs = _f32(shs)
        M = shs.shape[1]
    if colors is not None:
        colors = _f32(colors)
    o = dict(
        radii=np.zeros(P, np.int32), xy=np.zeros((P, 2), np.float32), depths=np.zeros(P, np.float32),
        conic_opacity=np.zeros((P, 4), np.float32), rgb=np.zeros((P, 3), np.float32),
        clamped=np.zeros((P, 3), np.uint8), tiles_touched=np.zeros(P, np.uint32), rects=np.zeros((P, 4), np.int32),
    )
    L.s3r_oracle_preprocess(
        C.c_int(P), C.c_int(deg), C.c_int(M), _p(means, f32p), _p(cov6, f32p), _p(shs, f32p), _p(colors, f32p),
        _p(opac, f32p), _p(view, f32p), _p(proj, f32p), _p(campos, f32p), C.c_int(W), C.c_int(H),
        C.c_float(tanx), C.c_float(tany), _p(o["radii"], i32p), _p(o["xy"], f32p), _p(o["depths"], f32p),
        _p(o["conic_opacity"], f32p), _p(o["rgb"], f32p), _p(o["clamped"], u8p), _p(o["tiles_touched"], u32p),
        _p(o["rects"], i32p))
    gx, gy = (W + 15) // 16, (H + 15) // 16
    R = int(o["tiles_touched"].astype(np.int64).sum())
    o.update(
        R=R, keys_unsorted=np.zeros(R, np.uint64), vals_unsorted=np.zeros(R, np.uint32),
        keys=np.zeros(R, np.uint64), point_list=np.zeros(R, np.uint32), ranges=np.zeros((gx * gy, 2), np.uint32),
    )
    L.s3r_oracle_bin_sort(
        C.c_int(P), C.c_int(W), C.c_int(H), _p(o["radii"], i32p), _p(o["depths"], f32p),
        _p(o["tiles_touched"], u32p), _p(o["rects"], i32p), _p(o["keys_unsorted"], u64p),
        _p(o["vals_unsorted"], u32p), _p(o["keys"], u64p), _p(o["point_list"], u32p), _p(o["ranges"], u32p))
    o.update(inputs=dict(means=means, cov6=cov6, opac=opac, view=view, proj=proj, campos=campos, W=W, H=H,
                         tanx=float(tanx), tany=float(tany), bg=bg, shs=shs, colors=colors, deg=deg, M=M))
    if render:
        o.update(
            color=np.zeros((3, H, W), np.float32), depth=np.zeros((H, W), np.float32),
            opacity=np.zeros((H, W), np.float32), final_T=np.zeros((H, W), np.float32),
            n_contrib=np.zeros((H, W), np.uint32), n_touched=np.zeros(P, np.int32), sens=np.zeros((H, W), np.uint32),
        )
        L.s3r_oracle_render(
            C.c_int(W), C.c_int(H), _p(o["ranges"], u32p), _p(o["point_list"], u32p), _p(o["xy"], f32p),
            _p(o["rgb"], f32p), _p(o["depths"], f32p), _p(o["conic_opacity"], f32p), _p(bg, f32p),
            _p(o["color"], f32p), _p(o["depth"], f32p), _p(o["opacity"], f32p), _p(o["final_T"], f32p),
            _p(o["n_contrib"], u32p), _p(o["n_touched"], i32p), _p(o["sens"], u32p))
    return o


def backward(fwd, dL_dcolor, dL_ddepth=None, proj_raw=None):
    """Gradients for one view given forward() output. proj_raw: 16 floats (projection only)."""
    L = lib()
    i = fwd["inputs"]
    P, W, H, M = i["means"].shape[0], i["W"], i["H"], i["M"]
    dL_dcolor = _f32(dL_dcolor)
    dL_ddepth = _f32(dL_ddepth) if dL_ddepth is not None else None
    g = dict(
        dL_dmean2D=np.zeros((P, 2), np.float32), dL_dconic=np.zeros((P, 3), np.float32),
        dL_dopacity=np.zeros(P, np.float32), dL_dcolor=np.zeros((P, 3), np.float32),
        dL_ddepthg=np.zeros(P, np.float32),
    )
    L.s3r_oracle_render_backward(
        C.c_int(W), C.c_int(H), _p(fwd["ranges"], u32p), _p(fwd["point_list"], u32p), _p(fwd["xy"], f32p),
        _p(fwd["rgb"], f32p), _p(fwd["depths"], f32p), _p(fwd["conic_opacity"], f32p), _p(i["bg"], f32p),
        _p(fwd["final_T"], f32p), _p(fwd["n_contrib"], u32p), _p(dL_dcolor, f32p), _p(dL_ddepth, f32p),
        _p(g["dL_dmean2D"], f32p), _p(g["dL_dconic"], f32p), _p(g["dL_dopacity"], f32p), _p(g["dL_dcolor"], f32p),
        _p(g["dL_ddepthg"], f32p))
    proj_raw = _f32(proj_raw).reshape(16)
    use_sh = i["shs"] is not None
    g.update(
        dL_dmeans=np.zeros((P, 3), np.float32), dL_dcov6=np.zeros((P, 6), np.float32),
        dL_dsh=np.zeros((P, M, 3), np.float32) if use_sh else None, dL_dtau_pg=np.zeros((P, 6), np.float32),
    )
    L.s3r_oracle_preprocess_backward(
        C.c_int(P), C.c_int(i["deg"]), C.c_int(M), _p(i["means"], f32p), _p(i["cov6"], f32p), _p(i["shs"], f32p),
        C.c_int(1 if use_sh else 0), _p(i["view"], f32p), _p(proj_raw, f32p), _p(i["campos"], f32p), C.c_int(W),
        C.c_int(H), C.c_float(i["tanx"]), C.c_float(i["tany"]), _p(fwd["radii"], i32p), _p(fwd["clamped"], u8p),
        _p(g["dL_dmean2D"], f32p), _p(g["dL_dconic"], f32p), _p(g["dL_dcolor"], f32p), _p(g["dL_ddepthg"], f32p),
        _p(g["dL_dmeans"], f32p), _p(g["dL_dcov6"], f32p), _p(g["dL_dsh"], f32p), _p(g["dL_dtau_pg"], f32p))
    g["dL_dtau"] = g["dL_dtau_pg"].astype(np.float64).sum(0).astype(np.float32)
    return g


# --------------------------------------------------------------------------- camera set-up
# Restates cuda_splatting.py:16-43 (projection), :65-72 (scale invariance), :81-88 (fov / view / proj)
# and projection.py:247-261 (get_fov) in fp32 numpy.


def get_fov(K):
    K = np.asarray(K, np.float32)
    Kinv = np.linalg.inv(K.astype(np.float64)).astype(np.float32)

    def ray(v):
        d = (Kinv @ np.asarray(v, np.float32)).astype(np.float32)
        return d / np.float32(np.sqrt(np.float32((d * d).sum())))

    l, r, t, b = ray([0, 0.5, 1]), ray([1, 0.5, 1]), ray([0.5, 0, 1]), ray([0.5, 1, 1])
    return np.float32(np.arccos(np.float32((l * r).sum()))), np.float32(np.arccos(np.float32((t * b).sum())))


def projection_matrix(near, far, fov_x, fov_y):
    near, far = np.float32(near), np.float32(far)
    tx, ty = np.float32(np.tan(np.float32(0.5) * fov_x)), np.float32(np.tan(np.float32(0.5) * fov_y))
    top, right = ty * near, tx * near
    bottom, left = -top, -right
    m = np.zeros((4, 4), np.float32)
    m[0, 0] = np.float32(2) * near / (right - left)
    m[1, 1] = np.float32(2) * near / (top - bottom)
    m[0, 2] = (right + left) / (right - left)
    m[1, 2] = (top + bottom) / (top - bottom)
    m[3, 2] = 1
    m[2, 2] = far / (far - near)
    m[2, 3] = -(far * near) / (far - near)
    return m


def camera_setup(extrinsics, intrinsics, near, far, scale_invariant=True):
    """Per-view camera quantities as the reference builds them before the rasterizer call.

    Returns dict(view16, proj16, projraw16, campos, tanx, tany, scale) — matrices flattened in the
    reference's transposed (m[4*col+row]) layout."""
    e = np.array(extrinsics, np.float32)
    near, far = np.float32(near), np.float32(far)
    scale = np.float32(1.0)
    if scale_invariant:
        scale = np.float32(1.0) / near
        e = e.copy()
        e[:3, 3] = e[:3, 3] * scale
        near, far = near * scale, far * scale
    fov_x, fov_y = get_fov(intrinsics)
    tanx, tany = np.float32(np.tan(np.float32(0.5) * fov_x)), np.float32(np.tan(np.float32(0.5) * fov_y))
    proj = projection_matrix(near, far, fov_x, fov_y).T.copy()
    view = np.linalg.inv(e.astype(np.float64)).astype(np.float32).T.copy()
    full = (view @ proj).astype(np.float32)
    return dict(view16=view.reshape(16), proj16=full.reshape(16), projraw16=proj.reshape(16),
                campos=e[:3, 3].copy(), tanx=tanx, tany=tany, scale=scale)


def scale_gaussians(means, covs, scale):
    """cuda_splatting.py:69-70: cov * scale**2, mean * scale (fp32, one rounding each)."""
    s = np.float32(scale)
    s2 = np.float32(s * s)
    return (_f32(means) * s).astype(np.float32), (_f32(covs) * s2).astype(np.float32)


def cov3x3_to_6(cov):
    cov = _f32(cov)
    r, c = np.triu_indices(3)
    return np.ascontiguousarray(cov[..., r, c])


# --------------------------------------------------------------------------- RoPE


def rope2d(tokens, positions, base=100.0, fwd=1.0):
    """tokens [B,N,H,D] fp32 (copied), positions [B,N,2] int64 -> rotated copy."""
    t = np.array(tokens, dtype=np.float32, order="C", copy=True)
    pos = np.ascontiguousarray(np.asarray(positions, dtype=np.int64))
    B, N, H, D = t.shape
    lib().s3r_oracle_rope2d(_p(t, f32p), _p(pos, i64p), C.c_int(B), C.c_int(N), C.c_int(H), C.c_int(D),
                            C.c_float(base), C.c_float(fwd))
    return t
