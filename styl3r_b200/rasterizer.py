"""Host side of the B200 rasterizer: torch tensors in, C-ABI calls (ctypes) underneath.

`rasterize(...)` is the batched, differentiable entry used by `styl3r_b200.decoder.render_cuda`; it replaces
the per-view loop around the third-party `GaussianRasterizer` in the reference
(src/model/decoder/cuda_splatting.py:93-133): one launch chain renders all views, without `.item()` syncs
and without materialising per-view copies of the Gaussians (a `view_set` index maps views to Gaussian sets).

PyTorch only provides device memory, the current stream and autograd plumbing here; there is no PyTorch/CPU
fallback for the computation.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib

_DT = {"u8": torch.uint8, "i32": torch.int32, "f32": torch.float32, "i64": torch.int64, "u16": torch.int16}


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


_layout_cache: dict = {}


def query_layout(n_views: int, P: int, W: int, H: int, capacity: int) -> _lib.RasterLayout:
    key = (n_views, P, W, H, capacity)
    lay = _layout_cache.get(key)
    if lay is None:
        lay = _lib.RasterLayout()
        _lib.check(_lib.lib().s3r_raster_layout_query(n_views, P, W, H, capacity, lay), "s3r_raster_layout_query")
        if len(_layout_cache) > 256:
            _layout_cache.clear()
        _layout_cache[key] = lay
    return lay


@dataclass
class RasterContext:
    """Everything the backward pass (and the parity tests) need from one forward call."""
    params: _lib.RasterParams
    layout: _lib.RasterLayout
    state: torch.Tensor
    capacity: int
    keep: tuple  # tensors referenced by raw pointers in `params`
    n_views: int
    n_sets: int
    P: int
    W: int
    H: int

    def view(self, name: str) -> torch.Tensor:
        """Typed view of one internal array of the state buffer (tile assignment, sort order, ...)."""
        L, nv, P, cap = self.layout, self.n_views, self.P, self.capacity
        HW = self.W * self.H
        spec = {
            "status": ("i64", 4), "depths": ("f32", nv * P), "xy": ("f32", nv * P * 2),
            "conic_opacity": ("f32", nv * P * 4), "rgb": ("f32", nv * P * 4), "rect": ("i32", nv * P),
            "tile_count": ("i32", nv * L.tiles), "ranges": ("i32", nv * L.tiles * 2),
            "keys_unsorted": ("i64", cap), "point_list": ("i32", cap), "point_keys": ("i64", cap),
            "records": ("f32", cap * 12), "final_T": ("f32", nv * HW), "n_contrib": ("i32", nv * HW),
            "chunk_hist": ("u16", nv * L.chunks * L.tiles), "chunk_base": ("i32", nv * L.chunks * L.tiles),
            "work_order": ("i32", nv * L.tiles), "blists": ("i32", cap * 8), "bcounts": ("i32", nv * L.tiles * 8),
            "n_contrib_blk": ("i32", nv * HW),
        }[name]
        dt = _DT[spec[0]]
        off = getattr(L, name)
        nbytes = spec[1] * torch.empty((), dtype=dt).element_size()
        return self.state[off:off + nbytes].view(dt)

    def status(self):
        out = (C.c_int64 * 4)()
        _lib.check(_lib.lib().s3r_raster_read_status(_ptr(self.state), out,
                                                    C.c_void_p(torch.cuda.current_stream(self.state.device).cuda_stream)),
                   "s3r_raster_read_status")
        return dict(num_instances=int(out[0]), overflow=bool(out[1]), max_tile_count=int(out[2]))


_capacity_hint: dict = {}
_pending: list = []  # deferred overflow checks: (event, pinned status copy, shape key)
_pinned_pool: list = []  # recycled pinned int64[4] buffers (cudaHostAlloc per call would cost ~100 us)


def _pinned_status() -> torch.Tensor:
    return _pinned_pool.pop() if _pinned_pool else torch.empty(4, dtype=torch.int64).pin_memory()



def validate_pending(block: bool = False) -> None:
    """Resolve deferred overflow checks (check="deferred").  Raises S3RError if an earlier asynchronous render
    dropped instances because its capacity was too small (the capacity hint is raised so a retry succeeds)."""
    bad = None
    keep = []
    for ev, host, key in _pending:
        if block:
            ev.synchronize()
        if ev.query():
            r_total, overflow = int(host[0]), bool(host[1])
            _capacity_hint[key] = max(_capacity_hint.get(key, 0), int(r_total * 1.5) + 1024, 1 << 16)
            if overflow:
                bad = (key, r_total)
            _pinned_pool.append(host)
        else:
            keep.append((ev, host, key))
    _pending[:] = keep
    if bad is not None:
        raise _lib.S3RError(f"a deferred-check render (S,P,V,W,H)={bad[0]} overflowed its instance capacity "
                            f"(needed {bad[1]}); its image is incomplete — re-render (capacity hint updated)")



def forward_raw(means, cov, opacities, viewmatrix, projmatrix, tanfov, background, W: int, H: int, *, shs=None,
                colors_precomp=None, sh_degree: int = 0, campos=None, projmatrix_raw=None, scales=None,
                view_set=None, want_n_touched: bool = False, capacity: Optional[int] = None, check: str = "sync"):
    """Non-differentiable forward. Shapes: means [S,P,3]; cov [S,P,6] or [S,P,3,3]; opacities [S,P];
    shs [S,P,M,3] or colors_precomp [S,P,3]; viewmatrix/projmatrix(/_raw) [V,4,4] in the reference's
    transposed layout; tanfov [V,2]; background [V,3]; campos [V,3]; scales [V]; view_set [V] int32.

    check: "sync"     — read the status word after the launches (one 32-byte D2H per *batch*) and transparently
                        re-run with a larger instance capacity on overflow;
           "deferred" — asynchronous once a capacity hint exists for this shape (first call syncs): the status word
                        is copied to pinned memory and inspected by a later call / validate_pending(), which raises
                        if the render had overflowed (capacity = 1.5x the largest R seen, so this is rare);
           "none"     — fully asynchronous (CUDA-graph friendly); the caller inspects ctx.status() later.
    Returns (color[V,3,H,W], depth[V,H,W], opacity[V,H,W], radii[V,P], n_touched[V,P] or None, ctx)."""
    L = _lib.lib()
    dev = means.device
    if dev.type != "cuda":
        raise _lib.S3RError("styl3r_b200 rasterizer needs CUDA tensors (no CPU fallback)")
    means, opacities = _f32c(means), _f32c(opacities)
    cov = _f32c(cov)
    S, P = means.shape[0], means.shape[1]
    cov_stride = 9 if cov.dim() == 4 else cov.shape[-1]
    viewmatrix, projmatrix = _f32c(viewmatrix), _f32c(projmatrix)
    V = viewmatrix.shape[0]
    tanfov, background = _f32c(tanfov), _f32c(background)
    shs = _f32c(shs) if shs is not None else None
    colors_precomp = _f32c(colors_precomp) if colors_precomp is not None else None
    campos = _f32c(campos) if campos is not None else None
    projmatrix_raw = _f32c(projmatrix_raw) if projmatrix_raw is not None else None
    scales = _f32c(scales) if scales is not None else None
    if view_set is not None:
        view_set = view_set.to(device=dev, dtype=torch.int32).contiguous()
    M = shs.shape[2] if shs is not None else 1

    key = (S, P, V, W, H)
    if _pending:
        validate_pending()
    if check == "deferred" and capacity is None and key not in _capacity_hint:
        check = "sync"  # establish a capacity for this shape first
    cap = int(capacity) if capacity is not None else _capacity_hint.get(key, max(4 * P * V, 1 << 16))
    tensors = (means, cov, opacities, shs, colors_precomp, viewmatrix, projmatrix, projmatrix_raw, campos, tanfov,
               scales, background, view_set)
    while True:
        plan = RasterPlan(tensors, S, P, V, W, H, M, sh_degree, cov_stride, cap, want_n_touched)
        plan.launch()
        ctx = plan.ctx
        if check == "none":
            break
        if check == "deferred":
            host = _pinned_status()
            host.copy_(ctx.view("status"), non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            _pending.append((ev, host, key))
            break
        st = ctx.status()
        if not st["overflow"]:
            # remember a comfortable capacity for this shape (next call allocates it up front)
            _capacity_hint[key] = max(int(st["num_instances"] * 1.5) + 1024, 1 << 16)
            break
        cap = int(st["num_instances"] * 1.25) + 1024
    return plan.color, plan.depth, plan.opacity, plan.radii, plan.n_touched, ctx


STAGE_PREPROCESS, STAGE_BIN, STAGE_SORT, STAGE_BLEND, STAGE_ALL = 1, 2, 4, 8, 15


class RasterPlan:
    """Pre-allocated state + outputs + parameter block for one rasterization shape; `launch()` enqueues the
    kernel chain (or the stages selected by `stage_mask`) on the current stream and can be called repeatedly or
    captured in a CUDA graph."""

    def __init__(self, tensors, S, P, V, W, H, M, sh_degree, cov_stride, cap, want_n_touched=False):
        (means, cov, opacities, shs, colors_precomp, viewmatrix, projmatrix, projmatrix_raw, campos, tanfov, scales,
         background, view_set) = tensors
        dev = means.device
        self.dev = dev
        lay = query_layout(V, P, W, H, cap)
        self.state = torch.empty(lay.total_bytes, dtype=torch.uint8, device=dev)
        self.color = torch.empty(V, 3, H, W, dtype=torch.float32, device=dev)
        self.depth = torch.empty(V, H, W, dtype=torch.float32, device=dev)
        self.opacity = torch.empty(V, H, W, dtype=torch.float32, device=dev)
        self.radii = torch.empty(V, P, dtype=torch.int32, device=dev)
        self.n_touched = torch.zeros(V, P, dtype=torch.int32, device=dev) if want_n_touched else None
        self.prm = _lib.RasterParams(
            n_views=V, n_sets=S, P=P, width=W, height=H, sh_degree=sh_degree, sh_coeffs=M, cov_stride=cov_stride,
            means3D=_ptr(means), cov3D=_ptr(cov), shs=_ptr(shs), colors_precomp=_ptr(colors_precomp),
            opacities=_ptr(opacities), view_set=_ptr(view_set), viewmatrix=_ptr(viewmatrix),
            projmatrix=_ptr(projmatrix), projmatrix_raw=_ptr(projmatrix_raw), campos=_ptr(campos),
            tanfov=_ptr(tanfov), scales=_ptr(scales), background=_ptr(background))
        self.out = _lib.RasterOutputs(color=_ptr(self.color), depth=_ptr(self.depth), opacity=_ptr(self.opacity),
                                      radii=_ptr(self.radii), n_touched=_ptr(self.n_touched))
        self.cap = cap
        self.ctx = RasterContext(self.prm, lay, self.state, cap, tensors, V, S, P, W, H)

    def launch(self, stage_mask: int = STAGE_ALL):
        with torch.cuda.device(self.dev):  # the C entry points launch on the CURRENT device: make it the tensors' device
            stream = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
            _lib.check(_lib.lib().s3r_raster_forward_stages(self.prm, self.out, _ptr(self.state),
                                                            self.ctx.layout.total_bytes, self.cap, stage_mask, stream),
                       "s3r_raster_forward")


def backward_raw(ctx: RasterContext, dL_dcolor, dL_ddepth=None, *, need_pose: bool = True, only_pose: bool = False,
                 needs: Optional[dict] = None):
    """Gradients w.r.t. the tensors given to forward_raw. Returns a dict of tensors shaped like the inputs
    (`cov` in the packing it was given), plus dL_dmeans2D [V,P,3] and dL_dtau [V,6] = (rho, theta).
    only_pose: the pose-align loop needs dL/dtau only — no Gaussian gradient buffers are allocated or written.
    needs: {"means", "cov", "opacities", "shs", "colors", "means2D"} -> bool (default all True): gradients that are not
    needed are neither allocated, zero-filled nor accumulated (stage-2 training needs dL/dSH only: the structure heads
    are frozen, model_wrapper_style.py:854-868)."""
    L = _lib.lib()
    prm = ctx.params
    dev = ctx.state.device
    V, S, P = ctx.n_views, ctx.n_sets, ctx.P
    M, cs = prm.sh_coeffs, prm.cov_stride
    dL_dcolor = _f32c(dL_dcolor)
    dL_ddepth = _f32c(dL_ddepth) if dL_ddepth is not None else None
    needs = needs or {}
    want = lambda k: not only_pose and needs.get(k, True)
    z = lambda k, *shape: torch.zeros(*shape, device=dev) if want(k) else None
    g = dict(
        means=z("means", S, P, 3), cov=z("cov", S, P, cs), opacities=z("opacities", S, P), means2D=z("means2D", V, P, 3),
        tau=torch.zeros(V, 6, device=dev),
        shs=z("shs", S, P, M, 3) if prm.shs else None,
        colors=z("colors", S, P, 3) if prm.colors_precomp else None,
    )
    nbytes = L.s3r_raster_backward_scratch_bytes(V, P)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    grads = _lib.RasterGrads(
        dL_dcolor=_ptr(dL_dcolor), dL_ddepth=_ptr(dL_ddepth), dL_dmeans3D=_ptr(g["means"]), dL_dcov3D=_ptr(g["cov"]),
        dL_dshs=_ptr(g["shs"]), dL_dcolors=_ptr(g["colors"]), dL_dopacities=_ptr(g["opacities"]),
        dL_dmeans2D=_ptr(g["means2D"]), dL_dtau=_ptr(g["tau"]) if need_pose else None, scratch=_ptr(scratch),
        scratch_bytes=nbytes)
    with torch.cuda.device(dev):
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(L.s3r_raster_backward(prm, _ptr(ctx.state), ctx.layout.total_bytes, ctx.capacity, grads, stream),
                   "s3r_raster_backward")
    return g


class _Rasterize(torch.autograd.Function):
    """Differentiable batched rasterization. Camera tensors are treated as constants except for the pose
    perturbation (cam_trans_delta -> rho, cam_rot_delta -> theta) whose gradient dL/dtau the kernel produces,
    matching the `theta=` / `rho=` arguments of the reference call (cuda_splatting.py:127-128).  `means2D` is
    upstream's screen-space gradient holder: its value is ignored, its .grad receives dL/d(NDC mean)."""

    @staticmethod
    def forward(ctx, means, cov, opacities, shs, colors_precomp, rho, theta, means2D, cfg):
        color, depth, opacity, radii, n_touched, rctx = forward_raw(
            means, cov, opacities, cfg["viewmatrix"], cfg["projmatrix"], cfg["tanfov"], cfg["background"], cfg["W"],
            cfg["H"], shs=shs, colors_precomp=colors_precomp, sh_degree=cfg["sh_degree"], campos=cfg["campos"],
            projmatrix_raw=cfg["projmatrix_raw"], scales=cfg.get("scales"), view_set=cfg.get("view_set"),
            want_n_touched=cfg.get("want_n_touched", False), capacity=cfg.get("capacity"),
            check=cfg.get("check", "sync"))
        ctx.rctx = rctx
        # the kernels reach the inputs through raw pointers (RasterContext.keep), which bypasses autograd's version
        # counters: remember the versions and refuse to differentiate through tensors that were modified in place since
        ctx.versions = [(t, t._version) for t in (means, cov, opacities, shs, colors_precomp) if t is not None]
        ctx.cov_shape = cov.shape
        ctx.m2d_shape = None if means2D is None else means2D.shape
        ctx.has = (shs is not None, colors_precomp is not None, rho is not None, theta is not None)
        if n_touched is None:
            n_touched = torch.empty(0, dtype=torch.int32, device=means.device)
        ctx.mark_non_differentiable(opacity, radii, n_touched)
        return color, depth, opacity, radii, n_touched

    @staticmethod
    def backward(ctx, g_color, g_depth, *_):
        rctx = ctx.rctx
        for t, ver in ctx.versions:
            if t._version != ver:
                raise RuntimeError("one of the tensors given to the rasterizer (means / covariances / opacities / SH) was "
                                   "modified in place between forward and backward; its gradient would be computed from "
                                   "the new values")
        has_sh, has_col, has_rho, has_theta = ctx.has
        ni = ctx.needs_input_grad  # (means, cov, opacities, shs, colors_precomp, rho, theta, means2D, cfg)
        needs = dict(means=ni[0], cov=ni[1], opacities=ni[2], shs=ni[3], colors=ni[4],
                     means2D=ctx.m2d_shape is not None and ni[7])
        g = backward_raw(rctx, g_color, g_depth, need_pose=(has_rho and ni[5]) or (has_theta and ni[6]), needs=needs)
        g_m2d = None if g["means2D"] is None else g["means2D"].reshape(ctx.m2d_shape)
        return (g["means"], None if g["cov"] is None else g["cov"].reshape(ctx.cov_shape), g["opacities"], g["shs"] if has_sh else None,
                g["colors"] if has_col else None, g["tau"][:, :3] if has_rho else None,
                g["tau"][:, 3:] if has_theta else None, g_m2d, None)


def rasterize(means, cov, opacities, *, shs=None, colors_precomp=None, rho=None, theta=None, means2D=None, **cfg):
    """Differentiable batched rasterization; see forward_raw for shapes.
    Returns (color, depth, opacity, radii, n_touched)."""
    return _Rasterize.apply(means, cov, opacities, shs, colors_precomp, rho, theta, means2D, cfg)
