"""Static-shape serving path of the renderer: one CUDA graph per session that does

    pinned host inputs --H2D--> camera set-up kernel --> rasterizer kernel chain --D2H--> pinned host image

so that a request costs one graph launch on the host (the eager `render_cuda` call spends ~150-250 us of Python per
request, more than the GPU work of a 256x256 view).  Same math as `render_cuda` (reference signature
src/model/decoder/cuda_splatting.py:46-61); the inputs are *bound* pinned host tensors which the caller refills
between launches (zero staging copies).

`host["covariances"]` may be the reference's [S,G,3,3] matrices or the packed upper triangle [S,G,6]
(xx, xy, xz, yy, yz, zz) - the layout the reference itself hands to the rasterizer (`cov3D_precomp`,
cuda_splatting.py:118,126): 24 instead of 36 bytes per Gaussian over PCIe.

    sess = RenderSession(host)            # host: dict of pinned tensors, see `KEYS`
    sess.run()                            # enqueue on the current stream; sess.color_host holds the image afterwards
    torch.cuda.current_stream().synchronize(); sess.check()
"""
from __future__ import annotations

from math import isqrt
from typing import Dict, Optional

import torch

from .. import _lib
from .. import rasterizer as _rz
from .cuda_splatting import _sh_layout, camera_setup

KEYS = ("extrinsics", "intrinsics", "near", "far", "background", "means", "covariances", "harmonics", "opacities")


class RenderSession:
    def __init__(self, host: Dict[str, torch.Tensor], image_shape, view_set: Optional[torch.Tensor] = None,
                 scale_invariant: bool = True, capacity: Optional[int] = None, device="cuda", want_depth: bool = False):
        for k in KEYS:
            if k not in host:
                raise KeyError(f"RenderSession needs host['{k}']")
            if not host[k].is_pinned():
                raise _lib.S3RError(f"host['{k}'] must be pinned host memory (tensor.pin_memory())")
        self.host = host
        self.device = torch.device(device)
        self.h, self.w = image_shape
        self.scale_invariant = scale_invariant
        V = host["extrinsics"].shape[0]
        S, P = host["means"].shape[:2]
        self.dev = {k: torch.empty_like(host[k], device=self.device) for k in KEYS}
        self.view_set = (view_set.to(self.device, torch.int32) if view_set is not None
                         else torch.arange(V, device=self.device, dtype=torch.int32) % S)
        self.color_host = torch.empty(V, 3, self.h, self.w).pin_memory()
        self.depth_host = torch.empty(V, self.h, self.w).pin_memory() if want_depth else None
        self.status_host = torch.zeros(4, dtype=torch.int64).pin_memory()
        self.degree = min(isqrt(host["harmonics"].shape[-1]) - 1, 3)
        self._plan = None
        self._graph = None
        # capacity: measured once with a synchronous probe render of the current host contents (+25 % head-room)
        self._copy_in()
        if capacity is None:
            plan = self._make_plan(max(4 * P * V, 1 << 16))
            plan.launch()
            st = plan.ctx.status()
            capacity = max(int(st["num_instances"] * 1.25) + 4096, 1 << 16)
            if st["overflow"]:
                capacity = int(st["num_instances"] * 1.25) + 4096
        self.capacity = int(capacity)
        self._capture()

    def _copy_in(self):
        for k in KEYS:
            self.dev[k].copy_(self.host[k], non_blocking=True)

    def _make_plan(self, cap):
        d = self.dev
        view_t, full, proj_t, campos, tan_fov, scale = camera_setup(d["extrinsics"], d["intrinsics"], d["near"], d["far"],
                                                                    self.scale_invariant)
        S, P = d["means"].shape[:2]
        V = d["extrinsics"].shape[0]
        shs = _sh_layout(d["harmonics"])
        tensors = (d["means"], d["covariances"], d["opacities"], shs, None, view_t, full, proj_t, campos, tan_fov,
                   scale if self.scale_invariant else None, d["background"], self.view_set)
        cov_stride = 6 if d["covariances"].dim() == 3 else 9
        return _rz.RasterPlan(tensors, S, P, V, self.w, self.h, shs.shape[2], self.degree, cov_stride, cap)

    def _capture(self):
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            self._copy_in()
            self._make_plan(self.capacity).launch()  # warm-up outside capture
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._copy_in()
            self._plan = self._make_plan(self.capacity)
            self._plan.launch()
            self.color_host.copy_(self._plan.color, non_blocking=True)
            if self.depth_host is not None:
                self.depth_host.copy_(self._plan.depth, non_blocking=True)
            self.status_host.copy_(self._plan.ctx.view("status"), non_blocking=True)

    def run(self) -> torch.Tensor:
        """Enqueue one request on the current stream (asynchronous). Returns the pinned image buffer."""
        self._graph.replay()
        return self.color_host

    def check(self) -> None:
        """Call after synchronising: raises if the last request overflowed the session's instance capacity."""
        if int(self.status_host[1]) != 0:
            raise _lib.S3RError(f"RenderSession capacity {self.capacity} overflowed (needed {int(self.status_host[0])}); "
                                "create the session with a larger `capacity`")
