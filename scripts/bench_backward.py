"""Forward + backward timing of the rasterizer on cfg2 / 6-view shapes (development aid)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from styl3r_b200 import synthetic as syn, rasterizer as rz
from oracle import raster_oracle as ro
from tests.helpers import gpu_scene

def run(v, V):
    sc = syn.make_scene(seed=1234, v=v, V=V, hw=256)
    cams = [ro.camera_setup(sc["extrinsics"][i], sc["intrinsics"][i], sc["near"][i], sc["far"][i], True) for i in range(V)]
    color, depth, opacity, radii, nt, ctx = gpu_scene(sc, cams, want_n_touched=False)
    gc = torch.randn_like(color); gd = torch.randn_like(depth) * 0.1
    def t(fn, it=20):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(it): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / it * 1000
    full = t(lambda: rz.backward_raw(ctx, gc, gd))
    pose = t(lambda: rz.backward_raw(ctx, gc, None, only_pose=True))
    only_sh = dict(means=False, cov=False, opacities=False, shs=True, colors=False, means2D=False)
    sh = t(lambda: rz.backward_raw(ctx, gc, None, need_pose=False, needs=only_sh))
    print(f"v={v} V={V} P={ctx.P} R={ctx.status()['num_instances']}: backward all grads {full:.0f} us ({full/V:.0f}/view), pose-only {pose:.0f} us ({pose/V:.0f}/view), SH-only {sh:.0f} us ({sh/V:.0f}/view)", flush=True)
run(2, 1); run(2, 6)
