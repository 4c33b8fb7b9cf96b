"""ncu target: one launch of each blend-backward variant (all / geometry only / colour only) on the bench scene (cfg2).
  ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:s3r_blend_bwd -o gpurun_out/r02b_bwd python scripts/ncu_backward.py"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from oracle import raster_oracle as ro
from styl3r_b200 import rasterizer as rz
from styl3r_b200 import synthetic as syn
from tests.helpers import gpu_scene

sc = syn.make_scene(seed=1234, v=2, V=1, hw=256)
cams = [ro.camera_setup(sc["extrinsics"][i], sc["intrinsics"][i], sc["near"][i], sc["far"][i], True) for i in range(1)]
color, depth, opacity, radii, nt, ctx = gpu_scene(sc, cams, want_n_touched=False)
gc = torch.randn_like(color)
gd = torch.randn_like(depth) * 0.1
only_sh = dict(means=False, cov=False, opacities=False, shs=True, colors=False, means2D=False)
runs = [lambda: rz.backward_raw(ctx, gc, gd), lambda: rz.backward_raw(ctx, gc, None, only_pose=True),
        lambda: rz.backward_raw(ctx, gc, None, need_pose=False, needs=only_sh)]
for r in runs:
    for _ in range(3):
        r()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for r in runs:
    r()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
