"""ncu target: one launch of every raster forward kernel on the bench scene (cfg2) inside a cudaProfilerStart/Stop
window, followed by the blend kernel on the heaviest tile alone (S3R_TUNE_BLEND_ONLY_TILE).  Run as
  ncu --set full --import-source on --clock-control none --profile-from-start off -o gpurun_out/r02_raster python scripts/ncu_raster.py"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch

from scripts.raster_probe import make_plan
from styl3r_b200 import _lib
from styl3r_b200 import rasterizer as rz

_lib.lib().s3r_set_tunable(5, int(sys.argv[1]) if len(sys.argv) > 1 else 0)
plan = make_plan(1234)
for _ in range(3):
    plan.launch()
torch.cuda.synchronize()
rg = plan.ctx.view("ranges").cpu().numpy().reshape(-1, 2)
heavy = int(np.argmax(rg[:, 1] - rg[:, 0]))
torch.cuda.profiler.start()
plan.launch()
torch.cuda.synchronize()
_lib.lib().s3r_set_tunable(8, heavy + 1)
plan.launch(rz.STAGE_BLEND)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
_lib.lib().s3r_set_tunable(8, 0)
print("done, heaviest tile", heavy)
