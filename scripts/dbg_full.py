import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from styl3r_b200 import synthetic as syn
from oracle import raster_oracle as ro
from tests.helpers import gpu_scene
hw = int(sys.argv[1]) if len(sys.argv) > 1 else 256
sc = syn.make_scene(seed=1234, v=2, V=1, hw=hw)
cams = [ro.camera_setup(sc["extrinsics"][i], sc["intrinsics"][i], sc["near"][i], sc["far"][i], True) for i in range(1)]
out = gpu_scene(sc, cams, want_n_touched=False, check="none")
torch.cuda.synchronize()
print("ok", out[-1].status())
