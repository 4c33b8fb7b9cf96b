"""Ad-hoc timing of the rasterizer stages (development aid, not the bench)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from styl3r_b200 import synthetic as syn, rasterizer as rz
from oracle import raster_oracle as ro
from tests.helpers import gpu_scene

def run(v, V, iters=20):
    sc = syn.make_scene(seed=1234, v=v, V=V, hw=256)
    cams = [ro.camera_setup(sc["extrinsics"][i], sc["intrinsics"][i], sc["near"][i], sc["far"][i], True) for i in range(V)]
    out = gpu_scene(sc, cams, want_n_touched=False)
    ctx = out[-1]; st = ctx.status()
    cap = ctx.capacity
    torch.cuda.synchronize()
    for _ in range(3): gpu_scene(sc, cams, want_n_touched=False, capacity=cap, check="none")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # keep inputs resident: rebuild the call with device tensors once
    import tests.helpers as h
    t0=time.time(); e0.record()
    for _ in range(iters): gpu_scene(sc, cams, want_n_touched=False, capacity=cap, check="none")
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/iters
    print(f"v={v} V={V} P={ctx.P} R={st['num_instances']} maxtile={st['max_tile_count']}: {ms*1000:.1f} us/call (incl. H2D of inputs), {V/ms*1000:.0f} views/s, wall {1e3*(time.time()-t0)/iters:.2f} ms")

if __name__ == "__main__":
    run(2, 1); run(2, 6); run(4, 6)
