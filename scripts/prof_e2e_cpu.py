import sys, cProfile, pstats
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch, numpy as np
from styl3r_b200 import synthetic as syn, rasterizer as rz
from styl3r_b200.decoder import cuda_splatting as cs
sc = syn.make_scene(seed=1, v=2, V=1, hw=256)
dev = "cuda"
pin = lambda a: torch.as_tensor(np.ascontiguousarray(a)).pin_memory()
h = dict(means=pin(sc["means"][None]), cov=pin(sc["covariances"][None]), sh=pin(sc["harmonics"][None]), opac=pin(sc["opacities"][None]),
         extr=pin(sc["extrinsics"]), intr=pin(sc["intrinsics"]), near=pin(sc["near"]), far=pin(sc["far"]), bg=torch.zeros(1, 3).pin_memory())
vs0 = torch.zeros(1, dtype=torch.int32, device=dev)
out = torch.empty(1, 3, 256, 256).pin_memory()
def step():
    d = {k: v.to(dev, non_blocking=True) for k, v in h.items()}
    color, _ = cs.render_cuda(d["extr"], d["intr"], d["near"], d["far"], (256, 256), d["bg"], d["means"], d["cov"], d["sh"], d["opac"],
                              scale_invariant=True, view_set=vs0, check="deferred")
    out.copy_(color, non_blocking=True)
with torch.no_grad():
    for _ in range(20): step()
    torch.cuda.synchronize()
    pr = cProfile.Profile(); pr.enable()
    for _ in range(200): step()
    pr.disable(); torch.cuda.synchronize()
st = pstats.Stats(pr); st.sort_stats("cumulative").print_stats(22)
