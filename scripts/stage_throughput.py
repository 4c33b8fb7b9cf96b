"""Development aid (GPU): aggregate throughput of each raster stage alone when replayed as CUDA graphs on 8 streams
over 16 resident scenes (the bench's pipelined mode) - i.e. each stage's share of the SM time at the `value` operating
point.  State from a full chain run stays valid, so any single stage can be replayed in isolation."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import bench
from styl3r_b200 import rasterizer as rz
from styl3r_b200.decoder import cuda_splatting as cs

dev = torch.device("cuda", 0)
HW, V = 256, 1
n_slots, n_streams = 16, int(sys.argv[1]) if len(sys.argv) > 1 else 8
slots = []
for s in range(n_slots):
    sc = bench.make_scene(1234 + s)
    t = lambda a: torch.as_tensor(a, device=dev)
    g = dict(means=t(sc["means"])[None], cov=t(sc["covariances"])[None], sh=t(sc["harmonics"])[None],
             opac=t(sc["opacities"])[None], extr=t(sc["extrinsics"]), intr=t(sc["intrinsics"]), near=t(sc["near"]), far=t(sc["far"]))
    scale = 1 / g["near"]
    extr = g["extr"].clone(); extr[:, :3, 3] = extr[:, :3, 3] * scale[:, None]
    fov = cs.get_fov(g["intr"])
    proj_t = cs.get_projection_matrix(g["near"] * scale, g["far"] * scale, fov[:, 0], fov[:, 1]).transpose(1, 2).contiguous()
    view_t = extr.inverse().transpose(1, 2).contiguous()
    full = (view_t @ proj_t).contiguous()
    tensors = (g["means"], g["cov"], g["opac"], g["sh"].reshape(1, -1, 1, 3), None, view_t, full, proj_t,
               extr[:, :3, 3].contiguous(), (0.5 * fov).tan().contiguous(), scale.contiguous(),
               torch.zeros(V, 3, device=dev), torch.zeros(V, dtype=torch.int32, device=dev))
    P = g["means"].shape[1]
    plan = rz.RasterPlan(tensors, 1, P, V, HW, HW, 1, 0, 9, 3 * P)
    plan.launch()
    slots.append(plan)
torch.cuda.synchronize()
streams = [torch.cuda.Stream() for _ in range(n_streams)]
main = torch.cuda.current_stream()
side = torch.cuda.Stream()
for name, mask in (("all", rz.STAGE_ALL), ("preprocess", rz.STAGE_PREPROCESS), ("bin", rz.STAGE_BIN), ("sort", rz.STAGE_SORT),
                   ("blend", rz.STAGE_BLEND)):
    graphs = []
    with torch.cuda.stream(side):
        for pl in slots:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                pl.launch(mask)
            graphs.append(g)
    torch.cuda.synchronize()
    def run(count):
        f = torch.cuda.Event(); f.record(main)
        for st in streams: st.wait_event(f)
        for i in range(count):
            with torch.cuda.stream(streams[(i % n_slots) % n_streams]):
                graphs[i % n_slots].replay()
        for st in streams:
            j = torch.cuda.Event(); j.record(st); main.wait_event(j)
    run(50); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 1000
    e0.record(); run(K); e1.record(); torch.cuda.synchronize()
    print(f"{name:10s} {n_streams} streams: {1e3 * e0.elapsed_time(e1) / K:7.2f} us per scene", flush=True)
