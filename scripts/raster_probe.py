"""Development probe (GPU): per-stage times of the raster forward chain on the bench scene (cfg2), each stage timed as a
CUDA graph of REP back-to-back launches on one stream (launch overhead amortised; L2-warm), the whole chain likewise,
and the blend kernel on single tiles (S3R_TUNE_BLEND_ONLY_TILE) to separate the critical path of the heaviest tile from
machine throughput.  python scripts/raster_probe.py [rep]"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch

import bench
from styl3r_b200 import _lib
from styl3r_b200 import rasterizer as rz
from styl3r_b200.decoder import cuda_splatting as cs

REP = 20
dev = torch.device("cuda", 0)


def make_plan(seed, V=1):
    sc = bench.make_scene(seed)
    t = lambda a: torch.as_tensor(a, device=dev)
    view_t, full, proj_t, campos, tanfov, scale = cs.camera_setup(t(sc["extrinsics"]), t(sc["intrinsics"]), t(sc["near"]),
                                                                    t(sc["far"]), True)
    tensors = (t(sc["means"])[None], t(sc["covariances"])[None], t(sc["opacities"])[None],
               t(sc["harmonics"])[None].reshape(1, -1, 1, 3), None, view_t, full, proj_t, campos, tanfov, scale,
               torch.zeros(V, 3, device=dev), torch.zeros(V, dtype=torch.int32, device=dev))
    P = sc["means"].shape[0]
    plan = rz.RasterPlan(tensors, 1, P, V, 256, 256, 1, 0, 9, 3 * P)
    plan.launch()
    torch.cuda.synchronize()
    return plan


def time_graph(fn, rep=REP, iters=20):
    side = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        fn()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=side):
            for _ in range(rep):
                fn()
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / (iters * rep)  # us per launch


def throughput(plans, n_streams=8, iters=40, mask=rz.STAGE_ALL):
    """bench.py's `value` operating point: chain graphs of independent scenes round-robin on n_streams streams."""
    side = torch.cuda.Stream()
    graphs = []
    with torch.cuda.stream(side):
        for pl in plans:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                pl.launch(mask)
            graphs.append(g)
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream() for _ in range(n_streams)]
    main_s = torch.cuda.current_stream()

    def run(count):
        ev = torch.cuda.Event()
        ev.record(main_s)
        for st in streams:
            st.wait_event(ev)
        for i in range(count):
            with torch.cuda.stream(streams[i % n_streams]):
                graphs[i % len(graphs)].replay()
        for st in streams:
            j = torch.cuda.Event()
            j.record(st)
            main_s.wait_event(j)

    run(len(plans) * 2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = len(plans) * iters
    e0.record()
    run(n)
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / n


def main():
    L = _lib.lib()
    quick = "quick" in sys.argv
    plan = make_plan(1234)
    out = {}
    L.s3r_set_tunable(5, 1)
    for mask in ((0,) if quick else (0, 5, 31)):
        L.s3r_set_tunable(9, mask)
        d = {}
        for name, m in [("preprocess", rz.STAGE_PREPROCESS), ("bin", rz.STAGE_BIN), ("sort", rz.STAGE_SORT),
                        ("blend", rz.STAGE_BLEND), ("front", 7), ("chain", rz.STAGE_ALL)]:
            d[name] = round(time_graph(lambda: plan.launch(m)), 2)
        out[f"pdlmask{mask}"] = d
    L.s3r_set_tunable(9, 0)
    plans = [plan] + [make_plan(1235 + i) for i in range(7)]
    out["tp8_chain_us_per_view"] = round(throughput(plans), 2)
    out["tp8_blend_us_per_view"] = round(throughput(plans, mask=rz.STAGE_BLEND), 2)
    out["tp8_front_us_per_view"] = round(throughput(plans, mask=7), 2)
    # single-tile blends
    rg = plan.ctx.view("ranges").cpu().numpy().reshape(-1, 2)
    cnt = rg[:, 1] - rg[:, 0]
    order = np.argsort(cnt)
    for tag, tile in [("heaviest", int(order[-1])), ("median", int(order[len(order) // 2]))]:
        L.s3r_set_tunable(8, tile + 1)
        out[f"blend_tile_{tag}"] = dict(tile=tile, n=int(cnt[tile]), us=round(time_graph(lambda: plan.launch(rz.STAGE_BLEND)), 2))
    L.s3r_set_tunable(8, 0)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
