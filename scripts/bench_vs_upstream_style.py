"""Measured GPU comparator for the '>= 10x the reference rasterizer' target: the upstream-style stand-in
(baseline/upstream_style: per-view launch chain, CUB scan + 64-bit radix sort, D2H num_rendered, scalar 16x16 render
kernel, driven by a restatement of the reference's per-view Python loop) vs styl3r_b200 on the same B200, same inputs.
Writes gpurun_out/vs_upstream_style.json."""
import json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np, torch
from baseline import upstream_style as ups
from styl3r_b200 import synthetic as syn
from styl3r_b200.decoder import render_cuda

def timeit(fn, iters):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters

out = {}
for name, (v, V, b) in {"cfg2 (1 scene x 1 view, P=131072)": (2, 1, 1), "cfg3-like (1 scene x 6 views, P=262144)": (4, 6, 1)}.items():
    sc = syn.make_scene(seed=1234, v=v, V=V, hw=256)
    t = lambda a: torch.as_tensor(a).cuda()
    extr, intr, near, far = t(sc["extrinsics"]), t(sc["intrinsics"]), t(sc["near"]), t(sc["far"])
    means, cov, sh, op = t(sc["means"])[None], t(sc["covariances"])[None], t(sc["harmonics"])[None], t(sc["opacities"])[None]
    bg = torch.zeros(V, 3, device="cuda")
    rep = lambda x: x.expand(V, *x.shape[1:]).contiguous()   # the reference's decoder repeats the Gaussians per view
    vs = torch.zeros(V, dtype=torch.int32, device="cuda")
    with torch.no_grad():
        ref_img, _ = ups.render_cuda_upstream_style(extr, intr, near, far, (256, 256), bg, rep(means), rep(cov), rep(sh), rep(op))
        our_img, _ = render_cuda(extr, intr, near, far, (256, 256), bg, means, cov, sh, op, view_set=vs)
        diff = (ref_img - our_img).abs()
        t_ref = timeit(lambda: ups.render_cuda_upstream_style(extr, intr, near, far, (256, 256), bg, rep(means), rep(cov), rep(sh), rep(op)), 20)
        t_ours = timeit(lambda: render_cuda(extr, intr, near, far, (256, 256), bg, means, cov, sh, op, view_set=vs), 50)
        t_ours_async = timeit(lambda: render_cuda(extr, intr, near, far, (256, 256), bg, means, cov, sh, op, view_set=vs, check="deferred"), 50)
    out[name] = {"upstream_style_views_per_s": V / t_ref, "styl3r_b200_render_cuda_views_per_s": V / t_ours,
                 "styl3r_b200_render_cuda_deferred_views_per_s": V / t_ours_async, "speedup_eager_call": t_ref / t_ours,
                 "image_max_abs_diff": float(diff.max()), "image_frac_gt_1e-4": float((diff > 1e-4).float().mean())}
    print(name, json.dumps(out[name]), flush=True)
Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "vs_upstream_style.json").write_text(json.dumps(out, indent=1))
