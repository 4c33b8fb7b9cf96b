"""Supplementary legs of bench.py: the BASELINE.json configurations beyond the headline cfg2-raster line.

Every leg is bounded (a few seconds on a B200), uses synthetic inputs and random weights (`data: synthetic`), times on
the device with CUDA events after warm-up, and returns a plain dict that bench.py stores under its own key:

  cfg2_full      host images -> H2D -> encoder (GraphedEncoder) -> DecoderSplattingCUDA.forward -> D2H image, 1 scene,
                 v=2, 1 target view (infer_model_re10k.py:404-560 without the pose-align loop)
  cfg3           b=4 scenes, v=4 context views (G = 262 144), 6 target views each: the same pipeline, 24 views per pass
  cfg4           = cfg3 on every rank (4 scenes x 6 views per GPU; 32 x 6 on 8 GPUs), aggregated by bench.py
  cfg5           stage-2 training step fwd+bwd (style + identity pass, raster backward, VGG losses, AdamW) at batch
                 `train_batch` per GPU, DDP gradient all-reduce when world > 1
  e2e_render_cuda  the reference-signature eager call: host tensors -> .cuda() -> render_cuda -> .cpu()
  encoder_comparator  the same module tree on the torch libraries (bf16 cuBLAS / cuDNN / SDPA under a CUDA graph; TF32
                 eager = the reference's own GPU numerics) next to ours
  pose_align     50 graph-captured pose-alignment steps (raster fwd + bwd + Adam + SE3), cfg2 scene
"""
from __future__ import annotations

import time

import numpy as np
import torch

HW = 256


def _events():
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def _timeit(fn, iters, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = _events()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def make_encoder(dev, inference=True):
    from styl3r_b200.encoder import EncoderNoPoSplatTokenStyleCfg, get_encoder
    torch.manual_seed(0)
    enc, _ = get_encoder(EncoderNoPoSplatTokenStyleCfg(stylized=True))
    enc = enc.to(dev)
    if inference:
        enc = enc.eval().to_inference(torch.bfloat16)
    return enc


def _scene_inputs(b, v, V, seed=1234, pin=True):
    """Host-side request: context images / intrinsics, style image, target cameras (SURVEY §8d recipe)."""
    from styl3r_b200 import synthetic as syn
    g = torch.Generator().manual_seed(seed)
    p = (lambda t: t.pin_memory()) if pin else (lambda t: t)
    K = torch.tensor([[0.8, 0, 0.5], [0, 0.8, 0.5], [0, 0, 1.0]])
    ctx = {"image": p(torch.rand(b, v, 3, HW, HW, generator=g) * 2 - 1), "intrinsics": p(K.expand(b, v, 3, 3).contiguous())}
    sty = {"image": p(torch.rand(b, 3, HW, HW, generator=g) * 2 - 1)}
    extr = np.stack([syn.make_scene(seed=seed + s, v=2, V=V, hw=8)["extrinsics"] for s in range(b)])
    cams = dict(extrinsics=p(torch.as_tensor(extr)), intrinsics=p(K.expand(b, V, 3, 3).contiguous()),
                near=p(torch.full((b, V), 0.1)), far=p(torch.full((b, V), 100.0)))
    return ctx, sty, cams


def _trained_like_gaussians(fast, ctx_dev, sty_dev, b, v, dev, seed=1234):
    """Random weights scatter the Gaussians (|z| up to 60, most behind the camera), which would leave the rasterizer with
    almost no work.  For the full-pipeline legs the encoder runs on its real inputs (its time is input-independent) and
    the rasterizer renders "trained-like" pixel-aligned Gaussians of the same shapes (SURVEY §8d recipe) written into the
    encoder's static output buffers - so both halves do the work they do in production."""
    from styl3r_b200 import synthetic as syn
    out = fast(ctx_dev, sty_dev)
    scs = [syn.make_scene(seed=seed + s, v=v, V=1, hw=HW) for s in range(b)]
    t = lambda k: torch.as_tensor(np.stack([sc[k] for sc in scs]), device=dev)
    return out, (t("means"), t("covariances"), t("harmonics"), t("opacities"))


def full_pipeline(dev, b, v, V, iters, label):
    """host images -> H2D -> encoder graph -> decoder (b*V views in one launch chain) -> D2H colour."""
    from styl3r_b200.decoder import DecoderSplattingCUDA, DecoderSplattingCUDACfg
    from styl3r_b200.encoder import GraphedEncoder
    enc = make_encoder(dev)
    fast = GraphedEncoder(enc)
    dec = DecoderSplattingCUDA(DecoderSplattingCUDACfg("splatting_cuda", [0.0, 0.0, 0.0], True)).to(dev)
    ctx, sty, cams = _scene_inputs(b, v, V)
    to = lambda d: {k: t.to(dev, non_blocking=True) for k, t in d.items()}
    out, trained = _trained_like_gaussians(fast, to(ctx), to(sty), b, v, dev)
    color_host = torch.empty(b, V, 3, HW, HW).pin_memory()
    h2d = sum(t.numel() * t.element_size() for d in (ctx, sty, cams) for t in d.values())
    d2h = color_host.numel() * 4
    stage = {}

    def step(measure=False):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if measure else None
        if measure:
            e[0].record()
        g = fast(ctx, sty)                       # pinned host -> static device buffers -> graph replay
        g.means.copy_(trained[0]); g.covariances.copy_(trained[1]); g.harmonics.copy_(trained[2]); g.opacities.copy_(trained[3])
        if measure:
            e[1].record()
        c = to(cams)
        o = dec(g, c["extrinsics"], c["intrinsics"], c["near"], c["far"], (HW, HW))
        if measure:
            e[2].record()
        color_host.copy_(o.color, non_blocking=True)
        if measure:
            e[3].record()
            torch.cuda.synchronize()
            stage.update(encoder_ms=e[0].elapsed_time(e[1]), decoder_ms=e[1].elapsed_time(e[2]), d2h_ms=e[2].elapsed_time(e[3]))

    with torch.no_grad():
        ms = _timeit(step, iters)                # back-to-back requests on one stream (throughput)
        step(measure=True)
        t0 = time.perf_counter()
        step()
        torch.cuda.synchronize()
        latency = (time.perf_counter() - t0) * 1e3
    flops = {2: 1270.8, 4: 2437.2}.get(v)
    res = {"workload": label, "ms_per_pass": ms, "views_per_s": b * V / (ms * 1e-3), "scenes_per_s": b / (ms * 1e-3),
           "latency_ms_sync": latency, "h2d_bytes_per_pass": h2d, "d2h_bytes_per_pass": d2h, **stage,
           "api": "GraphedEncoder(encoder.to_inference(bf16))(host context, host style) -> DecoderSplattingCUDA.forward -> "
                  "pinned host colour; rasterized Gaussians are trained-like (encoder outputs of random weights are degenerate)"}
    if flops:
        res["encoder_tflops"] = b * flops / stage["encoder_ms"]
    del fast, enc
    torch.cuda.empty_cache()
    return res


def e2e_render_cuda(dev, slots_host, iters=100):
    """The call an unchanged infer script makes: host tensors -> device -> render_cuda(...) -> host, eager Python."""
    from styl3r_b200.decoder import render_cuda
    n = len(slots_host)

    def step(i=[0]):
        h = slots_host[i[0] % n]
        i[0] += 1
        d = {k: v.to(dev, non_blocking=True) for k, v in h.items() if k != "cov6"}
        color, depth = render_cuda(d["extr"], d["intr"], d["near"], d["far"], (HW, HW), d["bg"], d["means"], d["cov"], d["sh"],
                                   d["opac"], scale_invariant=True)
        return color.cpu()

    with torch.no_grad():
        for _ in range(5):
            step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(iters):
            step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / iters
    return {"views_per_s": 1.0 / dt, "ms_per_view": dt * 1e3, "api": "styl3r_b200.decoder.render_cuda (reference signature, "
            "cuda_splatting.py:46-61), pinned host inputs .to(device), result .cpu(); eager, one stream, synchronous"}


def encoder_comparator(dev, iters=10):
    """Same module tree, same shapes (cfg2: b=1, v=2): ours vs the torch libraries on this GPU."""
    from baseline import library_encoder as le
    from styl3r_b200.encoder import GraphedEncoder
    ctx, sty, _ = _scene_inputs(1, 2, 1, pin=False)
    to = lambda d: {k: t.to(dev) for k, t in d.items()}
    ctx, sty = to(ctx), to(sty)
    res = {}
    with torch.no_grad():
        enc = make_encoder(dev)
        fast = GraphedEncoder(enc)
        res["ours_tcgen05_graph_ms"] = _timeit(lambda: fast(ctx, sty), iters)
        del fast, enc
        torch.cuda.empty_cache()
        res.update(le.time_library_encoder(dev, ctx, sty, iters))
    res["what"] = ("cfg2 encoder (b=1, v=2, 256x256 + style), random weights.  library_bf16_graph: the same nn.Module tree with bf16 "
                   "trunks on cuBLAS / SDPA / ATen LayerNorm and the DPT heads on cuDNN (bf16 autocast, channels_last), same stream "
                   "branches, replayed as a CUDA graph; library_tf32_eager: fp32 modules with TF32 matmuls and convolutions, "
                   "eager - the reference's own configuration (croco.py:13)")
    return res


def pose_align_leg(dev, steps=50):
    from types import SimpleNamespace
    from styl3r_b200 import synthetic as syn
    from styl3r_b200.decoder import render_cuda
    from styl3r_b200.pose_align import pose_align
    sc = syn.make_scene(seed=1234, v=2, V=1, hw=HW)
    t = lambda a: torch.as_tensor(a, device=dev)
    g = SimpleNamespace(means=t(sc["means"])[None], covariances=t(sc["covariances"])[None], harmonics=t(sc["harmonics"])[None],
                        opacities=t(sc["opacities"])[None])
    extr, intr, near, far = t(sc["extrinsics"])[None], t(sc["intrinsics"])[None], t(sc["near"])[None], t(sc["far"])[None]
    with torch.no_grad():
        target, _ = render_cuda(extr[0], intr[0], near[0], far[0], (HW, HW), torch.zeros(1, 3, device=dev), g.means,
                                g.covariances, g.harmonics, g.opacities)
    pert = extr.clone()
    pert[0, :, 0, 3] += 0.02
    pose_align(g, pert, intr, near, far, (HW, HW), target[None], steps=3)  # warm-up (allocations, capacity probe)
    torch.cuda.synchronize()

    def timed(n):
        t0 = time.perf_counter()
        refined, losses = pose_align(g, pert, intr, near, far, (HW, HW), target[None], steps=n)
        torch.cuda.synchronize()
        return time.perf_counter() - t0, losses

    dt, losses = timed(steps)
    dt3, _ = timed(3 * steps)  # the marginal cost of a step = graph replays only (set-up and capture cancel out)
    l = losses.cpu().numpy()
    return {"steps": steps, "total_ms": dt * 1e3, "ms_per_step": (dt3 - dt) * 1e3 / (2 * steps), "loss_first": float(l[0]),
            "loss_last": float(l[-1]),
            "what": "test_step_align (infer_model_re10k.py:79-161) on device: cfg2 scene, 1 target view 256x256, MSE loss; one CUDA "
                    "graph per iteration (camera kernel + raster fwd + loss grad + raster bwd (dL/dtau only) + Adam + SE3 update); "
                    "total_ms = wall time of the 50-step call incl. the capacity probe and graph capture, ms_per_step = marginal "
                    "cost of a step (150-step call minus 50-step call, / 100)"}


def train_step_leg(dev, batch, steps=3, world=1, layout="bf16"):
    """BASELINE cfg5: 2 context views + style image, `batch` scenes per GPU, 4 target views, style loss + identity pass."""
    from styl3r_b200 import synthetic as syn
    from styl3r_b200.decoder import DecoderSplattingCUDA, DecoderSplattingCUDACfg
    from styl3r_b200.train import IdentityLoss, LossStyle, LossStyleCfg, LossStyleCfgWrapper, TrainStep
    V = 4
    enc = make_encoder(dev, inference=False)
    enc.to_training(torch.bfloat16 if layout == "bf16" else None)
    dec = DecoderSplattingCUDA(DecoderSplattingCUDACfg("splatting_cuda", [0.0, 0.0, 0.0], True)).to(dev)
    style_loss = LossStyle(LossStyleCfgWrapper(LossStyleCfg(10.0))).to(dev)
    ident = IdentityLoss().to(dev)
    step = TrainStep(enc, dec, [style_loss], ident)
    n_train = sum(p.numel() for g in step.optimizer.param_groups for p in g["params"])
    gsel = torch.Generator().manual_seed(7)
    K = torch.tensor([[0.8, 0, 0.5], [0, 0.8, 0.5], [0, 0, 1.0]])
    scs = [syn.make_scene(seed=100 + s, v=2, V=V, hw=8) for s in range(batch)]
    batch_d = {
        "context": {"image": torch.rand(batch, 2, 3, HW, HW, generator=gsel).to(dev), "intrinsics": K.expand(batch, 2, 3, 3).contiguous().to(dev),
                    "extrinsics": torch.as_tensor(np.stack([s["context_extrinsics"] for s in scs])).to(dev),
                    "near": torch.full((batch, 2), 0.1, device=dev), "far": torch.full((batch, 2), 100.0, device=dev)},
        "target": {"image": torch.rand(batch, V, 3, HW, HW, generator=gsel).to(dev), "intrinsics": K.expand(batch, V, 3, 3).contiguous().to(dev),
                   "extrinsics": torch.as_tensor(np.stack([s["extrinsics"] for s in scs])).to(dev),
                   "near": torch.full((batch, V), 0.1, device=dev), "far": torch.full((batch, V), 100.0, device=dev)},
        "style": {"image": torch.rand(batch, 3, HW, HW, generator=gsel).to(dev)},
    }
    enc.train()
    torch.backends.cuda.matmul.allow_tf32 = True   # the reference's setting (croco.py:13)
    torch.backends.cudnn.allow_tf32 = True
    for _ in range(2):                             # warm-up (allocator growth, cuDNN plans, DDP buckets)
        step(batch_d)
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats(dev)
    per_step = []
    for _ in range(steps + 1):                     # every step timed on its own; the median is reported (a single step
        e0, e1 = _events()                         # that hits an allocator re-shuffle after the previous legs took 1.5x)
        e0.record()
        loss, logs = step(batch_d)
        e1.record()
        torch.cuda.synchronize()
        per_step.append(e0.elapsed_time(e1))
    ms = sorted(per_step)[len(per_step) // 2]
    res = {"ms_per_step": ms, "scenes_per_s_per_gpu": batch / (ms * 1e-3), "batch_per_gpu": batch, "target_views": V,
           "trainable_params": n_train, "grad_allreduce_bytes": 4 * n_train if world > 1 else 0,
           "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 1e9, "loss": float(loss),
           "ms_per_step_each": [round(t, 1) for t in per_step], "timing": "median of the individually timed steps",
           "what": "stage-2 step: encoder fwd x2 (style + identity pass) -> rasterizer fwd+bwd (our kernels) -> VGG style / identity "
                   "losses (tcgen05 convolutions, fwd + dgrad) -> encoder backward -> "
                   + ("DDP bucketed NCCL all-reduce -> " if world > 1 else "") + "clip -> AdamW",
           "layout": layout, "encoder_backward": step.backward_kind()}
    enc.to_training(None)
    del step, enc
    torch.cuda.empty_cache()
    return res
