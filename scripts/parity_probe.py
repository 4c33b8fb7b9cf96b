"""Parity calibration probe (GPU): numbers quoted in DESIGN.md §4 and used to set the bars in tests/.
  1. raster: count and max error of the oracle's threshold-sensitive ("sens") pixels per scene, all-pixel max RGB error
  2. encoder: bf16 tcgen05 path vs reference golden (sampled) on all six outputs, and rendered-RGB drift of the bf16-path
     Gaussians vs the fp32-path Gaussians through the rasterizer.
python scripts/parity_probe.py [raster] [encoder]"""
import copy
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from styl3r_b200 import synthetic as syn  # noqa: E402
from tests.helpers import gpu_scene, oracle_scene  # noqa: E402


def raster():
    scenes = {"cfg2": syn.make_scene(seed=1234, v=2, V=1, hw=256), "small0": syn.make_small_scene(seed=0, P=600, W=64, H=48, V=2),
              "small4": syn.make_small_scene(seed=4, P=5000, W=250, H=130, V=2), "v3_128": syn.make_scene(seed=7, v=2, V=3, hw=128)}
    for name, sc in scenes.items():
        outs, cams = oracle_scene(sc)
        color, depth, opacity, radii, nt, ctx = gpu_scene(sc, cams)
        torch.cuda.synchronize()
        H, W = sc["image_shape"]
        nc = ctx.view("n_contrib").cpu().numpy().reshape(len(cams), H, W).view(np.uint32)
        for v, o in enumerate(outs):
            sens = o["sens"] > 0
            dc = np.abs(color[v].cpu().numpy() - o["color"])
            ncd = nc[v] != o["n_contrib"]
            print(json.dumps(dict(scene=name, view=v, pixels=int(sens.size), sens_pixels=int(sens.sum()),
                                  max_err_all=float(dc.max()), max_err_sens=float(dc[:, sens].max(initial=0)),
                                  max_err_nonsens=float(dc[:, ~sens].max(initial=0)), n_over_1e4=int((dc.max(0) > 1e-4).sum()),
                                  n_contrib_diff=int(ncd.sum()), n_contrib_diff_nonsens=int((ncd & ~sens).sum()),
                                  colour_max=float(np.abs(o["color"]).max()))))


def encoder():
    from styl3r_b200.decoder import render_cuda
    from styl3r_b200.encoder import EncoderNoPoSplatTokenStyleCfg, GraphedEncoder, get_encoder
    from tests.encoder_weights import fill_named_weights, make_inputs
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    g = np.load(ROOT / "tests/golden/encoder_golden.npz")
    enc, _ = get_encoder(EncoderNoPoSplatTokenStyleCfg(stylized=True))
    fill_named_weights(enc)
    enc = enc.cuda().eval()
    context, style = make_inputs(1, 2, 256, seed=1234, device="cuda")

    def sample(t, n=4096):
        f = t.detach().reshape(-1)
        idx = torch.linspace(0, f.numel() - 1, min(n, f.numel())).long().to(f.device)
        return f[idx].float().cpu().numpy()

    d32 = {}
    with torch.no_grad():
        o32 = enc(context, style, visualization_dump=d32)
    res = {}
    # the reference's own numerics on a GPU: fp32 modules with TF32 matmuls / convolutions (croco.py:13)
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    dtf = {}
    with torch.no_grad():
        otf = enc(context, style, visualization_dump=dtf)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    for name, tt in [("means", otf.means), ("covariances", otf.covariances), ("harmonics", otf.harmonics),
                     ("opacities", otf.opacities), ("scales", dtf["scales"]), ("rotations", dtf["rotations"])]:
        ref, scale = g[f"b1v2_{name}"], float(g[f"b1v2_{name}_stats"][2])
        e = np.abs(sample(tt) - ref)
        res[f"tf32_eager_{name}"] = dict(max_over_sigma=float(e.max() / scale), mean_over_sigma=float(e.mean() / scale))
    from styl3r_b200.encoder import vit
    for variant in ("tcgen05", "tcgen05_bf16stream", "cudnn"):
        vit.STREAM_FP32 = variant != "tcgen05_bf16stream"
        fast = copy.deepcopy(enc).to_inference(torch.bfloat16, heads="cudnn" if variant == "cudnn" else "tcgen05")
        dump = {}
        with torch.no_grad():
            ob = fast(context, style, visualization_dump=dump)
        torch.cuda.synchronize()
        for name, tb, t32 in [("means", ob.means, o32.means), ("covariances", ob.covariances, o32.covariances),
                              ("harmonics", ob.harmonics, o32.harmonics), ("opacities", ob.opacities, o32.opacities),
                              ("scales", dump["scales"], d32["scales"]), ("rotations", dump["rotations"], d32["rotations"])]:
            ref, scale = g[f"b1v2_{name}"], float(g[f"b1v2_{name}_stats"][2])
            e = np.abs(sample(tb) - ref)
            efull = (tb.float() - t32.float()).abs()
            res[f"{variant}_{name}"] = dict(max_over_sigma=float(e.max() / scale), mean_over_sigma=float(e.mean() / scale),
                                            p99_over_sigma=float(np.quantile(e, 0.99) / scale),
                                            vs_fp32_full_max_over_sigma=float(efull.max() / scale), sigma=scale)
        # rendered-RGB drift: bf16-path Gaussians vs fp32-path Gaussians from the two context cameras + a novel one
        z = o32.means[0, :, 2]
        zmed = float(z.median())
        print("means z quantiles", [float(q) for q in torch.quantile(z[::16], torch.tensor([0.01, 0.5, 0.99], device=z.device))])
        V = 3
        extr = torch.eye(4, device="cuda").repeat(V, 1, 1)
        if zmed < 0:
            extr[:, 0, 0] = extr[:, 2, 2] = -1.0
        extr[1, 0, 3] = 0.1 * abs(zmed)
        extr[2, 0, 3] = -0.05 * abs(zmed)
        extr[2, 1, 3] = 0.05 * abs(zmed)
        K = context["intrinsics"][0, :1].float().expand(V, 3, 3).contiguous()
        near = torch.full((V,), max(1e-3, 0.05 * abs(zmed)), device="cuda")
        far = torch.full((V,), 1000 * abs(zmed), device="cuda")
        bg = torch.zeros(V, 3, device="cuda")
        vs = torch.zeros(V, dtype=torch.int32, device="cuda")
        with torch.no_grad():
            ca, da = render_cuda(extr, K, near, far, (256, 256), bg, o32.means, o32.covariances, o32.harmonics, o32.opacities, view_set=vs)
            cb, db = render_cuda(extr, K, near, far, (256, 256), bg, ob.means.float(), ob.covariances.float(), ob.harmonics.float(),
                                 ob.opacities.float(), view_set=vs)
        d = (ca - cb).abs()
        mse = float(((ca - cb) ** 2).mean())
        peak = float(ca.abs().max())
        res[f"{variant}_render"] = dict(max_abs=float(d.max()), mean_abs=float(d.mean()), p99=float(torch.quantile(d.flatten()[::7], 0.99)),
                                        psnr_db=float(10 * np.log10(peak * peak / max(mse, 1e-30))), peak=peak,
                                        coverage=float((ca.abs().sum(1) > 0).float().mean()))
    vit.STREAM_FP32 = True
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    what = sys.argv[1:] or ["raster", "encoder"]
    if "raster" in what:
        raster()
    if "encoder" in what:
        encoder()
