"""Training-side host logic on CPU (SURVEY.md §8 rows a16 / f4 / e):
  * LossStyle / IdentityLoss (fp32 torch path) against golden values AND gradients produced by the REFERENCE's own
    loss modules (tests/golden/make_loss_golden.py);
  * parameter selection / optimiser set-up of configure_optimizers;
  * TrainStep under a world_size-2 gloo group: the gradient all-reduce leaves both ranks with the mean gradient and
    identical weights."""
import os
import socket
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn

from tests.encoder_weights import fill_vgg_named

GOLD = np.load(Path(__file__).parent / "golden" / "loss_golden.npz")


class _Out:
    def __init__(self, color):
        self.color, self.depth = color, None


def _batch(seed=0, b=2, v=2, hw=32):
    g = torch.Generator().manual_seed(seed)
    pred = torch.rand(b, v, 3, hw, hw, generator=g)
    batch = {"target": {"image": torch.rand(b, v, 3, hw, hw, generator=g)}, "style": {"image": torch.rand(b, 3, hw, hw, generator=g)}}
    return pred, batch


def test_vgg_state_dict_names_match_reference():
    from styl3r_b200.train import VGGEncoder
    assert list(VGGEncoder().state_dict().keys()) == list(GOLD["vgg_buffer_names"])


def test_losses_match_reference_values_and_gradients():
    from styl3r_b200.train import IdentityLoss, LossStyle, LossStyleCfg, LossStyleCfgWrapper
    style, ident = LossStyle(LossStyleCfgWrapper(LossStyleCfg(10.0))), IdentityLoss()
    for m in (style, ident):
        fill_vgg_named(m.vgg.named_parameters())
        assert len(list(m.parameters())) == 0 and len(m.state_dict()) == 0   # frozen, outside the checkpoint
    pred, batch = _batch()
    pred.requires_grad_(True)
    ls = style(_Out(pred), batch, None, 0)
    ls.backward()
    assert abs(float(ls.detach()) - float(GOLD["style_loss"])) <= 1e-5 * abs(float(GOLD["style_loss"]))
    assert np.abs(pred.grad.numpy() - GOLD["style_grad"]).max() <= 1e-5 * np.abs(GOLD["style_grad"]).max()
    pred.grad = None
    li = ident(_Out(pred), batch, None, 0)
    li.backward()
    assert abs(float(li.detach()) - float(GOLD["identity_loss"])) <= 1e-5 * abs(float(GOLD["identity_loss"]))
    assert np.abs(pred.grad.numpy() - GOLD["identity_grad"]).max() <= 1e-5 * np.abs(GOLD["identity_grad"]).max()


class _Gauss:
    def __init__(self, means, covariances, harmonics, opacities):
        self.means, self.covariances, self.harmonics, self.opacities = means, covariances, harmonics, opacities


class TinyEncoder(nn.Module):
    """CPU stand-in with the reference's parameter NAMES (what select_trainable keys on)."""
    stylized = True

    def __init__(self):
        super().__init__()
        self.backbone = nn.Linear(3, 3)
        self.token_stylizer = nn.Module()
        self.token_stylizer.enc_blocks = nn.Linear(3, 3)
        self.token_stylizer.dec_blocks = nn.Linear(3, 3)
        self.token_stylizer.patch_embed = nn.Linear(3, 3)
        self.gaussian_appearance_head = nn.Linear(3, 3)
        self.gaussian_param_head = nn.Linear(3, 3)
        self.unused = nn.Linear(3, 3)   # never touched by forward: find_unused_parameters must cope

    def forward(self, context, style, global_step=0):
        x = context["image"].mean(dim=(1, 3, 4))                                   # [b, 3]
        s = style["image"].mean(dim=(2, 3))
        h = self.token_stylizer.dec_blocks(self.token_stylizer.enc_blocks(self.token_stylizer.patch_embed(s)) + self.backbone(x))
        c = self.gaussian_appearance_head(h) + self.gaussian_param_head(x)
        return _Gauss(c, None, c[:, :, None], c[:, 0])


class TinyDecoder(nn.Module):
    def forward(self, g, extr, intr, near, far, shape, depth_mode=None):
        b, V = extr.shape[:2]
        return _Out(torch.sigmoid(g.harmonics[:, None, :, :, None]).expand(b, V, 3, *shape))


class _MSE(nn.Module):
    name = "mse"

    def forward(self, out, batch, g, step):
        return ((out.color - batch["target"]["image"]) ** 2).mean()


def _tiny_batch(seed, b=2, V=2, hw=8):
    g = torch.Generator().manual_seed(seed)
    return {"context": {"image": torch.rand(b, 2, 3, hw, hw, generator=g)},
            "target": {"image": torch.rand(b, V, 3, hw, hw, generator=g), "extrinsics": torch.eye(4).expand(b, V, 4, 4),
                       "intrinsics": torch.eye(3).expand(b, V, 3, 3), "near": torch.ones(b, V), "far": torch.ones(b, V)},
            "style": {"image": torch.rand(b, 3, hw, hw, generator=g)}}


def test_select_trainable_and_optimizer_groups():
    from styl3r_b200.train import configure_optimizers, select_trainable
    enc = TinyEncoder()
    new, pre, frozen = select_trainable(enc)
    assert {id(p) for p in new} == {id(p) for m in (enc.token_stylizer.dec_blocks, enc.gaussian_appearance_head) for p in m.parameters()}
    assert {id(p) for p in pre} == {id(p) for m in (enc.token_stylizer.enc_blocks, enc.token_stylizer.patch_embed) for p in m.parameters()}
    assert all(not p.requires_grad for p in enc.backbone.parameters()) and "encoder.backbone.weight" in frozen
    enc = TinyEncoder()
    opt, sched = configure_optimizers(enc, lr=2e-4, backbone_lr_multiplier=0.1, warm_up_steps=4, max_steps=10)
    assert opt.defaults["betas"] == (0.9, 0.95) and opt.defaults["weight_decay"] == 0.05
    lrs = []
    for _ in range(6):
        lrs.append([g["lr"] for g in opt.param_groups])
        opt.step(); sched.step()
    assert lrs[0][0] == pytest.approx(2e-4 / 4) and lrs[0][1] == pytest.approx(2e-5 / 4)      # LinearLR warm-up from lr / steps
    assert lrs[4][0] == pytest.approx(2e-4) and lrs[5][0] < 2e-4                              # then cosine decay


def _train_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from styl3r_b200.train import TrainStep
        torch.manual_seed(0)
        enc = TinyEncoder()
        step = TrainStep(enc, TinyDecoder(), [_MSE()], identity_loss=None, lr=1e-2, warm_up_steps=1, max_steps=10,
                         gradient_clip_val=0.5)
        assert step.ddp is not None
        w0 = enc.gaussian_appearance_head.weight.detach().clone()
        loss, logs = step(_tiny_batch(100 + rank))          # different data per rank
        grads = enc.gaussian_appearance_head.weight.grad.clone()
        q.put((rank, float(loss), grads.numpy(), enc.gaussian_appearance_head.weight.detach().numpy(),
               enc.backbone.weight.grad is None, bool((enc.gaussian_appearance_head.weight != w0).any())))
    finally:
        dist.destroy_process_group()


def test_train_step_world_size_2_gloo_allreduces_gradients():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, l0, g0, w0, frozen0, moved0), (_, l1, g1, w1, frozen1, moved1) = res
    assert l0 != l1                                    # ranks saw different batches
    assert np.array_equal(g0, g1) and np.array_equal(w0, w1)     # ... but hold the same (mean) gradient and weights
    assert frozen0 and frozen1 and moved0 and moved1

    # the all-reduced gradient equals the mean of the single-process gradients of the two batches
    from styl3r_b200.train import training_step, select_trainable
    singles = []
    for r in range(2):
        torch.manual_seed(0)
        enc = TinyEncoder()
        select_trainable(enc)
        loss, _ = training_step(enc, TinyDecoder(), [_MSE()], _tiny_batch(100 + r))
        loss.backward()
        singles.append(enc.gaussian_appearance_head.weight.grad.numpy())
    mean = 0.5 * (singles[0] + singles[1])
    assert np.allclose(g0 / np.abs(g0).max(), mean / np.abs(mean).max(), atol=1e-5)   # same direction (clip rescales the norm)


def test_compute_psnr_matches_reference_formula():
    from styl3r_b200.train.step import compute_psnr
    g = torch.Generator().manual_seed(0)
    a, b = torch.rand(3, 3, 8, 8, generator=g) * 1.2 - 0.1, torch.rand(3, 3, 8, 8, generator=g)
    ref = -10 * ((a.clip(0, 1) - b.clip(0, 1)) ** 2).mean(dim=(1, 2, 3)).log10()
    assert torch.allclose(compute_psnr(a, b), ref)


def test_vgg_weights_loader_and_never_loaded_warning():
    """the reference's VGG comes from torchvision (non-persistent buffers, absent from checkpoints) - the losses
    need an explicit loader, must warn when they run on never-loaded weights, and cached operands must follow reloads."""
    import warnings
    import torch
    from styl3r_b200.train import IdentityLoss, VGGEncoder
    from styl3r_b200.train.vgg import VGG19_CFG
    g = torch.Generator().manual_seed(0)
    sd = {}
    for item in VGG19_CFG:
        if item != "M":
            i, cin, cout = item
            sd[f"features.{i}.weight"] = torch.randn(cout, cin, 3, 3, generator=g) * 0.05
            sd[f"features.{i}.bias"] = torch.randn(cout, generator=g) * 0.01
    sd["features.21.weight"] = torch.zeros(1)  # deeper layers of the full model are ignored
    img = torch.rand(1, 3, 32, 32, generator=g)
    vgg = VGGEncoder(fast=False)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        vgg(img)
        vgg(img)
    assert sum(issubclass(x.category, RuntimeWarning) for x in w) == 1   # warned once
    vgg.load_vgg19_features(sd)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        f = vgg(img)
    assert not w
    ref = torch.relu(torch.nn.functional.conv2d(img, sd["features.0.weight"], sd["features.0.bias"], padding=1))
    assert torch.allclose(f[0], ref, atol=1e-6)
    with pytest.raises(KeyError):
        VGGEncoder(fast=False).load_vgg19_features({"features.0.weight": sd["features.0.weight"]})
    loss = IdentityLoss(fast=False)
    assert "vgg" not in "".join(loss.state_dict().keys())        # like the reference: not part of checkpoints
    loss.load_vgg19_features(sd)
    assert torch.equal(loss.vgg.slice1[0].weight, sd["features.0.weight"])
    # the fast path's cached operands are keyed on the parameter versions: an in-place reload invalidates them
    v0 = loss.vgg._param_versions()
    loss.vgg.load_vgg19_features(sd)
    assert loss.vgg._param_versions() != v0 and loss.vgg._prep is None
