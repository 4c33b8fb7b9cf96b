"""Training step of Styl3R's stage 2 (BASELINE cfg5) - mirror of
`ModelWrapperStyle.training_step` (src/model/model_wrapper_style.py:118-315, the non-distillation branch),
`configure_optimizers` (:843-916) and the DDP strategy of src/main_style.py:103-108:

    data shim (context images -> [-1,1]) -> style image -> [-1,1] -> encoder -> decoder (all b*V target views in one
    rasterizer launch chain, backward through the CUDA rasterizer) -> sum of losses (+ identity pass with the first
    context image as the style) -> backward -> gradient all-reduce -> clip 0.5 -> AdamW -> LR schedule.

Multi-GPU: one process per GPU; the only collective on this path is the gradient all-reduce of the trainable set
(stage 2: token-stylizer + appearance head), issued bucket by bucket from the backward pass by
`torch.nn.parallel.DistributedDataParallel(find_unused_parameters=True)` (NCCL on GPUs, gloo in the CPU tests) - the
reference's `ddp_find_unused_parameters_true`."""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence, Tuple

import torch
from torch import Tensor, nn


def select_trainable(encoder: nn.Module, prefix: str = "encoder.") -> Tuple[List[nn.Parameter], List[nn.Parameter], List[str]]:
    """Parameter groups of configure_optimizers (:849-868) for a stylized encoder: (new, pretrained, frozen names).
    Names are matched with the Lightning prefix the reference sees ("encoder.token_stylizer.dec_blocks...")."""
    new, pre, frozen = [], [], []
    for name, p in encoder.named_parameters():
        full = prefix + name
        if not p.requires_grad:
            continue
        if getattr(encoder, "stylized", True):
            if "stylizer.dec" in full or "gaussian_appearance_head" in full:
                new.append(p)
            elif "stylizer.enc" in full or "stylizer.mask_token" in full or "stylizer.patch_embed" in full:
                pre.append(p)
            else:
                p.requires_grad = False
                frozen.append(full)
        else:
            if any(k in full for k in ("stylizer.dec", "gaussian_appearance_head", "gaussian_param_head", "intrinsic_encoder")):
                new.append(p)
            else:
                pre.append(p)
    return new, pre, frozen


def configure_optimizers(encoder: nn.Module, lr: float = 2e-4, backbone_lr_multiplier: float = 0.1,
                         warm_up_steps: int = 125, max_steps: int = 18751):
    """AdamW(beta 0.9/0.95, wd 0.05) over (new @ lr, pretrained @ lr * multiplier), LinearLR warm-up then cosine to
    0.1 * lr (:887-916)."""
    new, pre, _ = select_trainable(encoder)
    groups = [{"params": new, "lr": lr}, {"params": pre, "lr": lr * backbone_lr_multiplier}]
    opt = torch.optim.AdamW(groups, lr=lr, weight_decay=0.05, betas=(0.9, 0.95))
    warm = torch.optim.lr_scheduler.LinearLR(opt, 1 / warm_up_steps, 1, total_iters=warm_up_steps)
    cos = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=max_steps, eta_min=lr * 0.1)
    sched = torch.optim.lr_scheduler.SequentialLR(opt, schedulers=[warm, cos], milestones=[warm_up_steps])
    return opt, sched


@torch.no_grad()
def compute_psnr(ground_truth: Tensor, predicted: Tensor) -> Tensor:
    """src/evaluation/metrics.py:12-19: per-image PSNR of [n,c,h,w] tensors clipped to [0,1]."""
    mse = ((ground_truth.clip(min=0, max=1) - predicted.clip(min=0, max=1)) ** 2).flatten(1).mean(dim=1)
    return -10 * mse.log10()


def training_step(encoder: nn.Module, decoder: nn.Module, losses: Sequence[nn.Module], batch: dict, global_step: int = 0,
                  identity_loss: Optional[nn.Module] = None, data_shim=None, depth_mode=None) -> Tuple[Tensor, dict]:
    """Returns (total_loss, logs).  `batch` follows BatchedExample: context {image [b,v,3,h,w] in [0,1], intrinsics,
    ...}, target {image [b,V,3,h,w], extrinsics, intrinsics, near, far}, style {image [b,3,h,w] in [0,1]}."""
    if data_shim is not None:
        batch = data_shim(batch)
    h, w = batch["target"]["image"].shape[-2:]
    if not getattr(encoder, "stylized", True):
        style = {"image": batch["context"]["image"][:, 0]}
    else:
        style = dict(batch["style"])
        style["image"] = (style["image"].clone() - 0.5) / 0.5
    tgt = batch["target"]
    gaussians = encoder(batch["context"], style, global_step)
    output = decoder.forward(gaussians, tgt["extrinsics"], tgt["intrinsics"], tgt["near"], tgt["far"], (h, w),
                             depth_mode=depth_mode)
    logs, total = {}, 0
    logs["train/psnr_probabilistic"] = compute_psnr(tgt["image"].flatten(0, 1), output.color.flatten(0, 1)).mean()
    for loss_fn in losses:
        val = loss_fn.forward(output, batch, gaussians, global_step)
        logs[f"loss/{getattr(loss_fn, 'name', type(loss_fn).__name__)}"] = val.detach()
        total = total + val
    if identity_loss is not None:
        identity_style = {"image": batch["context"]["image"][:, 0]}
        ig = encoder(batch["context"], identity_style, global_step)
        io = decoder.forward(ig, tgt["extrinsics"], tgt["intrinsics"], tgt["near"], tgt["far"], (h, w), depth_mode=depth_mode)
        val = identity_loss(io, batch, ig, global_step)
        logs["loss/identity_loss"] = val.detach()
        total = total + val
    logs["loss/total"] = total.detach()
    return total, logs


class _EncoderForDDP(nn.Module):
    """DDP wraps one module whose forward produces everything the backward needs: the encoder.  (The decoder and the
    losses hold no trainable parameters.)"""

    def __init__(self, encoder: nn.Module):
        super().__init__()
        self.encoder = encoder

    def forward(self, context: dict, styles, global_step: int = 0):
        """`styles`: the style dicts of ALL encoder passes of one training step (stylised pass, identity pass) - like
        Lightning's DDP wrapper around `training_step`, ONE wrapped forward covers every use of the parameters before
        the single backward (the reducer's unused-parameter search runs once per wrapped forward)."""
        out = []
        for style in styles:
            g = self.encoder(context, style, global_step)
            out += [g.means, g.covariances, g.harmonics, g.opacities]
        return tuple(out)


class TrainStep:
    """One optimisation step (forward, backward with the bucketed gradient all-reduce, clip, AdamW, LR step).

        step = TrainStep(encoder, decoder, [LossStyle(cfg)], IdentityLoss(), lr=2e-4)    # after init_process_group
        loss, logs = step(batch)
    """

    def __init__(self, encoder: nn.Module, decoder: nn.Module, losses: Iterable[nn.Module],
                 identity_loss: Optional[nn.Module] = None, lr: float = 2e-4, backbone_lr_multiplier: float = 0.1,
                 warm_up_steps: int = 125, max_steps: int = 18751, gradient_clip_val: float = 0.5, data_shim=None):
        import torch.distributed as dist
        self.encoder, self.decoder = encoder, decoder
        self.losses, self.identity_loss = list(losses), identity_loss
        self.optimizer, self.scheduler = configure_optimizers(encoder, lr, backbone_lr_multiplier, warm_up_steps, max_steps)
        self.clip, self.data_shim, self.global_step = gradient_clip_val, data_shim, 0
        self.ddp = None
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dev = next(encoder.parameters()).device
            # buffers are constants (SH mask, VGG statistics): no per-forward broadcast (which would also bump their
            # version between the two encoder passes of a step)
            self.ddp = nn.parallel.DistributedDataParallel(
                _EncoderForDDP(encoder), device_ids=[dev.index] if dev.type == "cuda" else None,
                find_unused_parameters=True, broadcast_buffers=False)

    def backward_kind(self) -> str:
        """Which kernels run the encoder's backward (reported by bench.py next to the cfg5 number)."""
        from ..encoder import vit
        return getattr(vit, "TRAIN_BACKWARD_KIND", "torch autograd over the reference's fp32/TF32 ops (cuBLAS / cuDNN / SDPA)")

    def _encoder_proxy(self, batch: dict):
        """The `encoder` callable handed to training_step.  Single process: the encoder itself.  DDP: both passes of the
        step (training_step's stylised and identity calls, in that order) are computed by ONE wrapped forward up front
        and handed out in call order."""
        stylized = getattr(self.encoder, "stylized", True)
        if self.ddp is None:
            return type("Enc", (), {"stylized": stylized, "__call__": lambda s, c, st, gs=0: self.encoder(c, st, gs)})()
        from ..encoder.encoder import Gaussians
        b = self.data_shim(batch) if self.data_shim is not None else batch
        ctx = b["context"]
        if stylized:
            st0 = dict(b["style"])
            st0["image"] = (st0["image"].clone() - 0.5) / 0.5
        else:
            st0 = {"image": ctx["image"][:, 0]}
        styles = [st0] + ([{"image": ctx["image"][:, 0]}] if self.identity_loss is not None else [])
        flat = self.ddp(ctx, styles, self.global_step)
        queue = [Gaussians(*flat[4 * i:4 * i + 4]) for i in range(len(styles))]
        return type("Enc", (), {"stylized": stylized, "__call__": lambda s, c, st, gs=0: queue.pop(0)})()

    def __call__(self, batch: dict):
        enc = self._encoder_proxy(batch)
        self.optimizer.zero_grad(set_to_none=True)
        loss, logs = training_step(enc, self.decoder, self.losses, batch, self.global_step, self.identity_loss,
                                   self.data_shim)
        loss.backward()
        params = [p for g in self.optimizer.param_groups for p in g["params"]]
        if self.clip:
            torch.nn.utils.clip_grad_norm_(params, self.clip)
        self.optimizer.step()
        self.scheduler.step()
        self.global_step += 1
        return loss.detach(), logs
