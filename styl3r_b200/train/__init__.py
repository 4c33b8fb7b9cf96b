"""Training-side callers of the hot path: VGG feature encoder, style / identity losses,
optimiser set-up and the training step of ModelWrapperStyle, behind the reference's names."""
from .losses import IdentityLoss, LossStyle, LossStyleCfg, LossStyleCfgWrapper
from .step import TrainStep, configure_optimizers, select_trainable, training_step
from .vgg import VGGEncoder, calc_mean_std

__all__ = ["IdentityLoss", "LossStyle", "LossStyleCfg", "LossStyleCfgWrapper", "TrainStep", "configure_optimizers",
           "select_trainable", "training_step", "VGGEncoder", "calc_mean_std"]
