"""Decision losses of the attacks, computed by libsgb200's loss kernel (reference
attack/utils.py:7-125: SEC4SR_CrossEntropy, SEC4SR_MarginLoss, resolve_loss, resolve_prediction).

The modules keep the reference's names and constructor arguments; forward() runs
``sg_loss_fwd_bwd`` and registers its analytic gradient with autograd, so
``loss.backward(torch.ones_like(loss))`` works as in adaptive_attack/EOT.py:35 without the
reference's host-side index lists (no device->host sync)."""
from collections import Counter

import numpy as np
import torch
import torch.nn as nn

from ..engine import default_engine, make_loss_params


class _LossFn(torch.autograd.Function):

    @staticmethod
    def forward(ctx, scores, label, lp):
        eng = default_engine(scores.device)
        loss, ds = eng.loss(scores, label, lp, want_grad=True)
        ctx.save_for_backward(ds)
        return loss

    @staticmethod
    def backward(ctx, g):
        (ds,) = ctx.saved_tensors
        return ds * g.unsqueeze(1), None, None


class SEC4SR_CrossEntropy(nn.Module):
    """Per-utterance cross entropy over the enrolled speakers; label -1 (imposter) -> 0 loss."""

    def __init__(self, weight=None, size_average=None, ignore_index=-100, reduce=None, reduction="none", task="CSI"):
        super().__init__()
        assert task == "CSI", "CrossEntropy only supports the CSI task"
        assert reduction == "none", "the attacks use per-utterance losses (reduction='none')"
        self.lp = make_loss_params("Entropy", False, "CSI")

    def forward(self, scores, label):
        return _LossFn.apply(scores, label, self.lp)


class SEC4SR_MarginLoss(nn.Module):
    """CW-style margin loss for CSI / SV / OSI, targeted or not, optionally clipped at 0."""

    def __init__(self, targeted=False, confidence=0., task="CSI", threshold=None, clip_max=True):
        super().__init__()
        self.targeted, self.confidence, self.task = targeted, confidence, task
        self.threshold, self.clip_max = threshold, clip_max
        self.lp = make_loss_params("Margin", targeted, task, confidence, threshold, clip_max)

    def forward(self, scores, label):
        return _LossFn.apply(scores, label, self.lp)


def resolve_loss(loss_name="Entropy", targeted=False, confidence=0., task="CSI", threshold=None, clip_max=True):
    """-> (loss module, grad_sign).  SV/OSI always use the margin loss; grad_sign is +1 for
    untargeted cross entropy (ascend), -1 for targeted cross entropy and for any margin loss."""
    assert loss_name in ["Entropy", "Margin"]
    assert task in ["CSI", "SV", "OSI"]
    if task in ("SV", "OSI") or loss_name == "Margin":
        loss = SEC4SR_MarginLoss(targeted=targeted, confidence=confidence, task=task, threshold=threshold,
                                 clip_max=clip_max)
    else:
        loss = SEC4SR_CrossEntropy(reduction="none", task="CSI")
    grad_sign = (1 - 2 * int(targeted)) if loss_name == "Entropy" else -1
    return loss, grad_sign


def resolve_prediction(decisions):
    """Mode of the per-EOT-sample decisions of every utterance (list[list[int]] -> np.ndarray)."""
    return np.array([Counter(d).most_common(1)[0][0] for d in decisions])
