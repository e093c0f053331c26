"""Expectation-over-transformation wrapper (reference adaptive_attack/EOT.py:16-54).

Generic path: works with any model exposing ``make_decision`` (including defended models with
randomised feature-level defenses).  Scores / loss / gradient are the sum over EOT batches of
the mean within each batch (the caller divides by the number of batches, attack/FGSM.py:50-53).
Differences from the reference that do not change results: the input gradient is taken with
``torch.autograd.grad`` on the repeated input (one read per EOT batch, no parameter gradients), decisions are gathered on the device and copied to
the host once per call, and ``use_grad=False`` really skips the backward pass (reference quirk Q1).
"""
import torch
import torch.nn as nn

from ..functional import input_grad_only


class EOT(nn.Module):

    def __init__(self, model, loss, EOT_size=1, EOT_batch_size=1, use_grad=True):
        super().__init__()
        self.model = model
        self.loss = loss
        self.EOT_size = EOT_size
        self.EOT_batch_size = EOT_batch_size
        self.EOT_num_batches = self.EOT_size // self.EOT_batch_size
        self.use_grad = use_grad

    def forward(self, x_batch, y_batch, EOT_num_batches=None, EOT_batch_size=None, use_grad=None, need_decisions=True):
        rounds = EOT_num_batches or self.EOT_num_batches
        copies = EOT_batch_size or self.EOT_batch_size
        want_grad = self.use_grad if use_grad is None else use_grad
        n = x_batch.shape[0]
        totals = {"scores": None, "loss": None, "grad": None}
        votes = []

        def accumulate(key, value):
            totals[key] = value if totals[key] is None else totals[key] + value

        for _ in range(rounds):
            tiled = x_batch.detach().repeat(copies, 1, 1).requires_grad_(want_grad)     # a leaf: its .grad is read once
            labels = y_batch.repeat(copies)
            with torch.set_grad_enabled(want_grad):
                decided, sc = self.model.make_decision(tiled)
                per_copy_loss = self.loss(sc, labels)
            if want_grad:
                # only d(loss)/d(input) is needed here: parameter gradients of a trainable model (adver_train.py runs the
                # attack on the train-mode network) are not formed, and nothing accumulates into their .grad
                with input_grad_only():
                    (g,) = torch.autograd.grad(per_copy_loss, tiled, torch.ones_like(per_copy_loss))
                accumulate("grad", g.view(copies, *x_batch.shape).mean(0))
            accumulate("scores", sc.detach().view(copies, n, -1).mean(0))
            accumulate("loss", per_copy_loss.detach().view(copies, n).mean(0))
            votes.append(decided.detach().view(copies, n))
        if not need_decisions:            # intermediate attack iterations: no device->host copy, the host keeps enqueueing
            return totals["scores"], totals["loss"], totals["grad"], None
        ballot = torch.cat(votes, 0).cpu().numpy()                                     # one device->host copy per call
        decisions = [list(ballot[:, i]) for i in range(n)]
        return totals["scores"], totals["loss"], totals["grad"], decisions
