"""Expectation-over-transformation wrapper (reference adaptive_attack/EOT.py:16-54).

Generic path: works with any model exposing ``make_decision`` (including defended models with
randomised feature-level defenses).  Scores / loss / gradient are the sum over EOT batches of
the mean within each batch (the caller divides by the number of batches, attack/FGSM.py:50-53).
Differences from the reference that do not change results: the repeated input is a leaf tensor
whose ``.grad`` is read once per EOT batch, decisions are gathered on the device and copied to
the host once per call, and ``use_grad=False`` really skips the backward pass (reference quirk Q1).
"""
import torch
import torch.nn as nn


class EOT(nn.Module):

    def __init__(self, model, loss, EOT_size=1, EOT_batch_size=1, use_grad=True):
        super().__init__()
        self.model = model
        self.loss = loss
        self.EOT_size = EOT_size
        self.EOT_batch_size = EOT_batch_size
        self.EOT_num_batches = self.EOT_size // self.EOT_batch_size
        self.use_grad = use_grad

    def forward(self, x_batch, y_batch, EOT_num_batches=None, EOT_batch_size=None, use_grad=None):
        EOT_num_batches = EOT_num_batches if EOT_num_batches else self.EOT_num_batches
        EOT_batch_size = EOT_batch_size if EOT_batch_size else self.EOT_batch_size
        use_grad = self.use_grad if use_grad is None else use_grad
        n_audios, n_channels, max_len = x_batch.size()
        grad, scores, loss = None, None, None
        all_dec = []
        for _ in range(EOT_num_batches):
            xr = x_batch.detach().repeat(EOT_batch_size, 1, 1).requires_grad_(use_grad)
            yr = y_batch.repeat(EOT_batch_size)
            with torch.set_grad_enabled(use_grad):
                dec, sc = self.model.make_decision(xr)
                ls = self.loss(sc, yr)
            if use_grad:
                ls.backward(torch.ones_like(ls))
                g = xr.grad.view(EOT_batch_size, n_audios, n_channels, max_len).mean(0)
                grad = g if grad is None else grad + g
            s = sc.detach().view(EOT_batch_size, n_audios, -1).mean(0)
            l = ls.detach().view(EOT_batch_size, n_audios).mean(0)
            scores = s if scores is None else scores + s
            loss = l if loss is None else loss + l
            all_dec.append(dec.detach().view(EOT_batch_size, n_audios))
        dec_host = torch.cat(all_dec, 0).cpu().numpy()          # one D2H copy per call
        decisions = [list(dec_host[:, i]) for i in range(n_audios)]
        return scores, loss, grad, decisions
