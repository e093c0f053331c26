#!/usr/bin/env python
"""BENCH-ONLY library baseline (never imported by speakerguard_b200/): the same PGD iteration written as plain batched
PyTorch on the GPU, i.e. what the existing Blackwell library kernels give when the reference's per-utterance Python loops
are removed: cuFFT for the 512-point rFFT, cuDNN for the five dilated Conv1d layers and their input gradients (autograd,
weights frozen so no wgrad is computed), cuBLAS for mel / DCT / fc1 / LDA / PLDA.  `bench.py --impl torch_gpu` times it on
the headline workload; `tools/layer_table.py` uses the per-layer pieces.

Restates reference model/xv_plda.py:107-174 (-> torchaudio kaldi.mfcc, kaldi.py:514-813), model/iv_plda.py:296-377, :411-443,
model/_xv_plda/xvecTDNN.py:46-64, model/_xv_plda/plda.py:73-97, :140-190, attack/FGSM.py:38-70 - batched over utterances.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

EPS = 1.1920928955078125e-07
TDNN = [(30, 512, 5, 1), (512, 512, 5, 2), (512, 512, 7, 3), (512, 512, 1, 1), (512, 1500, 1, 1)]


class TorchGpuXv:
    def __init__(self, p: Dict[str, torch.Tensor], device, precision: str = "bf16"):
        self.dev, self.precision = torch.device(device), precision
        d = self.dev
        self.p = {k: v.to(d) for k, v in p.items()}
        j = torch.arange(400, dtype=torch.float64)
        self.window = (0.5 - 0.5 * torch.cos(2 * math.pi * j / 399)).pow(0.85).float().to(d)
        mel = lambda f: 1127.0 * torch.log(1.0 + f / 700.0)
        lo, hi = mel(torch.tensor(20.0, dtype=torch.float64)), mel(torch.tensor(7600.0, dtype=torch.float64))
        delta = (hi - lo) / 31
        c = torch.arange(30, dtype=torch.float64).unsqueeze(1)
        left, center, right = lo + c * delta, lo + (c + 1) * delta, lo + (c + 2) * delta
        mf = mel(31.25 * torch.arange(256, dtype=torch.float64)).unsqueeze(0)
        w = torch.clamp(torch.minimum((mf - left) / (center - left), (right - mf) / (right - center)), min=0)
        self.mel = F.pad(w, (0, 1)).float().to(d)                                  # [30, 257]
        n, k = torch.arange(30, dtype=torch.float64).unsqueeze(1), torch.arange(30, dtype=torch.float64).unsqueeze(0)
        dct = math.sqrt(2 / 30) * torch.cos(math.pi / 30 * (n + 0.5) * k)
        dct[:, 0] = math.sqrt(1 / 30)
        lift = 1 + 11 * torch.sin(math.pi * torch.arange(30, dtype=torch.float64) / 22)
        self.dct = (dct * lift).float().to(d)                                      # [n, k]
        dt = torch.bfloat16 if precision == "bf16" else torch.float32
        self.W = [self.p[f"tdnn{i}.weight"].to(dt) for i in range(1, 6)]
        self.b = [self.p[f"tdnn{i}.bias"].to(dt) for i in range(1, 6)]
        self.bn_m = [self.p[f"bn{i}.mean"].view(1, -1, 1).to(dt) for i in range(1, 6)]
        self.bn_s = [(self.p[f"bn{i}.var"] + 1e-5).rsqrt().view(1, -1, 1).to(dt) for i in range(1, 6)]
        self.dt = dt
        tf32 = precision != "fp32"
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.benchmark = True

    # ---- features -------------------------------------------------------------------------------
    def mfcc(self, x: torch.Tensor) -> torch.Tensor:
        B, N = x.shape
        m = (N + 80) // 160
        xs = x * 32768.0
        need = 160 * (m - 1) + 400 - 120 - N                                         # right reflection length
        xp = torch.cat([xs[:, :120].flip(1), xs, xs.flip(1)[:, :max(need, 0)]], 1)
        fr = xp.unfold(1, 400, 160)[:, :m]                                           # [B, m, 400]
        fr = fr + torch.randn_like(fr)                                               # dither 1.0 (kaldi.py:179-181)
        fr = fr - fr.mean(2, keepdim=True)
        logE = torch.log(torch.clamp(fr.pow(2).sum(2), min=EPS))
        prev = torch.cat([fr[:, :, :1], fr[:, :, :-1]], 2)
        g = (fr - 0.97 * prev) * self.window
        P = torch.fft.rfft(g, n=512).abs().pow(2)                                    # [B, m, 257]
        M = torch.log(torch.clamp(P @ self.mel.t(), min=EPS))
        cep = M @ self.dct
        return torch.cat([logE.unsqueeze(2), cep[:, :, 1:]], 2)

    @staticmethod
    def cmvn(feat: torch.Tensor) -> torch.Tensor:
        T = feat.shape[1]
        if T <= 300:
            return feat - feat.mean(1, keepdim=True)
        cs = F.pad(feat.cumsum(1), (0, 0, 1, 0))
        t = torch.arange(T, device=feat.device)
        ws = (t - 150).clamp(min=0)
        we = ws + 300
        over = (we - T).clamp(min=0)
        ws, we = (ws - over).clamp(min=0), we.clamp(max=T)
        return feat - (cs[:, we] - cs[:, ws]) / (we - ws).view(1, -1, 1)

    # ---- TDNN -----------------------------------------------------------------------------------
    def layer(self, i: int, h: torch.Tensor) -> torch.Tensor:
        ci, co, k, d = TDNN[i]
        return (F.relu(F.conv1d(h, self.W[i], self.b[i], dilation=d)) - self.bn_m[i]) * self.bn_s[i]

    def embed(self, feat: torch.Tensor) -> torch.Tensor:
        h = feat.transpose(1, 2).to(self.dt)
        for i in range(5):
            h = self.layer(i, h)
        h = h.float()
        stats = torch.cat([h.mean(2), h.std(2)], 1)
        p = self.p
        e = stats @ p["fc1.weight"].t() + p["fc1.bias"] - p["emb_mean"]
        e = e @ p["lda"][:, :512].t() + p["lda"][:, 512]
        L = e.shape[1]
        e = e * (math.sqrt(L) / e.norm(dim=1, keepdim=True).detach())               # quirk Q2: detached norm
        t = (e - p["plda.mean"]) @ p["plda.transform"].t()
        return t * torch.sqrt(L / (t.pow(2) / (p["plda.psi"] + 1)).sum(1, keepdim=True))

    def scores(self, q: torch.Tensor) -> torch.Tensor:
        p = self.p
        psi = p["plda.psi"]
        mean = (psi / (psi + 1)) * p["enroll"]                                       # [S, L]
        var = 1 + psi / (psi + 1)
        a = -0.5 * (torch.log(var).sum() + ((q.unsqueeze(1) - mean.unsqueeze(0)).pow(2) / var).sum(2))
        b = -0.5 * (torch.log(psi + 1).sum() + (q.pow(2) / (psi + 1)).sum(1, keepdim=True))
        return a - b

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.scores(self.embed(self.cmvn(self.mfcc(x))))

    # ---- attack ---------------------------------------------------------------------------------
    def pgd(self, x0: torch.Tensor, y: torch.Tensor, max_iter: int, eps: float = 0.002, step: float = 0.0004):
        """attack/FGSM.py:38-70 with EOT_size 1: max_iter gradient passes + the evaluation pass."""
        lower, upper = (x0 - eps).clamp(min=-1), (x0 + eps).clamp(max=1)
        x = x0.clone()
        for _ in range(max_iter):
            xr = x.detach().requires_grad_(True)
            loss = F.cross_entropy(self.forward(xr), y, reduction="sum")
            (g,) = torch.autograd.grad(loss, xr)
            x = torch.minimum(torch.maximum(x + step * g.sign(), lower), upper)
        with torch.no_grad():
            dec = self.forward(x).argmax(1)
        return x, dec
