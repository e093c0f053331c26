"""CPU oracle for the SpeakerGuard attack-iteration hot path (TEST INFRASTRUCTURE ONLY).

This module restates, in plain batched torch-CPU fp32 (fp64 on request), the arithmetic of the
reference's hot path.  It is the *checker*: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  Nothing under
``speakerguard_b200/`` imports it, and the product never falls back to it.

Parity status: the reference ships no tests / golden vectors, so the oracle
is pinned against outputs of the reference itself, generated in the build container by
``tests/golden/make_golden.py`` (imports /root/reference with the SURVEY 8(c) shims) and committed
as ``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks this file against them.
Exceptions (stated in DESIGN.md): the AudioNet mel matrix (librosa 0.8 absent -> Slaney
restatement, "parity unpinned" against real librosa) and FeCo's k-means ids (libKMCUDA /
kmeans_pytorch absent and randomised -> conditional parity only).

Citations are relative to /root/reference unless they start with ``kaldi.py`` (torchaudio
2.11 ``torchaudio/compliance/kaldi.py``, the third-party file holding the MFCC arithmetic).
"""
from __future__ import annotations

import math
from collections import Counter
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

EPS = float(torch.finfo(torch.float32).eps)  # kaldi.py:22
WIN_SHIFT, WIN_SIZE, WIN_PAD = 160, 400, 512  # kaldi.py:125-151 with 16 kHz / 10 ms / 25 ms
NUM_MEL, NUM_CEPS = 30, 30                     # model/xv_plda.py:133,141
TDNN_SPEC = [  # (C_in, C_out, kernel, dilation)  model/_xv_plda/xvecTDNN.py:16-34
    (30, 512, 5, 1), (512, 512, 5, 2), (512, 512, 7, 3), (512, 512, 1, 1), (512, 1500, 1, 1)]
BN_EPS = 1e-5
_torch_stft = torch.stft  # bound at import: the golden generator monkey-patches torch.stft later


# ----------------------------------------------------------------------------------------------
# synthetic model: random-init xvecTDNN + synthetic PLDA / LDA / enrol files
# ----------------------------------------------------------------------------------------------
def _round6(t: torch.Tensor) -> torch.Tensor:
    """Values as they come back from the reference's text parsers after being written with
    '%.6f' (model/_xv_plda/plda.py:27-51, model/utils.py:50-80): float(text) -> float32."""
    flat = [float("%.6f" % v) for v in t.double().flatten().tolist()]
    return torch.tensor(flat, dtype=torch.float32).reshape(t.shape)


def make_xv_params(seed: int = 0, L: int = 200, S: int = 10, num_spk: int = 100,
                   random_bn: bool = True) -> Dict[str, torch.Tensor]:
    """Random-init x-vector system.  Consumes the torch CPU generator exactly like
    ``torch.manual_seed(seed); xvecTDNN(numSpkrs=num_spk)`` (xvecTDNN.py:14-43: conv1..5, fc1,
    fc2, fc3 in that order), then draws the auxiliary tensors from a second generator."""
    torch.manual_seed(seed)
    p: Dict[str, torch.Tensor] = {}
    for i, (ci, co, k, d) in enumerate(TDNN_SPEC, 1):
        conv = torch.nn.Conv1d(ci, co, k, dilation=d)
        p[f"tdnn{i}.weight"] = conv.weight.detach().clone()
        p[f"tdnn{i}.bias"] = conv.bias.detach().clone()
    fc1 = torch.nn.Linear(3000, 512)
    torch.nn.Linear(512, 512)       # fc2: unused by embedding(), consumes RNG (xvecTDNN.py:40)
    torch.nn.Linear(512, num_spk)   # fc3
    p["fc1.weight"] = fc1.weight.detach().clone()
    p["fc1.bias"] = fc1.bias.detach().clone()
    g = torch.Generator().manual_seed(seed + 7919)
    for i, (_, co, _, _) in enumerate(TDNN_SPEC, 1):
        if random_bn:
            p[f"bn{i}.mean"] = 0.1 * torch.randn(co, generator=g)
            p[f"bn{i}.var"] = 0.5 + torch.rand(co, generator=g)
        else:
            p[f"bn{i}.mean"] = torch.zeros(co)
            p[f"bn{i}.var"] = torch.ones(co)
    p["plda.mean"] = _round6(0.1 * torch.randn(L, generator=g))
    p["plda.transform"] = _round6(torch.randn(L, L, generator=g) / math.sqrt(L))
    p["plda.psi"] = _round6(torch.randn(L, generator=g).abs() + 0.1)
    p["emb_mean"] = _round6(0.1 * torch.randn(512, generator=g))
    p["lda"] = _round6(torch.randn(L, 513, generator=g) / math.sqrt(512))
    p["enroll"] = torch.randn(S, L, generator=g)
    return p


def params_checksum(p: Dict[str, torch.Tensor]) -> float:
    return float(sum(v.double().abs().sum() for v in p.values()))


# ----------------------------------------------------------------------------------------------
# Kaldi MFCC (model/xv_plda.py:107-156 -> kaldi.py:669-813)
# ----------------------------------------------------------------------------------------------
def num_frames(N: int) -> int:
    return (N + WIN_SHIFT // 2) // WIN_SHIFT  # kaldi.py:70 (snip_edges=False)


def frame_index(N: int) -> torch.Tensor:
    """[m,400] gather indices into the waveform: reflect padding of kaldi.py:69-77."""
    m = num_frames(N)
    pidx = (WIN_SHIFT * torch.arange(m).unsqueeze(1) + torch.arange(WIN_SIZE).unsqueeze(0)
            - (WIN_SIZE // 2 - WIN_SHIFT // 2))
    pidx = torch.where(pidx < 0, -pidx - 1, pidx)
    pidx = torch.where(pidx >= N, 2 * N - 1 - pidx, pidx)
    return pidx


def povey_window(dtype=torch.float32) -> torch.Tensor:
    return torch.hann_window(WIN_SIZE, periodic=False, dtype=dtype).pow(0.85)  # kaldi.py:98-100


def mel_banks(num_bins: int = NUM_MEL, low: float = 20.0, high: float = 7600.0,
              sr: float = 16000.0, npad: int = WIN_PAD) -> torch.Tensor:
    """[num_bins, npad/2+1] triangular filters, kaldi.py:436-511 (vtln_warp == 1) with the zero
    Nyquist column appended as in kaldi.py:627."""
    mel = lambda f: 1127.0 * math.log(1.0 + f / 700.0)
    ml, mh = mel(low), mel(high)
    delta = (mh - ml) / (num_bins + 1)
    b = torch.arange(num_bins).unsqueeze(1)
    left, center, right = ml + b * delta, ml + (b + 1.0) * delta, ml + (b + 2.0) * delta
    freqs = (sr / npad) * torch.arange(npad / 2)
    melf = (1127.0 * (1.0 + freqs / 700.0).log()).unsqueeze(0)
    up = (melf - left) / (center - left)
    down = (right - melf) / (right - center)
    bins = torch.max(torch.zeros(1), torch.min(up, down))
    return F.pad(bins, (0, 1))


def dct_matrix(n: int = NUM_MEL, nceps: int = NUM_CEPS) -> torch.Tensor:
    """[n_mel, n_ceps]; kaldi.py:648-658 over torchaudio.functional.create_dct(norm='ortho')."""
    nn_ = torch.arange(float(n))
    k = torch.arange(float(n)).unsqueeze(1)
    dct = torch.cos(math.pi / float(n) * (nn_ + 0.5) * k)
    dct[0] *= 1.0 / math.sqrt(2.0)
    dct *= math.sqrt(2.0 / float(n))
    dct = dct.t().contiguous()
    dct[:, 0] = math.sqrt(1 / float(n))
    return dct[:, :nceps]


def lifter(nceps: int = NUM_CEPS, q: float = 22.0) -> torch.Tensor:
    i = torch.arange(nceps)
    return 1.0 + 0.5 * q * torch.sin(math.pi * i / q)  # kaldi.py:661-666


def mfcc(x: torch.Tensor, dither: Optional[torch.Tensor] = None, scale: float = 32768.0,
         num_ceps: int = NUM_CEPS) -> torch.Tensor:
    """x [B,N] in [-1,1] -> raw MFCC [B,m,num_ceps].  ``scale`` is check_input_range's 2**15
    (model/utils.py:14).  ``dither`` [B,m,400] is the N(0,1) tensor kaldi.py:180 would draw
    (dither = 1.0, model/xv_plda.py:119); None = no dither."""
    dt = x.dtype
    B, N = x.shape
    f = (x * scale)[:, frame_index(N)]                       # [B,m,400]  kaldi.py:176
    if dither is not None:
        f = f + dither.to(dt)                                # kaldi.py:179-181
    f = f - f.mean(dim=2, keepdim=True)                      # kaldi.py:183-186
    eps = torch.tensor(EPS, dtype=dt)
    log_e = torch.max(f.pow(2).sum(2), eps).log()            # kaldi.py:116-122, 188-191
    prev = torch.cat([f[:, :, :1], f[:, :, :-1]], dim=2)     # replicate pad, kaldi.py:193-198
    g = (f - 0.97 * prev) * povey_window(dt)                 # kaldi.py:200-204
    g = F.pad(g, (0, WIN_PAD - WIN_SIZE))                    # kaldi.py:206-211
    spec = torch.fft.rfft(g).abs().pow(2.0)                  # kaldi.py:616-618
    mel = spec @ mel_banks().to(dt).T                        # kaldi.py:621-630
    mel = torch.max(mel, eps).log()                          # kaldi.py:631-633
    c = (mel @ dct_matrix().to(dt)[:, :num_ceps]) * lifter(num_ceps).to(dt)  # kaldi.py:788-796
    c = torch.cat([log_e.unsqueeze(2), c[:, :, 1:]], dim=2)  # kaldi.py:799-800
    return c


def cmvn_windows(T: int, win: int = 300) -> Tuple[torch.Tensor, torch.Tensor]:
    """Window [ws, we) for every frame: model/iv_plda.py:321-337 (center=True)."""
    t = torch.arange(T)
    ws = t - win // 2
    we = ws + win
    shift = torch.clamp(-ws, min=0)
    ws, we = ws + shift, we + shift
    over = torch.clamp(we - T, min=0)
    ws, we = torch.clamp(ws - over, min=0), we - over
    return ws, we


def cmvn(feat: torch.Tensor) -> torch.Tensor:
    """Sliding mean-only CMVN, model/iv_plda.py:296-377, closed form."""
    B, T, Fd = feat.shape
    ws, we = cmvn_windows(T)
    cs = torch.cat([torch.zeros(B, 1, Fd, dtype=torch.float64), feat.double().cumsum(1)], dim=1)
    mean = (cs[:, we] - cs[:, ws]) / (we - ws).double().view(1, T, 1)
    return (feat.double() - mean).to(feat.dtype)


def cmvn_loop(feat: torch.Tensor) -> torch.Tensor:
    """Literal running-sum restatement of model/iv_plda.py:319-366 (slow; pins the closed form)."""
    out = []
    for x in feat:
        T = x.shape[0]
        last_s = last_e = -1
        cur = torch.zeros(x.shape[1], dtype=x.dtype)
        y = x.clone()
        for t in range(T):
            s = t - 150
            e = s + 300
            if s < 0:
                e -= s
                s = 0
            if e > T:
                s -= e - T
                e = T
                if s < 0:
                    s = 0
            if last_s == -1:
                cur = x[s:e].sum(0)
            else:
                if s > last_s:
                    cur = cur - x[last_s]
                if e > last_e:
                    cur = cur + x[last_e]
            last_s, last_e = s, e
            y[t] = x[t] - cur / (e - s)
        out.append(y)
    return torch.stack(out)


# ----------------------------------------------------------------------------------------------
# x-vector TDNN + head (xvecTDNN.py:46-64, iv_plda.py:411-443, plda.py:73-97,140-190)
# ----------------------------------------------------------------------------------------------
def tdnn_layers(feat: torch.Tensor, p: Dict[str, torch.Tensor], flips=None, preacts: Optional[list] = None
                ) -> List[torch.Tensor]:
    """feat [B,T,30] -> list of the five post-BN activations [B,C,T_l].

    ``flips`` (test aid): per-layer bool tensors; where True the ReLU gate of that unit is inverted
    (the gradient of the network is discontinuous where a pre-activation is ~0, and two fp32
    summation orders can land on different sides; see tests/kink.py).  ``preacts`` collects the
    pre-activations."""
    x = feat.transpose(1, 2)
    outs = []
    for i, (_, _, _, d) in enumerate(TDNN_SPEC, 1):
        a = F.conv1d(x, p[f"tdnn{i}.weight"].to(x.dtype), p[f"tdnn{i}.bias"].to(x.dtype), dilation=d)
        if preacts is not None:
            preacts.append(a.detach())
        if flips is not None and flips[i - 1] is not None:
            r = a * ((a > 0) ^ flips[i - 1]).to(a.dtype)
        else:
            r = F.relu(a)                                       # ReLU precedes BN (xvecTDNN.py:49)
        x = (r - p[f"bn{i}.mean"].to(x.dtype).view(1, -1, 1)) / torch.sqrt(
            p[f"bn{i}.var"].to(x.dtype).view(1, -1, 1) + BN_EPS)
        outs.append(x)
    return outs


def xvector(feat: torch.Tensor, p: Dict[str, torch.Tensor], flips=None, preacts=None) -> torch.Tensor:
    x = tdnn_layers(feat, p, flips, preacts)[-1]
    stats = torch.cat((x.mean(dim=2), x.std(dim=2)), dim=1)     # xvecTDNN.py:62 (unbiased std)
    return stats @ p["fc1.weight"].to(x.dtype).T + p["fc1.bias"].to(x.dtype)


def process_emb(e: torch.Tensor, p: Dict[str, torch.Tensor]) -> torch.Tensor:
    dt = e.dtype
    L = p["plda.mean"].shape[0]
    e = e - p["emb_mean"].to(dt)                                 # xvector_extract.py:41-43
    A = p["lda"].to(dt)
    e = e @ A[:, :-1].T + A[:, -1]                               # iv_plda.py:423-435
    nrm = e.detach().norm(dim=1, keepdim=True)                   # .item(): detached, Q2
    e = e * (math.sqrt(L) / nrm)                                 # xvector_extract.py:31-38
    t = (e - p["plda.mean"].to(dt)) @ p["plda.transform"].to(dt).T   # plda.py:75
    inv = 1.0 / (p["plda.psi"].to(dt) + 1.0)
    factor = torch.sqrt(L / (t.pow(2) * inv).sum(1, keepdim=True))   # plda.py:92-97
    return t * factor


def plda_scores(q: torch.Tensor, p: Dict[str, torch.Tensor], enroll: Optional[torch.Tensor] = None
                ) -> torch.Tensor:
    """[B,L] test embeddings vs [S,L] enrolled -> LLR [B,S]; plda.py:140-190 term by term."""
    dt = q.dtype
    psi = p["plda.psi"].to(dt)
    en = (p["enroll"] if enroll is None else enroll).to(dt)
    L = psi.shape[0]
    mean = psi / (psi + 1.0) * en                                # [S,L]
    var = 1.0 + psi / (psi + 1.0)
    logdet = torch.log(var).sum() * torch.ones(en.shape[0], dtype=dt)
    c = torch.log(2 * torch.tensor(3.1415926, dtype=dt)) * L
    sq = (q.unsqueeze(1) - mean.unsqueeze(0)).pow(2)             # [B,S,L]
    given = -0.5 * (logdet + c + (sq * (1.0 / var)).sum(2))
    var2 = psi + 1.0
    without = -0.5 * (torch.log(var2).sum() + c + (q.pow(2) * (1.0 / var2)).sum(1))
    return given - without.unsqueeze(1)


def decide(scores: torch.Tensor, threshold: float = -math.inf) -> torch.Tensor:
    d = scores.argmax(1)                                          # defended_model.py:167-170
    return torch.where(scores.max(1)[0] > threshold, d, torch.full_like(d, -1))


def xv_forward(x: torch.Tensor, p: Dict[str, torch.Tensor], dither: Optional[torch.Tensor] = None,
               return_all: bool = False, flips=None, preacts=None):
    """x [B,N] (or [B,1,N]) -> scores [B,S] (iv_plda.py:155-169 with flag 0)."""
    if x.dim() == 3:
        x = x[:, 0]
    raw = mfcc(x, dither)
    feat = cmvn(raw)
    emb = process_emb(xvector(feat, p, flips, preacts), p)
    scores = plda_scores(emb, p)
    if return_all:
        return {"raw": raw, "feat": feat, "emb": emb, "scores": scores}
    return scores


# ----------------------------------------------------------------------------------------------
# losses (attack/utils.py:7-116)
# ----------------------------------------------------------------------------------------------
def loss_ce(scores: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    loss = torch.zeros(y.shape[0], dtype=scores.dtype)
    keep = y != -1
    if keep.any():
        loss = torch.where(keep, F.cross_entropy(scores, y.clamp(min=0), reduction="none"), loss)
    return loss + 0.0 * scores.sum(1) * (~keep)                  # attack/utils.py:25-27


def loss_margin(scores: torch.Tensor, y: torch.Tensor, targeted: bool = False, confidence: float = 0.0,
                task: str = "CSI", threshold: Optional[float] = None, clip_max: bool = True) -> torch.Tensor:
    B, S = scores.shape
    loss = torch.zeros(B, dtype=scores.dtype)
    if task == "SV":                                              # attack/utils.py:48-61
        s = scores[:, 0]
        enr = y == 0
        a = threshold + confidence - s
        b = s + confidence - threshold
        loss = torch.where(enr, a if targeted else b, b if targeted else a)
    else:                                                         # attack/utils.py:63-94
        keep = y != -1
        oh = F.one_hot(y.clamp(min=0), S).to(scores.dtype)
        real = (oh * scores).sum(1)
        other = ((1 - oh) * scores - oh * 10000).max(1)[0]
        if targeted:
            v = other + confidence - real if task == "CSI" else \
                torch.clamp(other, min=threshold) + confidence - real
        elif task == "CSI":
            v = real + confidence - other
        else:
            f_rej = scores.max(1)[0] + confidence - threshold
            f_mis = torch.clamp(real, min=threshold) + confidence - other
            v = torch.minimum(f_rej, f_mis)
        if task == "OSI":
            mx = scores.max(1)[0]
            imp = mx + confidence - threshold if targeted else threshold + confidence - mx
        else:
            imp = 0.0 * scores.sum(1)
        loss = torch.where(keep, v, imp)
    if clip_max:
        loss = torch.clamp(loss, min=0)
    return loss


def resolve_loss(loss_name="Entropy", targeted=False, confidence=0.0, task="CSI", threshold=None,
                 clip_max=True):
    """attack/utils.py:104-116 -> (callable(scores, y), grad_sign)."""
    if task in ("SV", "OSI") or loss_name == "Margin":
        fn = lambda s, y: loss_margin(s, y, targeted, confidence, task, threshold, clip_max)
    else:
        fn = loss_ce
    grad_sign = (1 - 2 * int(targeted)) if loss_name == "Entropy" else -1
    return fn, grad_sign


# ----------------------------------------------------------------------------------------------
# FGSM / PGD (attack/FGSM.py:38-98, attack/PGD.py:40-79, adaptive_attack/EOT.py:16-54, EOT_size 1)
# ----------------------------------------------------------------------------------------------
def xv_loss_and_grad(x: torch.Tensor, y: torch.Tensor, p, loss_fn, dither=None, flips=None, preacts=None, system="xv"):
    """One EOT pass with E=1: scores, loss, d(sum loss)/dx, decisions.  system: 'xv' (x-vector) | 'iv' (i-vector)."""
    xr = x.detach().clone().requires_grad_(True)
    scores = iv_forward(xr, p, dither) if system == "iv" else xv_forward(xr, p, dither, flips=flips, preacts=preacts)
    loss = loss_fn(scores, y)
    loss.backward(torch.ones_like(loss))                          # EOT.py:35
    thr = p.get("threshold", -math.inf)
    return scores.detach(), loss.detach(), xr.grad.detach(), decide(scores.detach(), thr)


def pgd_attack(x: torch.Tensor, y: torch.Tensor, p, epsilon=0.002, step_size=0.0004, max_iter=10,
               loss_name="Entropy", targeted=False, task="CSI", dither: Optional[torch.Tensor] = None,
               fgsm: bool = False, x_init: Optional[torch.Tensor] = None, system: str = "xv"):
    """x [B,N].  dither [max_iter+1,B,m,400] or None.  Returns (adv [B,N], success list, info).
    FGSM = one step of size epsilon with [-1,1] bounds (attack/FGSM.py:35-36,74-81)."""
    thr = p.get("threshold", None)
    loss_fn, grad_sign = resolve_loss(loss_name, targeted, 0.0, task,
                                      thr if task in ("SV", "OSI") else None, clip_max=False)
    if fgsm:
        lower, upper = torch.full_like(x, -1.0), torch.full_like(x, 1.0)
        step_size, max_iter = epsilon, 1
    else:
        upper = torch.clamp(x + epsilon, max=1.0)                 # attack/PGD.py:48-49
        lower = torch.clamp(x - epsilon, min=-1.0)
    xa = (x if x_init is None else x_init).clone()
    info = {"loss": [], "decisions": []}
    success = None
    for it in range(max_iter + 1):
        d = None if dither is None else dither[it]
        scores, loss, grad, dec = xv_loss_and_grad(xa, y, p, loss_fn, d, system=system)
        info["loss"].append(loss)
        info["decisions"].append(dec)
        success = (dec == y).tolist() if targeted else (dec != y).tolist()   # Attack.py:11-15
        if it < max_iter:
            xa = xa + step_size * torch.sign(grad) * grad_sign    # attack/FGSM.py:65
            xa = torch.min(torch.max(xa, lower), upper)           # attack/FGSM.py:68
    info["scores"] = scores
    return xa, success, info


def resolve_prediction(decisions: List[List[int]]) -> np.ndarray:
    return np.array([Counter(d).most_common(1)[0][0] for d in decisions])   # attack/utils.py:118-125


# ----------------------------------------------------------------------------------------------
# FeCo conditional restatement (defense/feature_level.py:202-217): means given cluster ids
# ----------------------------------------------------------------------------------------------
def feco_means(feat: torch.Tensor, ids: np.ndarray, k: int, force: bool = True) -> torch.Tensor:
    """feat [n,dim], ids [n] in [0,k) -> [k,dim]; empty cluster i -> feat[i] when force."""
    rows = []
    for i in range(k):
        sel = np.argwhere(ids == i).flatten()
        if sel.size > 0:
            rows.append(feat[sel, :].mean(0, keepdim=True))
        elif force:
            rows.append(feat[i:i + 1, :])
    return torch.cat(rows, 0)


def lloyd_inertia(feat: np.ndarray, ids: np.ndarray, k: int) -> float:
    tot = 0.0
    for i in range(k):
        sel = feat[ids == i]
        if len(sel):
            tot += float(((sel - sel.mean(0)) ** 2).sum())
    return tot


# ----------------------------------------------------------------------------------------------
# AudioNet (model/audionet_csine.py:133-257, model/_audionet/Preprocessor.py:85-112)
# ----------------------------------------------------------------------------------------------
AN_NFFT, AN_HOP, AN_WIN, AN_MELS = 1024, 160, 800, 32
AN_CONVS = [  # (name, C_in, C_out, k, pad, pool)   audionet_csine.py:66-118
    ("conv2", 32, 64, 3, 1, True), ("conv3", 64, 128, 3, 1, False), ("conv4", 128, 128, 3, 1, False),
    ("conv5", 128, 128, 3, 1, True), ("conv6", 128, 128, 3, 1, False), ("conv7", 128, 64, 3, 1, True),
    ("conv8", 64, 32, 3, 0, False)]


def slaney_mel(sr=16000, n_fft=AN_NFFT, n_mels=AN_MELS, fmin=0.0, fmax=8000.0) -> torch.Tensor:
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax) restated (Slaney scale + area norm);
    returns [n_mels, n_fft/2+1].  librosa itself is absent: parity unpinned."""
    def hz2mel(f):
        f = np.asarray(f, dtype=np.float64)
        m = f / (200.0 / 3)
        lin = f >= 1000.0
        return np.where(lin, 15.0 + np.log(np.maximum(f, 1e-10) / 1000.0) / (np.log(6.4) / 27.0), m)

    def mel2hz(m):
        m = np.asarray(m, dtype=np.float64)
        return np.where(m >= 15.0, 1000.0 * np.exp((np.log(6.4) / 27.0) * (m - 15.0)), m * 200.0 / 3)
    fft_f = np.linspace(0, sr / 2, n_fft // 2 + 1)
    mel_f = mel2hz(np.linspace(hz2mel(fmin), hz2mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fft_f[None, :]
    w = np.zeros((n_mels, n_fft // 2 + 1))
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0, np.minimum(lower, upper))
    w *= (2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels]))[:, None]
    return torch.tensor(w, dtype=torch.float32)


def make_audionet_params(seed: int = 0, num_class: int = 251, random_bn: bool = True):
    """Random-init AudioNet in the construction order of audionet_csine.py:58-121."""
    torch.manual_seed(seed)
    p = {}
    c1 = torch.nn.Conv2d(1, 1, kernel_size=[5, 5], stride=1, padding=[2, 2])
    p["conv1.weight"], p["conv1.bias"] = c1.weight.detach().clone(), c1.bias.detach().clone()
    for name, ci, co, k, pad, _ in AN_CONVS:
        c = torch.nn.Conv1d(ci, co, k, padding=pad)
        p[f"{name}.weight"], p[f"{name}.bias"] = c.weight.detach().clone(), c.bias.detach().clone()
    fc = torch.nn.Linear(32, num_class)
    p["fc.weight"], p["fc.bias"] = fc.weight.detach().clone(), fc.bias.detach().clone()
    g = torch.Generator().manual_seed(seed + 104729)
    for name, co in [("conv1", 1)] + [(n, c) for n, _, c, _, _, _ in AN_CONVS]:
        if random_bn:
            p[f"{name}.bn_mean"] = 0.1 * torch.randn(co, generator=g)
            p[f"{name}.bn_var"] = 0.5 + torch.rand(co, generator=g)
            p[f"{name}.bn_gamma"] = 1.0 + 0.1 * torch.randn(co, generator=g)
            p[f"{name}.bn_beta"] = 0.1 * torch.randn(co, generator=g)
        else:
            p[f"{name}.bn_mean"], p[f"{name}.bn_var"] = torch.zeros(co), torch.ones(co)
            p[f"{name}.bn_gamma"], p[f"{name}.bn_beta"] = torch.ones(co), torch.zeros(co)
    return p


def audionet_logmel(x: torch.Tensor) -> torch.Tensor:
    """x [B,N] in [-1,1] -> [B,32,T]; Preprocessor.py:88-112."""
    dt = x.dtype
    w = x[:, 1:] - 0.97 * x[:, :-1]                                # Preprocessor.py:85-86
    spec = _torch_stft(w, n_fft=AN_NFFT, hop_length=AN_HOP, win_length=AN_WIN,
                      window=torch.hann_window(AN_WIN, dtype=dt), return_complex=True)
    power = spec.real.pow(2) + spec.imag.pow(2)                    # Preprocessor.py:28-37
    mel = torch.matmul(power.transpose(2, 1), slaney_mel().to(dt).T).transpose(2, 1)
    return 10 * torch.clamp(mel, 1e-16).log10()                    # Preprocessor.py:111


def _bn(x, p, name, dims):
    shape = [1, -1] + [1] * dims
    dt = x.dtype
    return ((x - p[f"{name}.bn_mean"].to(dt).view(shape)) /
            torch.sqrt(p[f"{name}.bn_var"].to(dt).view(shape) + BN_EPS) *
            p[f"{name}.bn_gamma"].to(dt).view(shape) + p[f"{name}.bn_beta"].to(dt).view(shape))


def audionet_forward(x: torch.Tensor, p, return_all: bool = False):
    """x [B,N] -> logits [B,C]; audionet_csine.py:176-224 (eval-mode BN)."""
    if x.dim() == 3:
        x = x[:, 0]
    dt = x.dtype
    feat = audionet_logmel(x)                                      # [B,32,T]
    h = F.conv2d(feat.unsqueeze(1), p["conv1.weight"].to(dt), p["conv1.bias"].to(dt), padding=2)
    h = _bn(h, p, "conv1", 2).squeeze(1)
    for name, _, _, _, pad, pool in AN_CONVS:
        if name == "conv8" and h.shape[2] < 3:                     # audionet_csine.py:192-200
            h = h.repeat(1, 1, math.ceil(3 / h.shape[2]))
        h = F.conv1d(h, p[f"{name}.weight"].to(dt), p[f"{name}.bias"].to(dt), padding=pad)
        h = F.relu(_bn(h, p, name, 1))
        if pool:
            h = F.max_pool1d(h, 2, stride=2)
    emb = h.max(2)[0]
    logits = emb @ p["fc.weight"].to(dt).T + p["fc.bias"].to(dt)
    if return_all:
        return {"feat": feat, "emb": emb, "logits": logits}
    return logits


# ----------------------------------------------------------------------------------------------
# AudioNet in training mode (adver_train.py:183-223, natural_train.py:127-160): BatchNorm with batch
# statistics, running statistics updated with momentum 0.1, all parameter gradients, Adam
# ----------------------------------------------------------------------------------------------
AN_BN_NAMES = ["conv1"] + [n for n, *_ in AN_CONVS]


def audionet_train_forward(x: torch.Tensor, p, momentum: float = 0.1, eps: float = 1e-5, from_feat: bool = False):
    """x [B,N] waveform (or log-mel [B,T,32] when from_feat) -> (logits, new running stats dict).
    p holds tensors (leaf tensors with requires_grad for the parameters to differentiate)."""
    if from_feat:
        feat = x.transpose(1, 2)
    else:
        feat = audionet_logmel(x[:, 0] if x.dim() == 3 else x)
    new_stats = {}

    def bn(h, name, red):
        mean = h.mean(red)
        var_b = h.var(red, unbiased=False)
        n = h.numel() / mean.numel()
        new_stats[f"{name}.bn_mean"] = (1 - momentum) * p[f"{name}.bn_mean"] + momentum * mean.detach()
        new_stats[f"{name}.bn_var"] = (1 - momentum) * p[f"{name}.bn_var"] + momentum * var_b.detach() * n / (n - 1)
        shape = [1, -1] + [1] * (h.dim() - 2)
        return (h - mean.view(shape)) / torch.sqrt(var_b.view(shape) + eps) * p[f"{name}.bn_gamma"].view(shape) \
            + p[f"{name}.bn_beta"].view(shape)

    h = F.conv2d(feat.unsqueeze(1), p["conv1.weight"], p["conv1.bias"], padding=2)
    h = bn(h, "conv1", (0, 2, 3)).squeeze(1)
    for name, _, _, _, pad, pool in AN_CONVS:
        h = F.conv1d(h, p[f"{name}.weight"], p[f"{name}.bias"], padding=pad)
        h = F.relu(bn(h, name, (0, 2)))
        if pool:
            h = F.max_pool1d(h, 2, stride=2)
    emb = h.max(2)[0]
    return emb @ p["fc.weight"].T + p["fc.bias"], new_stats


AN_PARAM_KEYS = [f"{n}.{k}" for n in AN_BN_NAMES for k in ("weight", "bias", "bn_gamma", "bn_beta")] + ["fc.weight", "fc.bias"]


def audionet_train_step(x: torch.Tensor, y: torch.Tensor, p, lr: float = 1e-3, from_feat: bool = False):
    """One optimisation step as in adver_train.py:216-221 (CrossEntropyLoss mean, torch.optim.Adam defaults, first step).
    Returns dict: logits, loss, grads (per parameter), input grad, new running stats, updated parameters."""
    q = {k: v.detach().clone() for k, v in p.items()}
    for k in AN_PARAM_KEYS:
        q[k].requires_grad_(True)
    xr = x.detach().clone().requires_grad_(True)
    logits, stats = audionet_train_forward(xr, q, from_feat=from_feat)
    loss = F.cross_entropy(logits, y)
    params = [q[k] for k in AN_PARAM_KEYS]
    opt = torch.optim.Adam(params, lr=lr)
    opt.zero_grad()
    loss.backward()
    grads = {k: q[k].grad.detach().clone() for k in AN_PARAM_KEYS}
    opt.step()
    return {"logits": logits.detach(), "loss": loss.detach(), "grads": grads, "xgrad": xr.grad.detach(), "stats": stats,
            "params": {k: q[k].detach().clone() for k in AN_PARAM_KEYS}}


# ----------------------------------------------------------------------------------------------
# CW2 (attack/CW2.py:41-132) against a generic score function
# ----------------------------------------------------------------------------------------------
def cw2_attack(x: torch.Tensor, y: torch.Tensor, score_fn, targeted=False, confidence=0.0,
               initial_const=1e-3, binary_search_steps=9, max_iter=1000, stop_early=True,
               stop_early_iter=1000, lr=1e-2, task="CSI", threshold=None):
    """x [B,N]; score_fn(x[B,N]) -> scores [B,S] (differentiable).  Mirrors the host control
    flow of attack/CW2.py literally (python floats, sentinel -2)."""
    B = x.shape[0]
    const = torch.full((B,), initial_const, dtype=torch.float32)
    lower_b = torch.zeros(B)
    upper_b = torch.full((B,), 1e10)
    g_best_l2 = [np.inf] * B
    g_best_x = x.clone()
    g_best_score = [-2] * B
    thr = -math.inf if threshold is None else threshold
    for _ in range(binary_search_steps):
        modifier = torch.zeros_like(x, requires_grad=True)
        opt = torch.optim.Adam([modifier], lr=lr)
        best_l2, best_score = [np.inf] * B, [-2] * B
        cont, prev = True, np.inf
        for n_iter in range(max_iter + 1):
            if not cont:
                break
            inp = torch.tanh(modifier + torch.atanh(x * 0.999999))
            scores = score_fn(inp)
            dec = decide(scores.detach(), thr)
            loss1 = loss_margin(scores, y, targeted, confidence, task, threshold, clip_max=True)
            loss2 = torch.sum(torch.square(inp - x), dim=1)
            loss = const * loss1 + loss2
            if n_iter < max_iter:
                loss.backward(torch.ones_like(loss))
                opt.step()
                modifier.grad.zero_()
            lo, l1, l2 = loss.detach().tolist(), loss1.detach().tolist(), loss2.detach().tolist()
            if stop_early and n_iter % stop_early_iter == 0:
                if np.mean(lo) > 0.9999 * prev:
                    cont = False
                prev = np.mean(lo)
            for ii in range(B):
                if l1[ii] <= 0 and l2[ii] < best_l2[ii]:
                    best_l2[ii], best_score[ii] = l2[ii], int(dec[ii])
                if l1[ii] <= 0 and l2[ii] < g_best_l2[ii]:
                    g_best_l2[ii], g_best_score[ii] = l2[ii], int(dec[ii])
                    g_best_x[ii] = inp[ii].detach()
        for jj in range(B):
            if best_score[jj] != -2:
                upper_b[jj] = min(upper_b[jj], const[jj])
                if upper_b[jj] < 1e9:
                    const[jj] = (lower_b[jj] + upper_b[jj]) / 2
            else:
                lower_b[jj] = max(lower_b[jj], const[jj])
                if upper_b[jj] < 1e9:
                    const[jj] = (lower_b[jj] + upper_b[jj]) / 2
                else:
                    const[jj] *= 10
    success = [s != -2 for s in g_best_score]
    return g_best_x, success, {"const": const, "best_l2": g_best_l2}


# ----------------------------------------------------------------------------------------------
# iv_plda path (model/iv_plda.py:197-293, :380-396; model/_iv_plda/gmm.py:120-171;
# model/_iv_plda/ivector_extract.py:94-125): MFCC (24 ceps) -> deltas -> CMVN -> UBM posteriors
# -> Baum-Welch statistics -> i-vector -> LDA / length-norm / PLDA (shared with xv_plda)
# ----------------------------------------------------------------------------------------------
def make_iv_params(seed: int = 0, C: int = 64, Fd: int = 72, D: int = 40, L: int = 30, S: int = 1):
    """Synthetic full-covariance UBM + i-vector extractor + back-end, sized by the arguments
    (BASELINE config 5 uses C=2048, Fd=72, D=400, L=200)."""
    g = torch.Generator().manual_seed(seed + 31337)
    p = {}
    mu = 3.0 * torch.randn(C, Fd, generator=g)
    A = torch.randn(C, Fd, Fd, generator=g) / math.sqrt(Fd)
    invcov = 0.05 * (A @ A.transpose(1, 2)) + 0.06 * torch.eye(Fd)           # SPD, std ~ 3-4 per dim
    invcov = 0.5 * (invcov + invcov.transpose(1, 2))
    w = torch.softmax(torch.randn(C, generator=g), 0)
    p["gmm.invcovars"] = invcov
    p["gmm.means_invcovars"] = (invcov @ mu.unsqueeze(-1)).squeeze(-1)
    logdet = torch.linalg.slogdet(invcov.double())[1].float()
    quad = (mu.unsqueeze(1) @ invcov @ mu.unsqueeze(-1)).flatten()
    p["gmm.gconsts"] = torch.log(w) - 0.5 * (Fd * math.log(2 * math.pi) - logdet + quad)
    p["gmm.weights"] = w
    p["ive.T"] = 0.3 * torch.randn(C, Fd, D, generator=g)
    B_ = torch.randn(C, Fd, Fd, generator=g) / math.sqrt(Fd)
    sig = 0.05 * (B_ @ B_.transpose(1, 2)) + 0.06 * torch.eye(Fd)
    p["ive.sigma_inv"] = 0.5 * (sig + sig.transpose(1, 2))
    p["ive.offset"] = torch.tensor(5.0)
    p["plda.mean"] = _round6(0.1 * torch.randn(L, generator=g))
    p["plda.transform"] = _round6(torch.randn(L, L, generator=g) / math.sqrt(L))
    p["plda.psi"] = _round6(torch.randn(L, generator=g).abs() + 0.1)
    p["emb_mean"] = _round6(0.1 * torch.randn(D, generator=g))
    p["lda"] = _round6(torch.randn(L, D + 1, generator=g) / math.sqrt(D))
    p["enroll"] = torch.randn(S, L, generator=g)
    return p


def delta_scales(window: int = 3, order: int = 2) -> List[torch.Tensor]:
    """Kaldi delta filters, model/iv_plda.py:274-293 (get_scales)."""
    scales = [torch.tensor([1.0])]
    for _ in range(order):
        prev = scales[-1]
        po = (prev.numel() - 1) // 2
        cur = torch.zeros(prev.numel() + 2 * window)
        norm = 0.0
        for j in range(-window, window + 1):
            norm += j * j
            for k in range(-po, po + 1):
                cur[j + k + po + window] += j * prev[k + po]
        scales.append(cur / norm)
    return scales


def add_delta(feat: torch.Tensor, window: int = 3, order: int = 2) -> torch.Tensor:
    """[B,T,F] -> [B,T,F*(order+1)], edge frames replicated (model/iv_plda.py:248-271)."""
    B, T, Fd = feat.shape
    outs = []
    for s in delta_scales(window, order):
        off = (s.numel() - 1) // 2
        idx = (torch.arange(T).view(-1, 1) + torch.arange(-off, off + 1).view(1, -1)).clamp(0, T - 1)   # [T,2off+1]
        outs.append((feat[:, idx, :] * s.to(feat.dtype).view(1, 1, -1, 1)).sum(2))
    return torch.cat(outs, dim=2)


def gmm_loglike(x: torch.Tensor, p) -> torch.Tensor:
    """x [T,Fd] -> component log-likelihoods [T,C] (gmm.py:120-131)."""
    dt = x.dtype
    lin = x @ p["gmm.means_invcovars"].to(dt).T
    quad = torch.einsum("tf,cfg,tg->tc", x, p["gmm.invcovars"].to(dt), x)
    return lin - 0.5 * quad + p["gmm.gconsts"].to(dt)


def iv_stats(x: torch.Tensor, p) -> Tuple[torch.Tensor, torch.Tensor]:
    post = torch.softmax(gmm_loglike(x, p), -1)                 # gmm.py:133-136
    return post.sum(0), post.T @ x                              # gmm.py:166-171


def ivector(N: torch.Tensor, Fs: torch.Tensor, p) -> torch.Tensor:
    """Zeroth [C] / first [C,Fd] order statistics -> i-vector [D] (ivector_extract.py:94-114)."""
    dt = N.dtype
    T, Si = p["ive.T"].to(dt), p["ive.sigma_inv"].to(dt)
    TtS = T.transpose(1, 2) @ Si                                # [C,D,Fd]
    Lm = torch.eye(T.shape[2], dtype=dt) + (N.view(-1, 1, 1) * (TtS @ T)).sum(0)
    lin = (TtS @ Fs.unsqueeze(-1)).sum(dim=(0, 2))
    off = p["ive.offset"].to(dt)
    lin = lin + off * F.one_hot(torch.tensor(0), lin.numel()).to(dt)
    iv = torch.linalg.solve(Lm, lin)                            # reference: torch.inverse(L) @ linear
    return iv - off * F.one_hot(torch.tensor(0), iv.numel()).to(dt)


def iv_forward(x: torch.Tensor, p, dither: Optional[torch.Tensor] = None, return_all: bool = False):
    """x [B,N] -> scores [B,S] through the i-vector system (iv_plda.forward with flag 0)."""
    if x.dim() == 3:
        x = x[:, 0]
    raw = mfcc(x, dither, num_ceps=24)                           # model/iv_plda.py:203-237
    delta = add_delta(raw)
    feat = cmvn(delta)
    ivs = []
    for f in feat:                                               # per utterance (model/iv_plda.py:380-396)
        N, Fs = iv_stats(f, p)
        ivs.append(ivector(N, Fs, p))
    iv = torch.stack(ivs)
    emb = process_emb(iv, p)
    scores = plda_scores(emb, p)
    if return_all:
        return {"raw": raw, "delta": delta, "feat": feat, "ivector": iv, "emb": emb, "scores": scores}
    return scores
