"""Drop-in for the reference's ``model.xv_plda.xv_plda`` (model/xv_plda.py:15-174 with the methods
it inherits from model/iv_plda.py:155-194, :296-443), backed by libsgb200.

Same constructor arguments, attributes (``threshold``, ``allowed_flags``, ``range_type``,
``spk_ids``, ``num_spks``, ``enroll_embs``, ``device``) and method signatures, so
``defended_model``, the attack classes and evaluation scripts run unchanged.  All arithmetic
happens in the CUDA library: there is no PyTorch/CPU implementation behind these methods.

Extra keyword arguments (engine options):
  precision  'fp32' (FFMA, parity mode) | 'tf32' | 'bf16' (tcgen05 tensor cores)
  dither     'philox' (in-kernel counter-based N(0,1), default) | 'torch' (torch.randn drawn per
             utterance in batch order on the model's device, the reference's own call sequence,
             kaldi.py:180) | 'off' | a callable (B, m) -> tensor [B,m,400]
  seed       base seed of the philox stream
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional, Union

import numpy as np
import torch
import torch.nn as nn

from .. import _lib
from ..engine import Engine
from ..functional import CmvnFn, EmbedFn, MfccFn, ScoreFn
from .utils import (check_input_range, parse_enroll_model_file, parse_mean_file, parse_plda_file,
                    parse_transform_mat_file)


def _state_dict_of(extractor_file) -> Dict[str, torch.Tensor]:
    if isinstance(extractor_file, str):
        sd = torch.load(extractor_file, map_location="cpu")
    elif isinstance(extractor_file, dict):
        sd = extractor_file
    elif hasattr(extractor_file, "state_dict"):
        sd = extractor_file.state_dict()
    else:
        raise NotImplementedError("extractor_file must be a checkpoint path, a state dict or an x-vector TDNN module")
    return {k: v.detach().cpu() for k, v in sd.items() if torch.is_tensor(v)}


class xv_plda(nn.Module):

    def __init__(self, extractor_file, plda_file, mean_file, transform_mat_file, model_file=None, threshold=None,
                 device="cuda", precision: str = "fp32", dither: Union[str, Callable] = "philox", seed: int = 0):
        super().__init__()
        dev = torch.device(device)
        if dev.type != "cuda":
            raise _lib.SgError("speakerguard_b200.xv_plda needs a CUDA device (no CPU fallback); got '%s'" % device)
        self.engine = Engine(dev, precision=precision)
        self.device = self.engine.device
        self.extractor_file, self.plda_file = extractor_file, plda_file

        sd = _state_dict_of(extractor_file)
        p: Dict[str, torch.Tensor] = {}
        for i in range(1, 6):
            p[f"tdnn{i}.weight"], p[f"tdnn{i}.bias"] = sd[f"tdnn{i}.weight"], sd[f"tdnn{i}.bias"]
            p[f"bn{i}.mean"], p[f"bn{i}.var"] = sd[f"bn_tdnn{i}.running_mean"], sd[f"bn_tdnn{i}.running_var"]
        p["fc1.weight"], p["fc1.bias"] = sd["fc1.weight"], sd["fc1.bias"]
        mean, transform, psi = parse_plda_file(plda_file)
        p["plda.mean"], p["plda.transform"], p["plda.psi"] = map(torch.from_numpy, (mean, transform, psi))
        self.emb_mean = parse_mean_file(mean_file, self.device)
        self.transform_mat = parse_transform_mat_file(transform_mat_file, self.device)
        p["emb_mean"], p["lda"] = self.emb_mean, self.transform_mat
        if model_file is not None:
            self.num_spks, self.spk_ids, self.z_norm_means, self.z_norm_stds, self.enroll_embs = \
                parse_enroll_model_file(model_file, self.device)
            p["enroll"] = self.enroll_embs
            # the fused paths score against the copy the engine took here: remember which tensor (and which version of it)
            # that was, so a later assignment to / in-place edit of model.enroll_embs falls back to the stage-wise path
            self._engine_enroll = (self.enroll_embs, self.enroll_embs._version)
        else:
            p["enroll"] = torch.zeros(1, mean.shape[0])     # placeholder; forward() then requires enroll_embs
        self.engine.load_xv(p)

        self.threshold = threshold if threshold else -np.inf   # CSI: -inf (model/xv_plda.py:43)
        self.allowed_flags = sorted([0, 1, 2])                 # 0: wav; 1: raw feat; 2: cmvn feat
        self.range_type = "origin"
        self.dither = dither
        self.seed = int(seed)
        self._pass = 0

    # nn.Module.eval()/train() keep working; the engine is always in inference mode (BN running stats)

    # ---- features ------------------------------------------------------------------------------
    def _draw_dither(self, B: int, m: int):
        if callable(self.dither):
            return _lib.DITHER_TENSOR, self.dither(B, m).to(self.device)
        if self.dither == "torch":
            d = torch.stack([torch.randn((m, 400), device=self.device, dtype=torch.float32) for _ in range(B)])
            return _lib.DITHER_TENSOR, d
        if self.dither == "off":
            return _lib.DITHER_OFF, None
        if self.dither == "philox":
            return _lib.DITHER_PHILOX, None
        raise ValueError(f"unknown dither mode {self.dither!r}")

    def raw(self, x):
        """x: (B, 1, T) waveform in int16 range (model/xv_plda.py:107-156) -> (B, frames, 30)."""
        x2 = x[:, 0, :] / float(2 ** 15)       # the kernel applies the 2^15 scale itself (exact)
        B, N = x2.shape
        mode, d = self._draw_dither(B, self.engine.num_frames(N))
        out = MfccFn.apply(x2, self.engine, mode, d, self.seed, self._pass)
        self._pass += 1
        return out

    def cmvn(self, batch_delta_feat):
        return CmvnFn.apply(batch_delta_feat, self.engine)

    def compute_feat(self, x, flag=1):
        assert flag in [f for f in self.allowed_flags if f != 0]
        x = check_input_range(x, range_type=self.range_type)
        feats = self.raw(x)
        if flag == 1:
            return feats
        return self.comput_feat_from_feat(feats, ori_flag=1, des_flag=2)

    def comput_feat_from_feat(self, feats, ori_flag=1, des_flag=2):
        assert ori_flag in [f for f in self.allowed_flags if f != 0]
        assert des_flag in [f for f in self.allowed_flags if f != 0]
        assert des_flag > ori_flag
        return self.cmvn(feats)

    # ---- embedding / scoring -------------------------------------------------------------------
    def extract_emb(self, x):
        """x: (B, T, F) CMVN features -> (B, L) embeddings in PLDA space (model/xv_plda.py:159-174)."""
        return EmbedFn.apply(x, self.engine)

    def embedding(self, x, flag=0):
        assert flag in self.allowed_flags
        if flag == 0:
            feats = self.compute_feat(x, flag=self.allowed_flags[-1])
        elif flag == 1:
            feats = self.comput_feat_from_feat(x, ori_flag=1, des_flag=self.allowed_flags[-1])
        else:
            feats = x
        return self.extract_emb(feats)

    def scoring_trials(self, enroll_embs, embs):
        return ScoreFn.apply(embs, enroll_embs, self.engine)

    def _fused_forward(self, x):
        """Forward-only wav -> (scores, decisions, emb) through sg_xv_forward: one call, no autograd
        bookkeeping (evaluation scripts and black-box attacks call make_decision thousands of times:
        reference test_attack.py:123-128, attack/FAKEBOB.py:170-174)."""
        x = check_input_range(x, range_type=self.range_type)
        x2 = (x[:, 0, :] / float(2 ** 15)).contiguous()
        B, N = x2.shape
        mode, d = self._draw_dither(B, self.engine.num_frames(N))
        # the engine keeps (and grows) one scratch block per stream; asking each time never pins a stale block
        out = self.engine.xv_forward(x2, mode, d, self.seed, self._pass, self.decision_threshold,
                                     ws=self.engine.pgd_ws(B, N))
        self._pass += 1
        return out

    def engine_enroll_current(self) -> bool:
        """True while ``self.enroll_embs`` is still the tensor (same object, unmodified) the engine copied at construction."""
        held = getattr(self, "_engine_enroll", None)
        cur = getattr(self, "enroll_embs", None)
        return held is not None and cur is held[0] and cur._version == held[1]

    def _can_fuse(self, x, flag, enroll_embs):
        return (flag == 0 and enroll_embs is None and self.engine_enroll_current()
                and not (torch.is_grad_enabled() and x.requires_grad))

    def forward(self, x, flag=0, return_emb=False, enroll_embs=None):
        if self._can_fuse(x, flag, enroll_embs):
            scores, _, emb = self._fused_forward(x)
            return (scores, emb) if return_emb else scores
        embedding = self.embedding(x, flag=flag)
        if not hasattr(self, "enroll_embs"):
            assert enroll_embs is not None
        enroll_embs = enroll_embs if enroll_embs is not None else self.enroll_embs
        scores = self.scoring_trials(enroll_embs=enroll_embs, embs=embedding)
        return (scores, embedding) if return_emb else scores

    def score(self, x, flag=0, enroll_embs=None):
        return self.forward(x, flag=flag, enroll_embs=enroll_embs)

    def make_decision(self, x, flag=0, enroll_embs=None):
        if self._can_fuse(x, flag, enroll_embs):
            scores, decisions, _ = self._fused_forward(x)
            return decisions, scores
        scores = self.score(x, flag=flag, enroll_embs=enroll_embs)
        decisions = torch.argmax(scores, dim=1)
        max_scores = torch.max(scores, dim=1)[0]
        decisions = torch.where(max_scores > self.threshold, decisions, torch.full_like(decisions, -1))
        return decisions, scores

    # ---- hooks used by the fused attack path ---------------------------------------------------
    def fused_dither(self, n_pass: int, B: int, N: int):
        """(mode, tensor [n_pass,B,m,400] or None, seed) for sg_pgd_run, honouring ``self.dither``."""
        m = self.engine.num_frames(N)
        if callable(self.dither) or self.dither == "torch":
            parts = [self._draw_dither(B, m)[1] for _ in range(n_pass)]
            return _lib.DITHER_TENSOR, torch.stack(parts), self.seed
        mode, _ = self._draw_dither(B, m)
        seed = self.seed + 0x9E3779B97F4A7C15 * self._pass & 0xFFFFFFFFFFFFFFFF
        self._pass += n_pass
        return mode, None, seed

    @property
    def decision_threshold(self) -> float:
        t = float(self.threshold)
        return t if math.isfinite(t) else -math.inf
