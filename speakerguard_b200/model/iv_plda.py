"""Drop-in for the reference's ``model.iv_plda.iv_plda`` (model/iv_plda.py:17-443: GMM-UBM i-vector
extractor + PLDA back-end), backed by libsgb200.

Same constructor arguments, attributes (``threshold``, ``allowed_flags`` 0 wav / 1 raw / 2 delta /
3 cmvn, ``range_type``, ``spk_ids``, ``num_spks``, ``enroll_embs``, ``device``) and method signatures
as the reference class.  ``gmm_frame_bs`` is accepted and ignored: the engine evaluates the UBM for all
frames of the batch in one contraction.  There is no PyTorch/CPU implementation behind these methods.

Extra keyword arguments (engine options):
  precision 'fp32' (every contraction in FFMA, the parity mode) | 'tf32x3' (the two frames-x-components contractions, UBM
           log-likelihoods and their adjoint, on the tcgen05 tensor cores as split-TF32: operands split into tf32-exact hi + lo
           parts, three products, fp32 accumulation: fp32-level accuracy at tensor-core speed)
  params   dict of dense tensors instead of the Kaldi text files ('gmm.gconsts', 'gmm.means_invcovars',
           'gmm.invcovars', 'ive.T', 'ive.sigma_inv', 'ive.offset', 'plda.mean/.transform/.psi',
           'emb_mean', 'lda', optionally 'enroll'); the five file arguments may then be None
  dither   'philox' | 'torch' | 'off' | callable, as in xv_plda
  seed     base seed of the philox stream
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional, Union

import numpy as np
import torch
import torch.nn as nn

from .. import _lib
from ..engine import Engine
from ..functional import CmvnColsFn, DeltaFn, IvEmbedFn, Mfcc24Fn, ScoreFn
from .utils import (check_input_range, parse_enroll_model_file, parse_fgmm_file, parse_ivector_extractor_file,
                    parse_mean_file, parse_plda_file, parse_transform_mat_file)

NUM_CEPS = 24      # model/iv_plda.py:232


class iv_plda(nn.Module):

    def __init__(self, fgmm_file, extractor_file, plda_file, mean_file, transform_mat_file, model_file=None, threshold=None,
                 device="cuda", gmm_frame_bs=200, params: Optional[Dict[str, torch.Tensor]] = None,
                 dither: Union[str, Callable] = "philox", seed: int = 0, precision: str = "fp32"):
        super().__init__()
        dev = torch.device(device)
        if dev.type != "cuda":
            raise _lib.SgError("speakerguard_b200.iv_plda needs a CUDA device (no CPU fallback); got '%s'" % device)
        if precision not in ("fp32", "tf32x3"):
            raise ValueError(f"iv_plda precision must be 'fp32' or 'tf32x3', got {precision!r}")
        self.engine = Engine(dev, precision="fp32" if precision == "fp32" else "tf32")
        self.device = self.engine.device
        self.fgmm_file, self.extractor_file, self.plda_file = fgmm_file, extractor_file, plda_file
        self.gmm_frame_bs = gmm_frame_bs

        p: Dict[str, torch.Tensor] = {}
        if params is not None:
            p.update({k: torch.as_tensor(v) for k, v in params.items()})
        else:
            gconsts, _, mic, inv = parse_fgmm_file(fgmm_file)
            p["gmm.gconsts"], p["gmm.means_invcovars"], p["gmm.invcovars"] = map(torch.from_numpy, (gconsts, mic, inv))
            T, sig, off = parse_ivector_extractor_file(extractor_file)
            p["ive.T"], p["ive.sigma_inv"], p["ive.offset"] = torch.from_numpy(T), torch.from_numpy(sig), torch.tensor(off)
            mean, transform, psi = parse_plda_file(plda_file)
            p["plda.mean"], p["plda.transform"], p["plda.psi"] = map(torch.from_numpy, (mean, transform, psi))
            p["emb_mean"] = parse_mean_file(mean_file, "cpu")
            p["lda"] = parse_transform_mat_file(transform_mat_file, "cpu")
        self.emb_mean = p["emb_mean"].to(self.device, torch.float32)
        self.transform_mat = p["lda"].to(self.device, torch.float32)
        if model_file is not None:
            self.num_spks, self.spk_ids, self.z_norm_means, self.z_norm_stds, self.enroll_embs = \
                parse_enroll_model_file(model_file, self.device)
            p["enroll"] = self.enroll_embs
        elif "enroll" in p:
            self.enroll_embs = p["enroll"].to(self.device, torch.float32)
            self.num_spks = int(self.enroll_embs.shape[0])
            self.spk_ids = [f"spk{i}" for i in range(self.num_spks)]
        else:
            p["enroll"] = torch.zeros(1, p["plda.mean"].shape[0])   # placeholder; forward() then requires enroll_embs
        self.engine.load_iv(p)

        self.threshold = threshold if threshold else -np.inf    # SV / OSI need a threshold; CSI: -inf (model/iv_plda.py:70)
        self.allowed_flags = sorted([0, 1, 2, 3])               # 0: wav; 1: raw feat; 2: delta feat; 3: cmvn feat
        self.range_type = "origin"
        self.dither = dither
        self.seed = int(seed)
        self._pass = 0

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
        """x: (B, 1, T) waveform in int16 range -> (B, frames, 24)  (model/iv_plda.py:196-246)."""
        x2 = x[:, 0, :] / float(2 ** 15)
        B, N = x2.shape
        mode, d = self._draw_dither(B, self.engine.num_frames(N))
        out = Mfcc24Fn.apply(x2, self.engine, mode, d, self.seed, self._pass, NUM_CEPS)
        self._pass += 1
        return out

    def add_delta(self, batch_raw_feat, window=3, order=2, mode="replicate"):
        """(B, T, F) -> (B, T, 3F)  (model/iv_plda.py:248-271); only the reference's default filter is built."""
        if (window, order, mode) != (3, 2, "replicate"):
            raise NotImplementedError("the engine implements add_delta for window=3, order=2, mode='replicate'")
        return DeltaFn.apply(batch_raw_feat, self.engine)

    def cmvn(self, batch_delta_feat):
        return CmvnColsFn.apply(batch_delta_feat, self.engine)

    def compute_feat(self, x, flag=1):
        assert flag in [f for f in self.allowed_flags if f != 0]
        x = check_input_range(x, range_type=self.range_type)
        feats = self.raw(x)
        if flag == 1:
            return feats
        return self.comput_feat_from_feat(feats, ori_flag=1, des_flag=flag)

    def comput_feat_from_feat(self, feats, ori_flag=1, des_flag=2):
        assert ori_flag in [f for f in self.allowed_flags if f != 0]
        assert des_flag in [f for f in self.allowed_flags if f != 0]
        assert des_flag > ori_flag
        if ori_flag == 1:
            feats = self.add_delta(feats)
            if des_flag == 2:
                return feats
        return self.cmvn(feats)

    # ---- embedding / scoring -------------------------------------------------------------------
    def extract_emb(self, x):
        """x: (B, T, 72) CMVN features -> (B, L) embeddings in PLDA space (model/iv_plda.py:380-396, :411-443)."""
        return IvEmbedFn.apply(x, self.engine)

    def embedding(self, x, flag=0):
        assert flag in self.allowed_flags
        if flag == 0:
            feats = self.compute_feat(x, flag=self.allowed_flags[-1])
        elif flag in (1, 2):
            feats = self.comput_feat_from_feat(x, ori_flag=flag, des_flag=self.allowed_flags[-1])
        else:
            feats = x
        return self.extract_emb(feats)

    def scoring_trials(self, enroll_embs, embs):
        return ScoreFn.apply(embs, enroll_embs, self.engine)

    def forward(self, x, flag=0, return_emb=False, enroll_embs=None):
        embedding = self.embedding(x, flag=flag)
        if not hasattr(self, "enroll_embs"):
            assert enroll_embs is not None
        enroll_embs = enroll_embs if enroll_embs is not None else self.enroll_embs
        scores = self.scoring_trials(enroll_embs=enroll_embs, embs=embedding)
        return (scores, embedding) if return_emb else scores

    def score(self, x, flag=0, enroll_embs=None):
        return self.forward(x, flag=flag, enroll_embs=enroll_embs)

    def make_decision(self, x, flag=0, enroll_embs=None):
        scores = self.score(x, flag=flag, enroll_embs=enroll_embs)
        decisions = torch.argmax(scores, dim=1)
        max_scores = torch.max(scores, dim=1)[0]
        decisions = torch.where(max_scores > self.threshold, decisions, torch.full_like(decisions, -1))
        return decisions, scores

    @property
    def decision_threshold(self) -> float:
        t = float(self.threshold)
        return t if math.isfinite(t) else -math.inf
