"""Drop-in for the reference's ``model.audionet_csine.audionet_csine`` (model/audionet_csine.py:20-257;
front end model/_audionet/Preprocessor.py:48-112), backed by libsgb200.  Same constructor arguments,
attributes, method signatures and ``state_dict`` keys ('conv2.0.weight', 'conv2.1.running_mean', ...), so
reference checkpoints load and ``torch.optim.Adam(model.parameters())`` trains it.

* ``eval()`` (what the reference enters when a checkpoint is given): BatchNorm with running statistics, folded
  into the convolutions inside the engine - the attack / evaluation path.
* ``train()`` (natural_train.py / adver_train.py; also the mode the attack inside adver_train.py runs in):
  batch-statistics BatchNorm with the momentum update of the running statistics, gradients for every parameter.

The submodules below only *hold* the parameters and buffers; the arithmetic runs in the CUDA library.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
import torch.nn as nn

from .. import _lib
from ..engine import Engine
from ..functional import AnCnnFn, AnCnnTrainFn, AnEmbFn, AnFcFn, AnLogMelFn
from .utils import check_input_range

_CONVS = ["conv2", "conv3", "conv4", "conv5", "conv6", "conv7", "conv8"]
_SPEC = [(32, 64, 1, True), (64, 128, 1, False), (128, 128, 1, False), (128, 128, 1, True), (128, 128, 1, False),
         (128, 64, 1, True), (64, 32, 0, False)]            # (C_in, C_out, padding, max-pool)  audionet_csine.py:66-118


def params_from_state_dict(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Reference state-dict keys ('conv2.0.weight', 'conv2.1.running_mean', ...) -> engine names."""
    p = {}
    for n in ["conv1"] + _CONVS:
        p[f"{n}.weight"], p[f"{n}.bias"] = sd[f"{n}.0.weight"], sd[f"{n}.0.bias"]
        p[f"{n}.bn_mean"], p[f"{n}.bn_var"] = sd[f"{n}.1.running_mean"], sd[f"{n}.1.running_var"]
        p[f"{n}.bn_gamma"], p[f"{n}.bn_beta"] = sd[f"{n}.1.weight"], sd[f"{n}.1.bias"]
    p["fc.weight"], p["fc.bias"] = sd["fc.weight"], sd["fc.bias"]
    return p


def state_dict_from_params(p: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    sd = {}
    for n in ["conv1"] + _CONVS:
        sd[f"{n}.0.weight"], sd[f"{n}.0.bias"] = p[f"{n}.weight"], p[f"{n}.bias"]
        sd[f"{n}.1.running_mean"], sd[f"{n}.1.running_var"] = p[f"{n}.bn_mean"], p[f"{n}.bn_var"]
        sd[f"{n}.1.weight"], sd[f"{n}.1.bias"] = p[f"{n}.bn_gamma"], p[f"{n}.bn_beta"]
    sd["fc.weight"], sd["fc.bias"] = p["fc.weight"], p["fc.bias"]
    return sd


class audionet_csine(nn.Module):

    def __init__(self, extractor_file=None, num_class=None, label_encoder=None, device="cuda", params=None,
                 precision="fp32"):
        """extractor_file: checkpoint path / state dict of the reference model (-> eval mode); without it the model is
        freshly initialised for training and ``num_class`` (or ``label_encoder``) is required, as in the reference.
        ``params`` (engine naming, see Engine.load_audionet) may be given instead of a checkpoint for synthetic models.
        ``precision``: "fp32" (FFMA convolutions, the parity mode) or "tf32" (the eval-mode convolutions and their adjoints on
        tcgen05 tiles with TF32 operands and fp32 storage - the precision class of the reference's own GPU default)."""
        super().__init__()
        assert precision in ("fp32", "tf32")
        self.precision = precision
        dev = torch.device(device)
        if dev.type != "cuda":
            raise _lib.SgError("speakerguard_b200.audionet_csine needs a CUDA device (no CPU fallback); got '%s'" % device)
        self.device = dev if dev.index is not None else torch.device("cuda", torch.cuda.current_device())
        sd = None
        if params is not None:
            sd = state_dict_from_params(params)
        elif extractor_file is not None:
            sd = torch.load(extractor_file, map_location="cpu") if isinstance(extractor_file, str) else extractor_file
        n_ckpt = int(sd["fc.bias"].shape[0]) if sd is not None else None
        n_enc = None
        if label_encoder is not None:
            id_label = np.loadtxt(label_encoder, dtype=str, converters={0: lambda s: s[1:-1]})
            self.id2label = {row[0]: int(row[1]) for row in id_label}
            self.label2id = {int(row[1]): row[0] for row in id_label}
            self.spk_ids = [self.label2id[i] for i in range(len(self.label2id))]
            n_enc = len(self.spk_ids)
        known = [k for k in (n_ckpt, n_enc) if k is not None]
        if len(known) == 2:
            assert n_ckpt == n_enc
        if known:
            if num_class is not None and n_ckpt is not None:
                assert num_class == n_ckpt
            num_class = known[0]
        assert num_class is not None, "num_class is required when neither a checkpoint nor a label encoder is given"
        self.num_spks = int(num_class)
        if not hasattr(self, "spk_ids"):
            self.spk_ids = [str(i) for i in range(self.num_spks)]

        # parameter / buffer containers in the reference's construction order (same default initialisation stream)
        self.conv1 = nn.Sequential(nn.Conv2d(1, 1, kernel_size=[5, 5], stride=1, padding=[2, 2]), nn.BatchNorm2d(1))
        for name, (ci, co, pad, pool) in zip(_CONVS, _SPEC):
            layers = [nn.Conv1d(ci, co, kernel_size=3, stride=1, padding=pad), nn.BatchNorm1d(co), nn.ReLU()]
            if pool:
                layers.append(nn.MaxPool1d(2, stride=2))
            setattr(self, name, nn.Sequential(*layers))
        self.fc = nn.Linear(32, self.num_spks)
        if sd is not None:
            self.load_state_dict({k: torch.as_tensor(v) for k, v in sd.items()}, strict=False)
            self.eval()
        else:
            self.train()
        self.to(self.device)

        self.engine = None
        self._engine_version = None
        self._sync_engine()
        self.threshold = -np.inf            # CSI-NE: never rejects
        self.allowed_flags = sorted([0, 1])  # 0: wav; 1: raw feat
        self.range_type = "scale"

    # ---- engine state ----------------------------------------------------------------------------
    def _bn_modules(self):
        return [self.conv1[1]] + [getattr(self, n)[1] for n in _CONVS]

    def _version(self):
        return tuple(t._version for t in list(self.parameters()) + list(self.buffers()))

    def _sync_engine(self):
        """(Re)build the inference-side engine state (BatchNorm folded into the convolutions) when a parameter or a running
        statistic has changed since it was last built: after optimiser steps, load_state_dict or train-mode forwards."""
        v = self._version()
        if self.engine is not None and v == self._engine_version:
            return
        eng = Engine(self.device, precision=self.precision)
        eng.load_audionet(params_from_state_dict({k: t.detach() for k, t in self.state_dict().items()}))
        self.engine, self._engine_version = eng, v

    def _train_params(self):
        convs = [getattr(self, n)[0] for n in _CONVS]
        bns = self._bn_modules()
        return ([self.conv1[0].weight, self.conv1[0].bias] + [c.weight for c in convs] + [c.bias for c in convs]
                + [b.weight for b in bns] + [b.bias for b in bns] + [self.fc.weight, self.fc.bias])

    # ---- reference API ---------------------------------------------------------------------------
    def compute_feat(self, x, flag=1):
        assert flag in [f for f in self.allowed_flags if f != 0]
        x = check_input_range(x, range_type=self.range_type)
        return self.raw(x)

    def raw(self, x):
        """x: (B, 1, T) in [-1,1] -> log-mel (B, frames, 32)."""
        return AnLogMelFn.apply(x[:, 0, :], self.engine)

    def _eval_only(self, what):
        if self.training:
            raise NotImplementedError(f"{what} is built for eval() mode (running-statistics BatchNorm); the train-mode kernels "
                                      "fuse the CNN with the final fc")
        self._sync_engine()

    def extract_emb(self, x):
        """x: (B, T, F) log-mel -> (B, 32) embedding (audionet_csine.py:176-207)."""
        self._eval_only("extract_emb")
        return AnEmbFn.apply(x, self.engine, (x.shape[1] - 1) * 160 + 1)

    def embedding(self, x, flag=0):
        """x: wav (flag 0) or log-mel (flag 1) -> (B, 32) (audionet_csine.py:159-173)."""
        assert flag in self.allowed_flags
        feats = self.compute_feat(x, flag=1) if flag == 0 else x
        return self.extract_emb(feats)

    def predict_from_embeddings(self, x):
        self._eval_only("predict_from_embeddings")
        return AnFcFn.apply(x, self.engine)

    def forward(self, x, flag=0, return_emb=False, enroll_embs=None):
        assert flag in self.allowed_flags
        if return_emb:
            embedding = self.embedding(x, flag=flag)
            return self.predict_from_embeddings(embedding), embedding
        if flag == 0:
            n = x.shape[2]
            feats = self.compute_feat(x, flag=1)
        else:
            feats = x
            n = (x.shape[1] - 1) * 160 + 1    # any length with the same frame count
        if self.training:
            bns = self._bn_modules()
            momentum = bns[0].momentum if bns[0].momentum is not None else 0.1
            running = ([b.running_mean for b in bns], [b.running_var for b in bns])
            out = AnCnnTrainFn.apply(feats, self.engine, n, running, momentum, bns[0].eps, *self._train_params())
            for b in bns:                    # the kernels wrote through raw pointers: tell autograd / the version check
                b.running_mean.add_(0)
                b.running_var.add_(0)
                b.num_batches_tracked += 1
            return out
        self._sync_engine()
        return AnCnnFn.apply(feats, self.engine, n)

    def score(self, x, flag=0, enroll_embs=None):
        return self.forward(x, flag=flag)

    def make_decision(self, x, flag=0, enroll_embs=None):
        scores = self.score(x, flag=flag)
        return torch.argmax(scores, dim=1), scores
