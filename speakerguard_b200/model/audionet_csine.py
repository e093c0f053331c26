"""Drop-in for the reference's ``model.audionet_csine.audionet_csine`` (model/audionet_csine.py:20-257;
front end model/_audionet/Preprocessor.py:48-112), inference / attack side only, backed by
libsgb200.  Same constructor arguments, attributes and method signatures; BatchNorm runs with its
running statistics (the reference calls ``.eval()`` when a checkpoint is given).  Training
(natural_train.py / adver_train.py) stays with the reference: SURVEY.md 8(f) rank 4.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
import torch.nn as nn

from .. import _lib
from ..engine import Engine
from ..functional import AnCnnFn, AnLogMelFn
from .utils import check_input_range

_CONVS = ["conv2", "conv3", "conv4", "conv5", "conv6", "conv7", "conv8"]


def params_from_state_dict(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Reference state-dict keys ('conv2.0.weight', 'conv2.1.running_mean', ...) -> engine names."""
    p = {}
    for n in ["conv1"] + _CONVS:
        p[f"{n}.weight"], p[f"{n}.bias"] = sd[f"{n}.0.weight"], sd[f"{n}.0.bias"]
        p[f"{n}.bn_mean"], p[f"{n}.bn_var"] = sd[f"{n}.1.running_mean"], sd[f"{n}.1.running_var"]
        p[f"{n}.bn_gamma"], p[f"{n}.bn_beta"] = sd[f"{n}.1.weight"], sd[f"{n}.1.bias"]
    p["fc.weight"], p["fc.bias"] = sd["fc.weight"], sd["fc.bias"]
    return p


class audionet_csine(nn.Module):

    def __init__(self, extractor_file=None, num_class=None, label_encoder=None, device="cuda", params=None):
        """extractor_file: checkpoint path / state dict of the reference model; ``params`` (engine
        naming, see Engine.load_audionet) may be given instead for synthetic models."""
        super().__init__()
        dev = torch.device(device)
        if dev.type != "cuda":
            raise _lib.SgError("speakerguard_b200.audionet_csine needs a CUDA device (no CPU fallback); got '%s'" % device)
        self.engine = Engine(dev)
        self.device = self.engine.device
        if params is None:
            if extractor_file is None:
                raise NotImplementedError("training AudioNet from scratch is out of scope here: give extractor_file or params")
            sd = torch.load(extractor_file, map_location="cpu") if isinstance(extractor_file, str) else extractor_file
            params = params_from_state_dict(sd)
        n_ckpt = int(params["fc.bias"].shape[0])
        if label_encoder is not None:
            id_label = np.loadtxt(label_encoder, dtype=str, converters={0: lambda s: s[1:-1]})
            self.id2label = {row[0]: int(row[1]) for row in id_label}
            self.label2id = {int(row[1]): row[0] for row in id_label}
            self.spk_ids = [self.label2id[i] for i in range(len(self.label2id))]
            assert len(self.spk_ids) == n_ckpt
        if num_class is not None:
            assert num_class == n_ckpt
        self.num_spks = n_ckpt
        if not hasattr(self, "spk_ids"):
            self.spk_ids = [str(i) for i in range(self.num_spks)]
        self.engine.load_audionet(params)
        self.threshold = -np.inf            # CSI-NE: never rejects
        self.allowed_flags = sorted([0, 1])  # 0: wav; 1: raw feat
        self.range_type = "scale"

    def compute_feat(self, x, flag=1):
        assert flag in [f for f in self.allowed_flags if f != 0]
        x = check_input_range(x, range_type=self.range_type)
        return self.raw(x)

    def raw(self, x):
        """x: (B, 1, T) in [-1,1] -> log-mel (B, frames, 32)."""
        return AnLogMelFn.apply(x[:, 0, :], self.engine)

    def extract_emb(self, x):
        raise NotImplementedError("the engine fuses extract_emb and the final fc; use forward()/score()")

    def embedding(self, x, flag=0):
        raise NotImplementedError("the engine fuses extract_emb and the final fc; use forward()/score()")

    def forward(self, x, flag=0, return_emb=False, enroll_embs=None):
        assert flag in self.allowed_flags
        if return_emb:
            raise NotImplementedError("return_emb is not supported by the fused AudioNet CNN")
        if flag == 0:
            n = x.shape[2]
            feats = self.compute_feat(x, flag=1)
        else:
            feats = x
            n = (x.shape[1] - 1) * 160 + 1    # any length with the same frame count
        return AnCnnFn.apply(feats, self.engine, n)

    def score(self, x, flag=0, enroll_embs=None):
        return self.forward(x, flag=flag)

    def make_decision(self, x, flag=0, enroll_embs=None):
        scores = self.score(x, flag=flag)
        return torch.argmax(scores, dim=1), scores
