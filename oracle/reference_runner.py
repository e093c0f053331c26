"""TEST / BENCH INFRASTRUCTURE - never imported by the product (speakerguard_b200/).

Runs the UNMODIFIED reference (SpeakerGuard) classes on the host CPU: ``bench.py --impl reference`` and
``bench.py``'s ``cpu_baseline`` leg time them, ``tests/golden/make_golden.py`` uses the same recipe to
dump fixtures.  The reference is pure Python, so "building" it is a file copy: ``tools/install_reference.py``
(called by ``__graft_entry__.build()`` when /root/reference exists) copies its *.py files into the
git-ignored ``baseline/_ref/`` so they travel to the GPU box with the snapshot; nothing under
``baseline/_ref`` is edited and nothing from it is committed.

Import shims (applied before any reference module is imported):
  * ``kaldi_io``  - imported at model/_xv_plda/plda.py:13, used only by ReadIvectors: an empty module;
  * ``np.infty``  - model/xv_plda.py:43 (removed in NumPy 2): alias of ``np.inf``.
"""
from __future__ import annotations

import os
import sys
import tempfile
import time
import types
from typing import Dict, Optional

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIRS = [os.path.join(ROOT, "baseline", "_ref"), "/root/reference"]


def reference_dir() -> Optional[str]:
    for d in REF_DIRS:
        if os.path.exists(os.path.join(d, "attack", "PGD.py")) and os.path.exists(os.path.join(d, "model", "xv_plda.py")):
            return d
    return None


def import_reference() -> str:
    """Put the reference on sys.path (once) with the shims in place; returns the directory used."""
    d = reference_dir()
    if d is None:
        raise FileNotFoundError("reference not found: run tools/install_reference.py in the build container "
                                "(copies /root/reference/*.py to baseline/_ref)")
    sys.modules.setdefault("kaldi_io", types.ModuleType("kaldi_io"))
    if not hasattr(np, "infty"):
        np.infty = np.inf
    if d not in sys.path:
        # after the repo root: the reference's top-level packages are called attack / model / defense, the product's live
        # inside speakerguard_b200/, so there is no clash
        sys.path.append(d)
    return d


def build_reference_xv(p: Dict[str, torch.Tensor], tmp: str, threshold=None):
    """The reference's own ``xv_plda`` with the synthetic parameters ``p`` (speakerguard_b200.synthetic.make_xv_params /
    oracle.make_xv_params), loaded back from the Kaldi-style text files its parsers read (model/utils.py:21-80,
    model/_xv_plda/plda.py:27-51)."""
    import_reference()
    from model._xv_plda.xvecTDNN import xvecTDNN
    from model.xv_plda import xv_plda
    from speakerguard_b200.synthetic import write_xv_model_files
    torch.manual_seed(0)
    net = xvecTDNN(numSpkrs=100)
    sd = net.state_dict()
    for i in range(1, 6):
        sd[f"tdnn{i}.weight"].copy_(p[f"tdnn{i}.weight"])
        sd[f"tdnn{i}.bias"].copy_(p[f"tdnn{i}.bias"])
        getattr(net, f"bn_tdnn{i}").running_mean.copy_(p[f"bn{i}.mean"])
        getattr(net, f"bn_tdnn{i}").running_var.copy_(p[f"bn{i}.var"])
    sd["fc1.weight"].copy_(p["fc1.weight"])
    sd["fc1.bias"].copy_(p["fc1.bias"])
    files = write_xv_model_files(p, tmp)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = xv_plda(net, files["plda.txt"], files["mean.vec"], files["transform.txt"],
                        model_file=files["speaker_model"], threshold=threshold)
    model.eval()
    return model


class ReferenceXv:
    """Stock ``FGSM`` / ``PGD`` of the reference against its stock ``xv_plda`` on the host cores."""

    def __init__(self, p: Dict[str, torch.Tensor], threads: Optional[int] = None):
        self.threads = threads or os.cpu_count() or 1
        torch.set_num_threads(self.threads)
        self._tmp = tempfile.TemporaryDirectory(prefix="sgb200_ref_")
        self.model = build_reference_xv(p, self._tmp.name)
        self.dir = reference_dir()

    def attack_seconds(self, kind: str, x: torch.Tensor, y: torch.Tensor, **kw) -> float:
        """Wall time of one ``attack(x, y)`` call (x [B,1,N] in [-1,1), y [B]) of the unmodified class."""
        from attack.FGSM import FGSM
        from attack.PGD import PGD
        cls = {"FGSM": FGSM, "PGD": PGD}[kind]
        att = cls(self.model, batch_size=x.shape[0], verbose=0, **kw)
        t0 = time.perf_counter()
        adv, success = att.attack(x, y)
        dt = time.perf_counter() - t0
        assert adv.shape == x.shape and len(success) == x.shape[0]
        return dt
