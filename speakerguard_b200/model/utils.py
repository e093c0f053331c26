"""Host-side helpers of the model facade: input-range handling and parsers for the Kaldi-style
text model files (same formats and semantics as the reference's model/utils.py:7-80 and
model/_xv_plda/plda.py:27-51; load-time only, never on the hot path)."""
from __future__ import annotations

import warnings
from typing import List, Tuple

import numpy as np
import torch

BITS = 16


def check_input_range(x: torch.Tensor, BITS: int = BITS, range_type: str = "scale") -> torch.Tensor:
    """Bring a waveform to the requested range: 'scale' = [-1,1], 'origin' = int16 range.  The
    current range is detected from the data (0.9*max <= 1 and 0.9*min >= -1 means 'scale')."""
    if range_type not in ("scale", "origin"):
        raise AssertionError("range_type must be 'scale' or 'origin'")
    lo, hi = torch.aminmax(x.detach())
    current = "scale" if (0.9 * float(hi) <= 1 and 0.9 * float(lo) >= -1) else "origin"
    if current == range_type:
        return x
    full = float(2 ** (BITS - 1))
    return x * full if range_type == "origin" else x / full


def _bracket_floats(line: str) -> List[float]:
    body = line.replace("<Plda>", " ").replace("[", " ").replace("]", " ")
    return [float(t) for t in body.split()]


def parse_plda_file(path: str) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Kaldi text PLDA: '<Plda>  [ mean ]', ' [', L rows of the transform (last ends with ']'),
    ' [ psi ]'.  Returns (mean [L], transform [L,L], psi [L]) as float32."""
    with open(path, "r") as f:
        lines = f.read().splitlines()
    mean = np.asarray(_bracket_floats(lines[0]), dtype=np.float32)
    L = mean.shape[0]
    rows = [_bracket_floats(lines[2 + i]) for i in range(L)]
    transform = np.asarray(rows, dtype=np.float32)
    psi = np.asarray(_bracket_floats(lines[2 + L]), dtype=np.float32)
    if transform.shape != (L, L) or psi.shape != (L,):
        raise ValueError(f"malformed PLDA file {path}: transform {transform.shape}, psi {psi.shape}, dim {L}")
    return mean, transform, psi


def parse_mean_file(path: str, device) -> torch.Tensor:
    with open(path, "r") as f:
        vals = _bracket_floats(f.readline())
    return torch.tensor(vals, dtype=torch.float32, device=device)


def parse_transform_mat_file(path: str, device) -> torch.Tensor:
    """LDA transform: first line is the opening bracket, then one row per line, ']' closes the last."""
    with open(path, "r") as f:
        lines = f.read().splitlines()[1:]
    rows = [_bracket_floats(ln) for ln in lines if ln.strip() and ln.strip() != "]"]
    return torch.tensor(np.asarray(rows, dtype=np.float64), dtype=torch.float32, device=device)


def parse_enroll_model_file(path: str, device):
    """Lines 'spk_id emb_path z_norm_mean z_norm_std'; emb_path holds a torch-saved [1,L] tensor."""
    info = np.loadtxt(path, dtype=str, comments=None)
    if info.ndim == 1:
        info = info[np.newaxis, :]
    spk_ids = list(info[:, 0])
    z_means = torch.tensor(info[:, 2].astype(np.float32), device=device)
    z_stds = torch.tensor(info[:, 3].astype(np.float32), device=device)
    embs = [torch.load(p, map_location=device) for p in info[:, 1]]
    enroll = torch.cat(embs, dim=0).to(torch.float32)
    if len(spk_ids) > 1:
        warnings.warn("model_file holds more than one speaker: make sure the task is not SV "
                      "(SV expects exactly one enrolled speaker).")
    return len(spk_ids), spk_ids, z_means, z_stds, enroll
