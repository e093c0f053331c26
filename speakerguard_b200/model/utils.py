"""Host-side helpers of the model facade: input-range handling and parsers for the Kaldi-style
text model files (same formats and semantics as the reference's model/utils.py:7-80 and
model/_xv_plda/plda.py:27-51; load-time only, never on the hot path)."""
from __future__ import annotations

import warnings
from typing import List, Tuple

import numpy as np
import torch

BITS = 16


_KNOWN_RANGE = []          # stack of ranges promised by callers (see known_input_range)


class known_input_range:
    """``with known_input_range("scale"):`` - the caller guarantees that every waveform handed to a model inside the block lies
    in that range, so check_input_range skips its data-dependent detection (a device->host read that stalls the host once
    per forward pass).  The attack loops use it: their iterates are clipped to [-1, 1] by construction, for which the
    detection below always answers 'scale'."""

    def __init__(self, kind: str):
        assert kind in ("scale", "origin")
        self.kind = kind

    def __enter__(self):
        _KNOWN_RANGE.append(self.kind)
        return self

    def __exit__(self, *exc):
        _KNOWN_RANGE.pop()
        return False


def check_input_range(x: torch.Tensor, BITS: int = BITS, range_type: str = "scale") -> torch.Tensor:
    """Bring a waveform to the requested range: 'scale' = [-1,1], 'origin' = int16 range.  The
    current range is detected from the data (0.9*max <= 1 and 0.9*min >= -1 means 'scale')."""
    if range_type not in ("scale", "origin"):
        raise AssertionError("range_type must be 'scale' or 'origin'")
    if _KNOWN_RANGE:
        current = _KNOWN_RANGE[-1]
    else:
        lo, hi = torch.aminmax(x.detach())
        lo, hi = torch.stack((lo, hi)).tolist()                       # one device->host read
        current = "scale" if (0.9 * hi <= 1 and 0.9 * lo >= -1) else "origin"
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


# ---- Kaldi text models of the i-vector system (model/_iv_plda/gmm.py:33-81, ivector_extract.py:28-70) ----
class _Tokens:
    """Whitespace token stream over a Kaldi text-mode model file."""

    def __init__(self, path: str):
        with open(path, "r") as f:
            self.t = f.read().split()
        self.i = 0

    def seek(self, tag: str) -> None:
        try:
            self.i = self.t.index(tag, self.i) + 1
        except ValueError:
            raise ValueError(f"tag {tag} not found") from None

    def take_bracketed(self) -> np.ndarray:
        """Numbers of the next '[ ... ]' group (vector, matrix or packed triangle), flattened."""
        while self.t[self.i] != "[":
            self.i += 1                       # e.g. the component count after <M> / <SigmaInv>
        j = self.t.index("]", self.i)
        vals = np.asarray(self.t[self.i + 1:j], dtype=np.float64)
        self.i = j + 1
        return vals


def _unpack_lower(packed: np.ndarray, dim: int) -> np.ndarray:
    """Kaldi SpMatrix text order (row-wise lower triangle) -> dense symmetric [dim, dim]."""
    if packed.size != dim * (dim + 1) // 2:
        raise ValueError(f"packed matrix has {packed.size} entries, expected {dim * (dim + 1) // 2}")
    m = np.zeros((dim, dim), dtype=np.float64)
    m[np.tril_indices(dim)] = packed
    return m + np.tril(m, -1).T


def parse_fgmm_file(path: str):
    """Full-covariance UBM, Kaldi text: <GCONSTS> [C], <WEIGHTS> [C], <MEANS_INVCOVARS> [C,F],
    <INV_COVARS> C packed triangles.  Returns float32 arrays (gconsts, weights, means_invcovars, invcovars [C,F,F])."""
    tk = _Tokens(path)
    tk.seek("<GCONSTS>")
    gconsts = tk.take_bracketed()
    tk.seek("<WEIGHTS>")
    weights = tk.take_bracketed()
    C = gconsts.size
    tk.seek("<MEANS_INVCOVARS>")
    mic = tk.take_bracketed()
    if mic.size % C != 0:
        raise ValueError(f"malformed UBM file {path}: {mic.size} mean entries for {C} components")
    F = mic.size // C
    tk.seek("<INV_COVARS>")
    inv = np.stack([_unpack_lower(tk.take_bracketed(), F) for _ in range(C)])
    f32 = np.float32
    return gconsts.astype(f32), weights.astype(f32), mic.reshape(C, F).astype(f32), inv.astype(f32)


def parse_ivector_extractor_file(path: str):
    """i-vector extractor, Kaldi text: <w_vec> [C], <M> C matrices [F,D], <SigmaInv> C packed triangles,
    <IvectorOffset> scalar.  Returns float32 (T [C,F,D], sigma_inv [C,F,F], offset)."""
    tk = _Tokens(path)
    tk.seek("<w_vec>")
    C = tk.take_bracketed().size
    tk.seek("<M>")
    mats = []
    for _ in range(C):
        start = tk.i
        flat = tk.take_bracketed()
        mats.append((flat, start))
    tk.seek("<SigmaInv>")
    packed = [tk.take_bracketed() for _ in range(C)]
    # F from the packed triangle size, D from the matrix size
    n = packed[0].size
    F = int((np.sqrt(8 * n + 1) - 1) / 2 + 0.5)
    D = mats[0][0].size // F
    T = np.stack([m.reshape(F, D) for m, _ in mats])
    sig = np.stack([_unpack_lower(p, F) for p in packed])
    tk.seek("<IvectorOffset>")
    offset = float(tk.t[tk.i])
    return T.astype(np.float32), sig.astype(np.float32), offset
