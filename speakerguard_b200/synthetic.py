"""Synthetic x-vector / PLDA system for benchmarks and smoke tests (no network, no checkpoints):
random-init TDNN in PyTorch's default Conv1d/Linear initialisation, synthetic PLDA / LDA /
enrolment tensors, plus writers for the Kaldi-style text model files the
drop-in ``xv_plda`` constructor parses."""
from __future__ import annotations

import math
import os
from typing import Dict

import torch

TDNN_SPEC = [(30, 512, 5, 1), (512, 512, 5, 2), (512, 512, 7, 3), (512, 512, 1, 1), (512, 1500, 1, 1)]


def _round6(t: torch.Tensor) -> torch.Tensor:
    return torch.tensor([float("%.6f" % v) for v in t.double().flatten().tolist()], dtype=torch.float32).reshape(t.shape)


def make_xv_params(seed: int = 0, L: int = 200, S: int = 10, num_spk: int = 100, random_bn: bool = True
                   ) -> Dict[str, torch.Tensor]:
    torch.manual_seed(seed)
    p: Dict[str, torch.Tensor] = {}
    for i, (ci, co, k, d) in enumerate(TDNN_SPEC, 1):
        conv = torch.nn.Conv1d(ci, co, k, dilation=d)
        p[f"tdnn{i}.weight"], p[f"tdnn{i}.bias"] = conv.weight.detach().clone(), conv.bias.detach().clone()
    fc1 = torch.nn.Linear(3000, 512)
    torch.nn.Linear(512, 512)
    torch.nn.Linear(512, num_spk)
    p["fc1.weight"], p["fc1.bias"] = fc1.weight.detach().clone(), fc1.bias.detach().clone()
    g = torch.Generator().manual_seed(seed + 7919)
    for i, (_, co, _, _) in enumerate(TDNN_SPEC, 1):
        p[f"bn{i}.mean"] = 0.1 * torch.randn(co, generator=g) if random_bn else torch.zeros(co)
        p[f"bn{i}.var"] = 0.5 + torch.rand(co, generator=g) if random_bn else torch.ones(co)
    p["plda.mean"] = _round6(0.1 * torch.randn(L, generator=g))
    p["plda.transform"] = _round6(torch.randn(L, L, generator=g) / math.sqrt(L))
    p["plda.psi"] = _round6(torch.randn(L, generator=g).abs() + 0.1)
    p["emb_mean"] = _round6(0.1 * torch.randn(512, generator=g))
    p["lda"] = _round6(torch.randn(L, 513, generator=g) / math.sqrt(512))
    p["enroll"] = torch.randn(S, L, generator=g)
    return p


def state_dict_of(p: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """x-vector TDNN state dict with the reference's key names (model/_xv_plda/xvecTDNN.py:16-43)."""
    sd = {}
    for i in range(1, 6):
        sd[f"tdnn{i}.weight"], sd[f"tdnn{i}.bias"] = p[f"tdnn{i}.weight"], p[f"tdnn{i}.bias"]
        sd[f"bn_tdnn{i}.running_mean"], sd[f"bn_tdnn{i}.running_var"] = p[f"bn{i}.mean"], p[f"bn{i}.var"]
    sd["fc1.weight"], sd["fc1.bias"] = p["fc1.weight"], p["fc1.bias"]
    return sd


def _vec(v) -> str:
    return " [ " + " ".join("%.6f" % x for x in v) + " ]\n"


def _rows(M) -> str:
    return "".join("  " + " ".join("%.6f" % x for x in r) + (" \n" if i < len(M) - 1 else " ]\n")
                   for i, r in enumerate(M))


def write_xv_model_files(p: Dict[str, torch.Tensor], out_dir: str) -> Dict[str, str]:
    """Writes plda.txt, mean.vec, transform.txt, speaker_model (+ one torch-saved embedding per
    speaker) in the formats of plda.py:27-51 and model/utils.py:21-80; returns their paths."""
    os.makedirs(out_dir, exist_ok=True)
    f = {k: os.path.join(out_dir, k) for k in ("plda.txt", "mean.vec", "transform.txt", "speaker_model")}
    with open(f["plda.txt"], "w") as fh:
        fh.write("<Plda> " + _vec(p["plda.mean"].tolist()) + " [\n" + _rows(p["plda.transform"].tolist())
                 + _vec(p["plda.psi"].tolist()) + "</Plda> \n")
    with open(f["mean.vec"], "w") as fh:
        fh.write(_vec(p["emb_mean"].tolist()))
    with open(f["transform.txt"], "w") as fh:
        fh.write(" [\n" + _rows(p["lda"].tolist()))
    with open(f["speaker_model"], "w") as fh:
        for s in range(p["enroll"].shape[0]):
            path = os.path.join(out_dir, f"spk{s}.emb")
            torch.save(p["enroll"][s:s + 1].clone(), path)
            fh.write(f"spk{s} {path} 0.0 1.0\n")
    return f


def synthetic_batch(B: int, N: int, S: int = 10, seed: int = 1234):
    """x [B,1,N] uniform in [-0.5, 0.5), y [B] in [0,S)  (SURVEY.md 8(d))."""
    g = torch.Generator().manual_seed(seed)
    x = (torch.rand(B, 1, N, generator=g) * 2 - 1) * 0.5
    y = torch.randint(0, S, (B,), generator=g)
    return x, y


# ---- i-vector system (BASELINE config 5) ---------------------------------------------------------
def make_iv_params(seed: int = 0, C: int = 64, F: int = 72, D: int = 40, L: int = 30, S: int = 1) -> Dict[str, torch.Tensor]:
    """Synthetic full-covariance UBM + i-vector extractor + back-end (config 5: C=2048, F=72, D=400, L=200).
    SPD inverse covariances with per-dimension std around 3-4, the range of CMVN'd MFCC+delta features."""
    g = torch.Generator().manual_seed(seed + 31337)
    p: Dict[str, torch.Tensor] = {}
    mu = 3.0 * torch.randn(C, F, generator=g)
    A = torch.randn(C, F, F, generator=g) / math.sqrt(F)
    invcov = 0.05 * (A @ A.transpose(1, 2)) + 0.06 * torch.eye(F)
    invcov = 0.5 * (invcov + invcov.transpose(1, 2))
    w = torch.softmax(torch.randn(C, generator=g), 0)
    p["gmm.invcovars"] = invcov
    p["gmm.means_invcovars"] = (invcov @ mu.unsqueeze(-1)).squeeze(-1)
    logdet = torch.linalg.slogdet(invcov.double())[1].float()
    quad = (mu.unsqueeze(1) @ invcov @ mu.unsqueeze(-1)).flatten()
    p["gmm.gconsts"] = torch.log(w) - 0.5 * (F * math.log(2 * math.pi) - logdet + quad)
    p["gmm.weights"] = w
    p["ive.T"] = 0.3 * torch.randn(C, F, D, generator=g)
    Bm = torch.randn(C, F, F, generator=g) / math.sqrt(F)
    sig = 0.05 * (Bm @ Bm.transpose(1, 2)) + 0.06 * torch.eye(F)
    p["ive.sigma_inv"] = 0.5 * (sig + sig.transpose(1, 2))
    p["ive.offset"] = torch.tensor(5.0)
    p["plda.mean"] = _round6(0.1 * torch.randn(L, generator=g))
    p["plda.transform"] = _round6(torch.randn(L, L, generator=g) / math.sqrt(L))
    p["plda.psi"] = _round6(torch.randn(L, generator=g).abs() + 0.1)
    p["emb_mean"] = _round6(0.1 * torch.randn(D, generator=g))
    p["lda"] = _round6(torch.randn(L, D + 1, generator=g) / math.sqrt(D))
    p["enroll"] = torch.randn(S, L, generator=g)
    return p


def _packed_rows(M) -> str:
    """Kaldi SpMatrix text: row i holds the i+1 lower-triangle entries; ']' closes the last row."""
    n = len(M)
    return "".join(" ".join("%.9g" % M[i][j] for j in range(i + 1)) + (" \n" if i < n - 1 else " ]\n") for i in range(n))


def _rows9(M) -> str:
    return "".join("  " + " ".join("%.9g" % x for x in r) + (" \n" if i < len(M) - 1 else " ]\n") for i, r in enumerate(M))


def _vec9(v) -> str:
    return " [ " + " ".join("%.9g" % x for x in v) + " ]\n"


def write_iv_model_files(p: Dict[str, torch.Tensor], out_dir: str) -> Dict[str, str]:
    """final_ubm.txt / final_ie.txt in the Kaldi text layout the reference's line parsers expect
    (gmm.py:33-81, ivector_extract.py:28-70) plus the back-end files of write_xv_model_files."""
    f = write_xv_model_files(p, out_dir)
    f["final_ubm.txt"], f["final_ie.txt"] = os.path.join(out_dir, "final_ubm.txt"), os.path.join(out_dir, "final_ie.txt")
    C = p["gmm.gconsts"].shape[0]
    with open(f["final_ubm.txt"], "w") as fh:
        fh.write("<FullGMM> \n<GCONSTS> " + _vec9(p["gmm.gconsts"].tolist()))
        fh.write("<WEIGHTS> " + _vec9(p["gmm.weights"].tolist()))
        fh.write("<MEANS_INVCOVARS>  [\n" + _rows9(p["gmm.means_invcovars"].tolist()))
        fh.write("<INV_COVARS>  [\n")
        for c in range(C):
            if c:
                fh.write(" [\n")
            fh.write(_packed_rows(p["gmm.invcovars"][c].tolist()))
        fh.write("</FullGMM> \n")
    with open(f["final_ie.txt"], "w") as fh:
        fh.write("<IvectorExtractor> \n<w> [ ]\n<w_vec> " + _vec9([1.0 / C] * C))
        fh.write(f"<M> {C}  [\n")
        for c in range(C):
            if c:
                fh.write(" [\n")
            fh.write(_rows9(p["ive.T"][c].tolist()))
        fh.write(f"<SigmaInv> {C}  [\n")
        for c in range(C):
            if c:
                fh.write(" [\n")
            fh.write(_packed_rows(p["ive.sigma_inv"][c].tolist()))
        fh.write("<IvectorOffset> %.9g \n</IvectorExtractor> \n" % float(p["ive.offset"]))
    return f
