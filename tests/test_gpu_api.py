"""GPU tests of the drop-in Python API (same names / arguments / behaviour as the reference's
model.xv_plda, model.defended_model, attack.FGSM/PGD/CWinf, adaptive_attack.EOT)."""
import os

import numpy as np
import pytest
import torch

from oracle import sg_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def params():
    return O.make_xv_params(seed=0)


@pytest.fixture(scope="module")
def model(params, tmp_path_factory):
    from speakerguard_b200.model.xv_plda import xv_plda
    from speakerguard_b200.synthetic import state_dict_of, write_xv_model_files
    d = str(tmp_path_factory.mktemp("xv"))
    f = write_xv_model_files(params, d)
    m = xv_plda(state_dict_of(params), f["plda.txt"], f["mean.vec"], f["transform.txt"], model_file=f["speaker_model"],
                device="cuda:0", dither="off")
    m.eval()
    return m


def wave(B, N, seed=1234):
    torch.manual_seed(seed)
    return (torch.rand(B, 1, N) * 2 - 1) * 0.5


def test_text_parsers_round_trip(model, params):
    assert torch.equal(model.emb_mean.cpu(), params["emb_mean"])
    assert torch.equal(model.transform_mat.cpu(), params["lda"])
    assert torch.equal(model.enroll_embs.cpu(), params["enroll"])
    assert model.num_spks == 10 and model.spk_ids[3] == "spk3"
    assert model.allowed_flags == [0, 1, 2] and model.range_type == "origin"
    assert model.threshold == -np.inf


def test_model_methods_match_oracle(model, params):
    x = wave(3, 32000)
    ref = O.xv_forward(x[:, 0], params, None, return_all=True)
    xc = x.cuda()
    raw = model.compute_feat(xc, flag=1)
    feat = model.compute_feat(xc, flag=2)
    assert raw.shape == (3, 200, 30)
    assert float((raw.cpu() - ref["raw"]).abs().max()) < 1e-4 * float(ref["raw"].abs().max())
    assert float((feat.cpu() - ref["feat"]).abs().max()) < 1e-4 * float(ref["feat"].abs().max())
    assert float((model.comput_feat_from_feat(raw, 1, 2).cpu() - ref["feat"]).abs().max()) < 1e-3
    for flag, inp in [(0, xc), (1, raw), (2, feat)]:
        emb = model.embedding(inp, flag=flag)
        assert float((emb.cpu() - ref["emb"]).abs().max()) < 1e-4 * float(ref["emb"].abs().max())
        dec, scores = model.make_decision(inp, flag=flag)
        assert float((scores.cpu() - ref["scores"]).abs().max()) < 2e-4 * float(ref["scores"].abs().max())
        assert torch.equal(dec.cpu(), O.decide(ref["scores"]))
    s2, e2 = model(xc, return_emb=True)
    assert s2.shape == (3, 10) and e2.shape == (3, 200)
    # int16-range input is detected and handled like [-1,1] input (model/utils.py:7-19)
    s3 = model.score(xc * 32768.0)
    assert float((s3 - s2).abs().max()) < 1e-3
    # custom enrolment set and SV-style threshold
    en = torch.randn(1, 200, generator=torch.Generator().manual_seed(1)).cuda()
    s4 = model.score(xc, enroll_embs=en)
    assert s4.shape == (3, 1)
    old = model.threshold
    model.threshold = float(ref["scores"].max(1)[0].median())
    dec_t, _ = model.make_decision(xc)
    assert torch.equal(dec_t.cpu(), O.decide(ref["scores"], model.threshold))
    model.threshold = old


def test_forward_only_fast_path_equals_stagewise(model):
    """make_decision on a tensor that needs no gradient takes the fused sg_xv_forward path and
    must equal the autograd (stage-by-stage) path bit for bit (dither off)."""
    x = wave(3, 24000, seed=21).cuda()
    d1, s1 = model.make_decision(x)                           # fused
    xg = x.clone().requires_grad_(True)
    d2, s2 = model.make_decision(xg)                          # stage-wise autograd path
    assert torch.equal(d1, d2) and torch.equal(s1, s2.detach())
    with torch.no_grad():
        d3, s3 = model.make_decision(xg)                      # no_grad -> fused again
    assert torch.equal(s3, s1)
    s4, e4 = model(x, return_emb=True)
    assert torch.equal(s4, s1) and e4.shape == (3, 200)


def test_autograd_through_the_stages(model, params):
    """loss.backward() through the CUDA stages == oracle autograd (what EOT.forward relies on)."""
    from speakerguard_b200.attack.utils import resolve_loss
    from tests import kink
    x = wave(2, 24000, seed=5)
    y = torch.tensor([4, 9])
    xc = x.cuda().requires_grad_(True)
    loss_mod, gs = resolve_loss("Margin", targeted=True, task="CSI", clip_max=False)
    dec, scores = model.make_decision(xc)
    loss = loss_mod(scores, y.cuda())
    loss.backward(torch.ones_like(loss))
    fn, gs_o = O.resolve_loss("Margin", True, 0.0, "CSI", None, False)
    assert gs == gs_o == -1
    for b in range(2):
        e, near, flipped = kink.resolve(kink.xv_input_grad_fn(x[b:b + 1, 0], y[b:b + 1], params, fn, None), xc.grad[b:b + 1, 0].cpu())
        assert e < 1e-4


def test_pgd_public_api_fused_equals_generic(model, params):
    """attack(x, y): the fused device loop and the generic EOT/autograd loop produce the same
    adversarial examples (dither off => deterministic), and both respect the reference contract."""
    from speakerguard_b200.attack.PGD import PGD
    x, y = wave(3, 32000, seed=8).cuda(), torch.tensor([1, 2, 3]).cuda()
    att = PGD(model, epsilon=0.002, step_size=0.0004, max_iter=3, batch_size=2, verbose=0)
    adv_f, suc_f = att.attack(x, y)
    att.use_fused = False
    adv_g, suc_g = att.attack(x, y)
    assert adv_f.shape == x.shape and isinstance(suc_f, list) and len(suc_f) == 3
    assert float((adv_f - x).abs().max()) <= 0.002 + 1e-7
    assert float((adv_f != adv_g).float().mean()) < 1e-3 and suc_f == suc_g
    xa, suc_o, _ = O.pgd_attack(x.cpu()[:, 0], y.cpu(), params, epsilon=0.002, step_size=0.0004, max_iter=3)
    assert float((adv_f.cpu()[:, 0] != xa).float().mean()) < 2e-2 and suc_f == suc_o


def test_fgsm_cwinf_random_init_and_asserts(model):
    from speakerguard_b200.attack.CWinf import CWinf
    from speakerguard_b200.attack.FGSM import FGSM
    from speakerguard_b200.attack.PGD import PGD
    x, y = wave(2, 16000, seed=9).cuda(), torch.tensor([0, 5]).cuda()
    adv, suc = FGSM(model, epsilon=0.002, batch_size=2, verbose=0).attack(x, y)
    d = (adv - x).abs()
    assert float(d.max()) <= 0.002 + 1e-7 and float((d > 0.0019).float().mean()) > 0.99
    adv2, suc2 = CWinf(model, epsilon=0.002, max_iter=2, batch_size=2, verbose=0).attack(x, y)
    assert float((adv2 - x).abs().max()) <= 0.002 + 1e-7
    np.random.seed(0)
    adv3, suc3 = PGD(model, epsilon=0.002, max_iter=2, num_random_init=2, batch_size=2, verbose=0).attack(x, y)
    assert float((adv3 - x).abs().max()) <= 0.002 + 1e-7
    with pytest.raises(AssertionError):
        PGD(model, verbose=0).attack(x * 4, y)                  # outside [-1, 1)
    with pytest.raises(AssertionError):
        PGD(model, verbose=0).attack(torch.cat([x, x], 1), y)    # stereo


def test_eot_wrapper_and_defended_model(model):
    """EOT over a randomised feature-level defense through defended_model (generic path)."""
    from speakerguard_b200.adaptive_attack.EOT import EOT
    from speakerguard_b200.attack.PGD import PGD
    from speakerguard_b200.attack.utils import resolve_loss
    from speakerguard_b200.model.defended_model import defended_model
    x, y = wave(2, 16000, seed=10).cuda(), torch.tensor([3, 7]).cuda()
    gen = torch.Generator(device="cuda").manual_seed(0)
    noise_defense = lambda f: f + 0.05 * torch.randn(f.shape, device=f.device, generator=gen)
    dm = defended_model(model, defense=[[1, noise_defense]], order="sequential")
    loss_mod, _ = resolve_loss("Entropy")
    scores, loss, grad, decisions = EOT(dm, loss_mod, 4, 2, True)(x, y)
    assert scores.shape == (2, 10) and loss.shape == (2,) and grad.shape == x.shape
    assert len(decisions) == 2 and len(decisions[0]) == 4
    s0, l0, g0, d0 = EOT(dm, loss_mod, 4, 2, True)(x, y, 1, 1, False)
    assert g0 is None and len(d0[0]) == 1
    adv, suc = PGD(dm, epsilon=0.002, max_iter=2, batch_size=2, EOT_size=4, EOT_batch_size=2, verbose=0).attack(x, y)
    assert float((adv - x).abs().max()) <= 0.002 + 1e-7 and len(suc) == 2
    # no defense: defended_model is transparent and the fused path is taken
    from speakerguard_b200.attack.FGSM import fused_target
    plain = defended_model(model)
    assert fused_target(plain) == (model, None) and fused_target(dm) == (None, None)
    d1, s1 = plain.make_decision(x)
    d2, s2 = model.make_decision(x)
    assert torch.equal(d1, d2) and torch.equal(s1, s2)


def test_dither_modes(model):
    x = wave(2, 16000, seed=11).cuda()
    model.dither = "philox"
    a = model.score(x)
    b = model.score(x)
    assert not torch.equal(a, b) and float((a - b).abs().max()) < 1e-2          # fresh noise per pass
    model.dither = "torch"
    torch.manual_seed(3)
    c = model.score(x)
    torch.manual_seed(3)
    d = model.score(x)
    assert torch.equal(c, d)                                                       # same torch seed -> same pass
    model.dither = lambda B, m: torch.zeros(B, m, 400)
    e = model.score(x)
    model.dither = "off"
    f = model.score(x)
    assert torch.equal(e, f)
