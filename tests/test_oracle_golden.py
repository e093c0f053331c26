"""Pins oracle/sg_oracle.py against fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import sg_oracle as O

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def xv():
    return np.load(os.path.join(G, "xv_golden.npz"))


@pytest.fixture(scope="module")
def params():
    return O.make_xv_params(seed=0)


def regen(g, tag, n_pass=None, check_y=True):
    seed, B, N = int(g[f"{tag}.seed"]), int(g[f"{tag}.B"]), int(g[f"{tag}.N"])
    torch.manual_seed(seed)
    x = (torch.rand(B, 1, N) * 2 - 1) * 0.5
    y = torch.randint(0, 10, (B,))
    m = O.num_frames(N)
    torch.manual_seed(seed + 1)
    n = B if n_pass is None else n_pass * B
    d = torch.stack([torch.randn((m, 400)) for _ in range(n)])
    d = d if n_pass is None else d.view(n_pass, B, m, 400)
    assert abs(float(x.double().abs().sum()) - float(g[f"{tag}.x_cks"])) < 1e-9, "input regeneration drifted"
    assert abs(float(d.double().abs().sum()) - float(g[f"{tag}.dither_cks"])) < 1e-6, "dither regeneration drifted"
    assert not check_y or np.array_equal(y.numpy(), g[f"{tag}.y"])
    return x[:, 0], y, d


def test_params_regenerate(xv, params):
    assert abs(O.params_checksum(params) - float(xv["params_cks"])) < 1e-6


@pytest.mark.parametrize("tag", ["fwd2s", "fwd5s", "fwd1p1s"])
def test_forward_stages(xv, params, tag):
    x, y, d = regen(xv, tag)
    o = O.xv_forward(x, params, d, return_all=True)
    assert np.array_equal(o["raw"].numpy(), xv[f"{tag}.raw"])            # same torch ops: bit-exact
    np.testing.assert_allclose(o["feat"].numpy(), xv[f"{tag}.feat"], atol=5e-5, rtol=0)
    np.testing.assert_allclose(o["emb"].numpy(), xv[f"{tag}.emb"], atol=5e-6, rtol=1e-5)
    np.testing.assert_allclose(o["scores"].numpy(), xv[f"{tag}.scores"], atol=2e-4, rtol=0)
    assert np.array_equal(O.decide(o["scores"]).numpy(), xv[f"{tag}.dec"])
    acts = O.tdnn_layers(O.cmvn(O.mfcc(x, d)), params)
    for i in range(1, 6):
        np.testing.assert_allclose(acts[i - 1][0, :48].numpy(), xv[f"{tag}.act{i}"], atol=2e-4, rtol=1e-4)


@pytest.mark.parametrize("tag", ["fwd2s", "fwd1p1s"])
def test_input_gradient(xv, params, tag):
    x, y, d = regen(xv, tag)
    scores, loss, grad, dec = O.xv_loss_and_grad(x, y, params, O.loss_ce, d)
    np.testing.assert_allclose(loss.numpy(), xv[f"{tag}.loss"], atol=2e-4, rtol=1e-4)
    ref = xv[f"{tag}.grad"]
    err = np.abs(grad.numpy() - ref).max(axis=1) / np.abs(ref).max(axis=1)
    assert err.max() < 1e-4, err


def test_cmvn_closed_form_equals_loop():
    g = torch.Generator().manual_seed(3)
    for T in (1, 7, 299, 300, 301, 450, 700):
        f = torch.randn(1, T, 5, generator=g) * 10
        np.testing.assert_allclose(O.cmvn(f).numpy(), O.cmvn_loop(f).numpy(), atol=2e-5, rtol=0)


@pytest.mark.parametrize("tag,kw", [
    ("fgsm", dict(fgsm=True, epsilon=0.002)),
    ("pgd3", dict(epsilon=0.002, step_size=0.0004, max_iter=3)),
    ("pgd3t", dict(epsilon=0.002, step_size=0.0004, max_iter=3, targeted=True)),
    ("cwinf3", dict(epsilon=0.002, step_size=0.0004, max_iter=3, loss_name="Margin")),
])
def test_attack_iterates(xv, params, tag, kw):
    x, y, d = regen(xv, tag, int(xv[f"{tag}.n_pass"]))
    adv, success, info = O.pgd_attack(x, y, params, dither=d, **kw)
    ref = xv[f"{tag}.adv"]
    mism = float((adv.numpy() != ref).mean())
    assert mism < 1e-3, mism          # sign flips only where |grad| ~ 0
    assert success == xv[f"{tag}.success"].tolist()


def test_audionet(params):
    g = np.load(os.path.join(G, "audionet_golden.npz"))
    p = O.make_audionet_params(seed=0, num_class=251)
    for tag in ("an1s", "an3s"):
        B, N = int(g[f"{tag}.B"]), int(g[f"{tag}.N"])
        torch.manual_seed(4321)
        x = ((torch.rand(B, 1, N) * 2 - 1) * 0.5)[:, 0]
        assert abs(float(x.double().abs().sum()) - float(g[f"{tag}.x_cks"])) < 1e-9
        y = torch.tensor(g[f"{tag}.y"])
        xr = x.clone().requires_grad_(True)
        o = O.audionet_forward(xr, p, return_all=True)
        np.testing.assert_allclose(o["feat"].transpose(1, 2).detach().numpy(), g[f"{tag}.feat"], atol=2e-4, rtol=0)
        np.testing.assert_allclose(o["logits"].detach().numpy(), g[f"{tag}.logits"], atol=1e-5, rtol=1e-5)
        loss = O.loss_margin(o["logits"], y, targeted=True, clip_max=True)
        loss.backward(torch.ones_like(loss))
        np.testing.assert_allclose(loss.detach().numpy(), g[f"{tag}.loss"], atol=1e-5, rtol=1e-5)
        ref = g[f"{tag}.grad"]
        err = np.abs(xr.grad.numpy() - ref).max(axis=1) / np.abs(ref).max(axis=1)
        assert err.max() < 1e-4, err


def test_feco_conditional():
    g = np.load(os.path.join(G, "feco_golden.npz"))
    feat = torch.tensor(g["feco.feat"], requires_grad=True)
    out = O.feco_means(feat, g["feco.ids"], int(g["feco.k"]), force=True)
    np.testing.assert_allclose(out.detach().numpy(), g["feco.out"], atol=1e-6, rtol=1e-6)
    (out * torch.tensor(g["feco.w"][0])).sum().backward()
    # the reference fixture fed the same feat twice (batch 2): its grad sums both batch rows
    w = torch.tensor(g["feco.w"])
    feat2 = torch.tensor(g["feco.feat"], requires_grad=True)
    tot = sum((O.feco_means(feat2, g["feco.ids"], int(g["feco.k"])) * w[i]).sum() for i in range(2))
    tot.backward()
    np.testing.assert_allclose(feat2.grad.numpy(), g["feco.grad"], atol=1e-6, rtol=1e-5)


def test_cw2_oracle_vs_reference_outcome():
    """CW2 is chaotic at the last bit (Adam's normalised steps), so the oracle is pinned on outcomes:
    success flags equal, best-example L2 distortion within 15 %."""
    g = np.load(os.path.join(G, "audionet_golden.npz"))
    p = O.make_audionet_params(seed=0, num_class=251)
    B, N = int(g["cw2.B"]), int(g["cw2.N"])
    torch.manual_seed(777)
    x = ((torch.rand(B, 1, N) * 2 - 1) * 0.5)[:, 0]
    y = torch.tensor(g["cw2.y"])
    xa, suc, _ = O.cw2_attack(x, y, lambda z: O.audionet_forward(z, p), targeted=False, initial_const=1e2,
                              binary_search_steps=2, max_iter=40, stop_early=True, stop_early_iter=10, lr=1e-2)
    assert suc == g["cw2.success"].tolist()
    l2 = lambda a: (a - x).pow(2).sum(1).numpy()
    np.testing.assert_allclose(l2(xa), l2(torch.tensor(g["cw2.adv"])), rtol=0.15)


# ---- i-vector system (BASELINE config 5) -----------------------------------------------------------
def iv_regen(g, n_pass=None, seeds=(606, 607)):
    B, N = int(g["iv.B"]), int(g["iv.N"])
    torch.manual_seed(seeds[0])
    x = (torch.rand(B, 1, N) * 2 - 1) * 0.5
    m = O.num_frames(N)
    torch.manual_seed(seeds[1])
    n = B if n_pass is None else n_pass * B
    d = torch.stack([torch.randn((m, 400)) for _ in range(n)])
    d = d if n_pass is None else d.view(n_pass, B, m, 400)
    assert abs(float(x.double().abs().sum()) - float(g["iv.x_cks"])) < 1e-9, "input regeneration drifted"
    return x[:, 0], torch.from_numpy(g["iv.y"]), d


@pytest.fixture(scope="module")
def ivg():
    return np.load(os.path.join(G, "iv_golden.npz"))


def test_iv_forward_stages(ivg):
    p = O.make_iv_params(seed=0)
    x, y, d = iv_regen(ivg)
    assert abs(float(d.double().abs().sum()) - float(ivg["iv.dither_cks"])) < 1e-6
    o = O.iv_forward(x, p, d, return_all=True)
    assert np.array_equal(o["raw"].numpy(), ivg["iv.raw"])
    np.testing.assert_allclose(o["delta"].numpy(), ivg["iv.delta"], atol=5e-6, rtol=0)
    np.testing.assert_allclose(o["feat"].numpy(), ivg["iv.feat"], atol=1e-5, rtol=0)
    np.testing.assert_allclose(o["emb"].numpy(), ivg["iv.emb"], atol=1e-5, rtol=1e-5)
    np.testing.assert_allclose(o["scores"].numpy(), ivg["iv.scores"], atol=1e-4, rtol=0)


def test_iv_sv_margin_gradient(ivg):
    p = O.make_iv_params(seed=0)
    x, y, d = iv_regen(ivg)
    thr = float(ivg["iv.thr"])
    loss_fn, _ = O.resolve_loss("Margin", False, 0.0, "SV", thr, clip_max=False)
    scores, loss, grad, dec = O.xv_loss_and_grad(x, y, p, loss_fn, d, system="iv")
    np.testing.assert_allclose(loss.numpy(), ivg["iv.loss"], atol=1e-4, rtol=0)
    ref = ivg["iv.grad"]
    err = np.abs(grad.numpy() - ref).max(axis=1) / np.abs(ref).max(axis=1)
    assert err.max() < 1e-3, err


def test_iv_pgd_sv(ivg):
    p = dict(O.make_iv_params(seed=0))
    p["threshold"] = float(ivg["iv.thr"])
    x, y, _ = iv_regen(ivg)
    _, _, d = iv_regen(ivg, n_pass=3, seeds=(606, 608))
    assert abs(float(d.double().abs().sum()) - float(ivg["ivpgd.dither_cks"])) < 1e-6
    adv, success, _ = O.pgd_attack(x, y, p, epsilon=0.002, step_size=0.0004, max_iter=2, task="SV", dither=d, system="iv")
    assert success == ivg["ivpgd.success"].tolist()
    ref = ivg["ivpgd.adv"]
    frac = float((np.abs(adv.numpy() - ref) < 1e-7).mean())
    assert frac > 0.995, frac


def test_iv_delta_filters_and_solver_adjoint():
    """add_delta equals two passes of the first-order filter away from the edges; the i-vector of zero
    statistics is the prior (offset in the first coordinate, removed again: ivector_extract.py:108-113)."""
    s = O.delta_scales()
    assert [t.numel() for t in s] == [1, 7, 13]
    np.testing.assert_allclose(s[1].numpy(), np.arange(-3, 4) / 28.0, atol=1e-7)
    p = O.make_iv_params(seed=0)
    C, Fd = p["gmm.gconsts"].shape[0], p["gmm.means_invcovars"].shape[1]
    iv = O.ivector(torch.zeros(C), torch.zeros(C, Fd), p)
    np.testing.assert_allclose(iv.numpy(), 0.0, atol=1e-6)


# ---- AudioNet training step ----------------------------------------------------
def test_audionet_train_step_against_reference():
    g = np.load(os.path.join(G, "antrain_golden.npz"))
    p = O.make_audionet_params(seed=0, num_class=251)
    B, N = int(g["antrain.B"]), int(g["antrain.N"])
    torch.manual_seed(9091)
    x = (torch.rand(B, 1, N) * 2 - 1) * 0.5
    assert abs(float(x.double().abs().sum()) - float(g["antrain.x_cks"])) < 1e-9
    y = torch.from_numpy(g["antrain.y"])
    o = O.audionet_train_step(x[:, 0], y, p)
    np.testing.assert_allclose(o["logits"].numpy(), g["antrain.logits"], atol=1e-4, rtol=0)
    np.testing.assert_allclose(float(o["loss"]), float(g["antrain.loss"]), atol=1e-5)
    ref = g["antrain.xgrad"]
    assert np.abs(o["xgrad"].numpy() - ref).max() < 1e-4 * np.abs(ref).max()
    for k in O.AN_PARAM_KEYS:
        r = g[f"antrain.grad.{k}"]
        got = o["grads"][k].numpy()
        if k.endswith(".bias") and k != "fc.bias":
            # a conv bias in front of a train-mode BatchNorm has zero gradient; both sides hold round-off only
            assert np.abs(got).max() < 1e-5 and np.abs(r).max() < 1e-5
            continue
        assert np.abs(got - r).max() < 2e-3 * np.abs(r).max(), k       # fp32 reductions over B*T in different orders
    for n in O.AN_BN_NAMES:
        np.testing.assert_allclose(o["stats"][f"{n}.bn_mean"].numpy(), g[f"antrain.stat.{n}.bn_mean"], atol=1e-5, rtol=1e-5)
        np.testing.assert_allclose(o["stats"][f"{n}.bn_var"].numpy(), g[f"antrain.stat.{n}.bn_var"], atol=1e-5, rtol=1e-4)
        # first Adam step moves every weight by lr * sign(grad) (up to eps): compare where the gradient is not round-off
        for key in ("weight", "bn_gamma"):
            r = g[f"antrain.new.{n}.{key}"]
            gr = np.abs(g[f"antrain.grad.{n}.{key}"])
            sel = gr > 1e-6
            assert np.abs(o["params"][f"{n}.{key}"].numpy() - r)[sel].max() < 1e-5     # 1 % of one Adam step (lr 1e-3)


# ---- round 2: long attack and the OSI / default-loss-name sign (tests/golden/xv_long_golden.npz) ---------------------
@pytest.fixture(scope="module")
def xvl():
    return np.load(os.path.join(G, "xv_long_golden.npz"))


@pytest.mark.parametrize("tag", ["pgd3osi", "pgd3osit"])
def test_osi_entropy_name_keeps_the_cross_entropy_sign(xvl, params, tag):
    """resolve_loss(task='OSI', loss_name='Entropy') -> margin loss, grad_sign +1 / -1 by `targeted` (attack/utils.py:107-114)."""
    x, y, d = regen(xvl, tag, 4, check_y=False)
    y[-1] = -1
    assert np.array_equal(y.numpy(), xvl[f"{tag}.y"])
    p = dict(params)
    p["threshold"] = float(xvl[f"{tag}.thr"])
    targeted = bool(int(xvl[f"{tag}.targeted"]))
    adv, success, _ = O.pgd_attack(x, y, p, dither=d, epsilon=0.002, step_size=0.0004, max_iter=3, loss_name="Entropy",
                                   targeted=targeted, task="OSI")
    assert success == xvl[f"{tag}.success"].tolist()
    assert float((adv != torch.tensor(xvl[f"{tag}.adv"])).float().mean()) < 2e-3


def test_pgd100_outcome(xvl, params):
    """The reference's PGD-100 at B = 8 x 3 s: the oracle reproduces the success list and the per-utterance SNR (the
    individual samples are chaotic after 100 sign steps: ~95 % end up bit-identical)."""
    torch.set_num_threads(max(torch.get_num_threads(), min(os.cpu_count() or 1, 16)))
    tag = "pgd100"
    x, y, d = regen(xvl, tag, int(xvl[f"{tag}.n_pass"]))
    adv, success, _ = O.pgd_attack(x, y, params, dither=d, epsilon=0.002, step_size=0.0004, max_iter=100)
    assert success == xvl[f"{tag}.success"].tolist()
    delta = (adv - x).double()
    snr = (10 * torch.log10(x.double().pow(2).sum(1) / delta.pow(2).sum(1))).numpy()
    assert np.abs(snr - xvl[f"{tag}.snr_db"]).max() < 0.05
    assert float((adv == torch.tensor(xvl[f"{tag}.adv"])).float().mean()) > 0.9
