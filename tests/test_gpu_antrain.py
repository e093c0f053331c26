"""GPU tests of AudioNet in training mode through the C-ABI and the drop-in class:
one optimisation step as adver_train.py:216-221 runs it, against the oracle (fp64) and the fixture produced
by the reference's own audionet_csine in train() mode (tests/golden/make_golden.py --only-antrain)."""
import os

import numpy as np
import pytest
import torch

from oracle import sg_oracle as O

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def rel(a, b):
    a, b = torch.as_tensor(a).detach().double().cpu(), torch.as_tensor(b).detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.fixture(scope="module")
def params():
    return O.make_audionet_params(seed=0, num_class=251)


def make_model(params, train=True):
    from speakerguard_b200.model.audionet_csine import audionet_csine
    m = audionet_csine(params=params, device="cuda:0")
    return m.train() if train else m


def grads_of(model):
    names = ["conv1"] + [f"conv{i}" for i in range(2, 9)]
    g = {}
    for n in names:
        seq = getattr(model, n)
        g[f"{n}.weight"], g[f"{n}.bias"] = seq[0].weight.grad, seq[0].bias.grad
        g[f"{n}.bn_gamma"], g[f"{n}.bn_beta"] = seq[1].weight.grad, seq[1].bias.grad
    g["fc.weight"], g["fc.bias"] = model.fc.weight.grad, model.fc.bias.grad
    return g


def test_state_dict_is_reference_compatible(params):
    m = make_model(params, train=False)
    keys = set(m.state_dict().keys())
    for n in ["conv1"] + [f"conv{i}" for i in range(2, 9)]:
        for k in ("0.weight", "0.bias", "1.weight", "1.bias", "1.running_mean", "1.running_var", "1.num_batches_tracked"):
            assert f"{n}.{k}" in keys
    assert {"fc.weight", "fc.bias"} <= keys and len(keys) == 8 * 7 + 2
    assert m.conv1[0].weight.shape == (1, 1, 5, 5) and m.conv8[0].weight.shape == (32, 64, 3)
    assert not m.training and len(list(m.parameters())) == 34
    # a fresh model (no checkpoint) starts in train mode like the reference (audionet_csine.py:120-124)
    from speakerguard_b200.model.audionet_csine import audionet_csine
    assert audionet_csine(num_class=7, device="cuda:0").training


@pytest.mark.parametrize("B,N", [(6, 16000), (3, 24000)])
def test_train_step_against_fp64_oracle(params, B, N):
    torch.manual_seed(B * 100 + 1)
    x = (torch.rand(B, 1, N) * 2 - 1) * 0.5
    y = torch.randint(0, 251, (B,))
    o = O.audionet_train_step(x[:, 0].double(), y, {k: v.double() for k, v in params.items()})
    model = make_model(params)
    opt = torch.optim.Adam(model.parameters())
    xr = x.cuda().requires_grad_(True)
    out = model(xr)
    loss = torch.nn.CrossEntropyLoss()(out, y.cuda())
    opt.zero_grad()
    loss.backward()
    e = {"logits": rel(out, o["logits"]), "loss": abs(float(loss) - float(o["loss"])), "xgrad": rel(xr.grad[:, 0], o["xgrad"])}
    got = grads_of(model)
    errs = {}
    for k in O.AN_PARAM_KEYS:
        if k.endswith(".bias") and k != "fc.bias":
            assert float(got[k].abs().max()) < 1e-4 * max(1.0, float(got[k.replace(".bias", ".weight")].abs().max())), k
            continue
        errs[k] = rel(got[k], o["grads"][k])
    top = sorted(errs.items(), key=lambda t: -t[1])[:4]
    print(f"AudioNet train step B={B} N={N}: {e} largest param-grad errors {top}")
    assert e["logits"] < 1e-4 and e["loss"] < 1e-5 and e["xgrad"] < 1e-4
    # the BatchNorm2d(1) of the pre-filter reduces B*32*T signed terms to ONE number: the sum cancels to ~1e-4 of its terms, so
    # fp32 inputs bound the relative accuracy near 1e-3 (the reference's own fp32 run differs from fp64 by 7e-4 there)
    loose = {"conv1.bn_gamma", "conv1.bn_beta"}
    for k, v in errs.items():
        assert v < (2e-3 if k in loose else 1e-4), (k, v)
    for i, n in enumerate(O.AN_BN_NAMES):
        bn = model._bn_modules()[i]
        assert rel(bn.running_mean, o["stats"][f"{n}.bn_mean"]) < 1e-5
        assert rel(bn.running_var, o["stats"][f"{n}.bn_var"]) < 1e-5
        assert int(bn.num_batches_tracked) == 1
    opt.step()
    for k in ("conv2.weight", "conv5.weight", "fc.weight"):
        n, key = k.split(".")
        mod = model.fc if n == "fc" else getattr(model, n)[0]
        sel = o["grads"][k].abs() > 1e-6
        assert float((mod.weight.detach().cpu() - o["params"][k].float()).abs()[sel].max()) < 1e-5


def test_train_step_against_reference_golden(params):
    g = np.load(os.path.join(G, "antrain_golden.npz"))
    B, N = int(g["antrain.B"]), int(g["antrain.N"])
    torch.manual_seed(9091)
    x = (torch.rand(B, 1, N) * 2 - 1) * 0.5
    assert abs(float(x.double().abs().sum()) - float(g["antrain.x_cks"])) < 1e-9
    y = torch.from_numpy(g["antrain.y"])
    model = make_model(params)
    opt = torch.optim.Adam(model.parameters())
    xr = x.cuda().requires_grad_(True)
    out = model(xr)
    loss = torch.nn.CrossEntropyLoss()(out, y.cuda())
    opt.zero_grad()
    loss.backward()
    opt.step()
    assert float((out.detach().cpu() - torch.from_numpy(g["antrain.logits"])).abs().max()) < 1e-4
    assert abs(float(loss) - float(g["antrain.loss"])) < 1e-5
    assert rel(xr.grad[:, 0], g["antrain.xgrad"]) < 1e-4
    got = grads_of(model)
    for k in O.AN_PARAM_KEYS:
        if k.endswith(".bias") and k != "fc.bias":
            continue
        assert rel(got[k], g[f"antrain.grad.{k}"]) < 2e-3, k          # the reference's own fp32 reduction order
    for i, n in enumerate(O.AN_BN_NAMES):
        bn = model._bn_modules()[i]
        np.testing.assert_allclose(bn.running_mean.cpu().numpy(), g[f"antrain.stat.{n}.bn_mean"], atol=1e-5, rtol=1e-5)
        np.testing.assert_allclose(bn.running_var.cpu().numpy(), g[f"antrain.stat.{n}.bn_var"], atol=1e-5, rtol=1e-4)
        for key, mod in (("weight", getattr(model, n)[0].weight), ("bn_gamma", getattr(model, n)[1].weight)):
            sel = np.abs(g[f"antrain.grad.{n}.{key}"]) > 1e-5
            assert np.abs(mod.detach().cpu().numpy() - g[f"antrain.new.{n}.{key}"])[sel].max() < 1e-5


def test_eval_after_training_uses_updated_weights(params):
    """eval() after optimiser steps must re-fold BatchNorm with the new weights and running statistics."""
    torch.manual_seed(3)
    x = ((torch.rand(4, 1, 16000) * 2 - 1) * 0.5).cuda()
    y = torch.randint(0, 251, (4,)).cuda()
    model = make_model(params)
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    before = model.eval()(x).detach().clone()
    model.train()
    for _ in range(3):
        opt.zero_grad()
        torch.nn.CrossEntropyLoss()(model(x), y).backward()
        opt.step()
    model.eval()
    after = model(x).detach()
    assert float((after - before).abs().max()) > 1e-3
    p2 = {k: v.detach().cpu() for k, v in __import__("speakerguard_b200.model.audionet_csine", fromlist=["x"]).params_from_state_dict(model.state_dict()).items()}
    ref = O.audionet_forward(x[:, 0].cpu(), p2)
    assert rel(after, ref) < 1e-4


def test_adversarial_training_step_like_adver_train(params):
    """adver_train.py:183-221: PGD inside the loop on the train-mode model (generic autograd path), then the optimisation step."""
    from speakerguard_b200.attack.PGD import PGD
    torch.manual_seed(4)
    x = ((torch.rand(4, 1, 16000) * 2 - 1) * 0.5).cuda()
    y = torch.randint(0, 251, (4,)).cuda()
    model = make_model(params)
    opt = torch.optim.Adam(model.parameters())
    attacker = PGD(model, targeted=False, step_size=0.0004, epsilon=0.002, max_iter=2, batch_size=4, loss="Entropy", verbose=0)
    nb0 = int(model.conv2[1].num_batches_tracked)
    adv, success = attacker.attack(x, y)
    assert float((adv - x).abs().max()) <= 0.002 + 1e-6 and len(success) == 4
    assert int(model.conv2[1].num_batches_tracked) == nb0 + 3            # train-mode forwards update the running statistics
    out = model(adv)
    loss = torch.nn.CrossEntropyLoss()(out, y)
    opt.zero_grad()
    loss.backward()
    opt.step()
    assert torch.isfinite(loss) and all(torch.isfinite(p).all() for p in model.parameters())


def test_adam_kernel_matches_torch():
    from speakerguard_b200.engine import Engine
    eng = Engine("cuda:0")
    g = torch.Generator().manual_seed(0)
    p0 = torch.randn(10007, generator=g)
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-3, weight_decay=0.01)
    p, m, v = p0.clone().cuda(), torch.zeros(10007).cuda(), torch.zeros(10007).cuda()
    for step in range(1, 6):
        gr = torch.randn(10007, generator=g)
        ref.grad = gr.clone()
        opt.step()
        eng.adam_step(p, gr.cuda(), m, v, step, lr=1e-3, weight_decay=0.01)
    assert float((p.cpu() - ref.detach()).abs().max()) < 1e-6
