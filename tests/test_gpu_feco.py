"""FeCo (k-means feature compression) on the GPU: conditional parity against the reference's own
mean-by-cluster code (fixture built from defense/feature_level.py with injected ids), Lloyd
fixed-point / inertia checks for the ids (libKMCUDA is absent and randomised: parity unpinned),
and EOT-PGD through a FeCo-defended xv_plda (BASELINE config 4, small)."""
import os

import numpy as np
import pytest
import torch

from oracle import sg_oracle as O

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def test_means_and_gradient_given_ids_match_reference():
    from speakerguard_b200.defense.feature_level import FeCo
    g = np.load(os.path.join(G, "feco_golden.npz"))
    feat = torch.tensor(g["feco.feat"]).cuda()
    ids = torch.tensor(g["feco.ids"]).int().cuda()
    k = int(g["feco.k"])
    f2 = torch.stack([feat, feat]).requires_grad_(True)            # the fixture fed the same utterance twice
    out = FeCo(f2, "kmeans", 0.5, "L2", ids=torch.stack([ids, ids]))
    assert out.shape == (2, k, 30)
    np.testing.assert_allclose(out[0].detach().cpu().numpy(), g["feco.out"], atol=1e-6, rtol=1e-6)
    (out * torch.tensor(g["feco.w"]).cuda()).sum().backward()
    np.testing.assert_allclose(f2.grad.sum(0).cpu().numpy(), g["feco.grad"], atol=1e-6, rtol=1e-5)
    # batch of one: the empty cluster is dropped like the reference does (force=False)
    out1 = FeCo(feat.unsqueeze(0), "kmeans", 0.5, "L2", ids=ids.unsqueeze(0))
    ref1 = O.feco_means(torch.tensor(g["feco.feat"]), g["feco.ids"], k, force=False)
    assert out1.shape[1] == ref1.shape[0] < k
    np.testing.assert_allclose(out1[0].cpu().numpy(), ref1.numpy(), atol=1e-6, rtol=1e-6)


@pytest.mark.parametrize("n,ratio", [(300, 0.5), (300, 0.2), (111, 0.5)])
def test_lloyd_ids_are_a_good_clustering(n, ratio):
    from sklearn.cluster import KMeans
    from speakerguard_b200.defense.feature_level import kmeans_ids
    g = torch.Generator().manual_seed(n)
    B, k = 6, int(n * ratio)
    # clustered data: frames drawn around k/3 prototypes + noise (MFCC-like scale)
    proto = torch.randn(B, max(k // 3, 2), 30, generator=g) * 6
    pick = torch.randint(0, proto.shape[1], (B, n), generator=g)
    feat = torch.gather(proto, 1, pick.unsqueeze(2).expand(B, n, 30)) + torch.randn(B, n, 30, generator=g)
    ids = kmeans_ids(feat.cuda(), k, seed=7).cpu().numpy()
    ids2 = kmeans_ids(feat.cuda(), k, seed=7).cpu().numpy()
    assert np.array_equal(ids, ids2)                              # deterministic given the seed
    assert ids.min() >= 0 and ids.max() < k
    for b in range(B):
        X = feat[b].numpy()
        cent = np.stack([X[ids[b] == c].mean(0) if (ids[b] == c).any() else np.full(30, 1e9) for c in range(k)])
        d = ((X[:, None, :] - cent[None]) ** 2).sum(2)
        nearest = d.argmin(1)
        frac_ok = float((d[np.arange(n), ids[b]] <= d[np.arange(n), nearest] * (1 + 1e-5) + 1e-6).mean())
        assert frac_ok >= 0.97, frac_ok                           # Lloyd fixed point up to the 1 % stop tolerance
        ours = O.lloyd_inertia(X, ids[b], k)
        sk = KMeans(n_clusters=k, n_init=1, init="k-means++", random_state=0).fit(X).inertia_
        assert ours <= 1.25 * sk + 1e-6, (ours, sk)


def test_eot_pgd_against_feco_defended_xv_plda(tmp_path):
    """BASELINE config 4 in miniature: PGD with EOT over the randomised FeCo defense at the raw-feature level."""
    from speakerguard_b200.attack.PGD import PGD
    from speakerguard_b200.defense.feature_level import FeCo
    from speakerguard_b200.model.defended_model import defended_model
    from speakerguard_b200.model.xv_plda import xv_plda
    from speakerguard_b200.synthetic import make_xv_params, state_dict_of, write_xv_model_files
    p = make_xv_params(0)
    f = write_xv_model_files(p, str(tmp_path))
    base = xv_plda(state_dict_of(p), f["plda.txt"], f["mean.vec"], f["transform.txt"], model_file=f["speaker_model"],
                   device="cuda:0", dither="philox")
    dm = defended_model(base, defense=[[1, lambda feat: FeCo(feat, "kmeans", 0.5, "L2")]], order="sequential")
    torch.manual_seed(3)
    x = ((torch.rand(2, 1, 32000) * 2 - 1) * 0.5).cuda()
    y = torch.tensor([1, 6]).cuda()
    dec, scores = dm.make_decision(x)                             # 200 frames -> 100 compressed frames -> TDNN
    assert scores.shape == (2, 10)
    att = PGD(dm, epsilon=0.002, step_size=0.0004, max_iter=2, batch_size=2, EOT_size=4, EOT_batch_size=2, verbose=0)
    adv, success = att.attack(x, y)
    assert adv.shape == x.shape and len(success) == 2
    d = (adv - x).abs()
    assert 0 < float(d.max()) <= 0.002 + 1e-7
    # the gradient through FeCo reaches the waveform: a finite-difference check of the means' adjoint
    feat = torch.randn(2, 60, 30, generator=torch.Generator().manual_seed(1)).cuda().requires_grad_(True)
    from speakerguard_b200.defense.feature_level import kmeans_ids
    ids = kmeans_ids(feat, 30, seed=1)
    out = FeCo(feat, "kmeans", 0.5, "L2", ids=ids)
    w = torch.randn(out.shape, generator=torch.Generator().manual_seed(2)).cuda()
    (out * w).sum().backward()
    ref_out = torch.stack([O.feco_means(feat[b].detach().cpu(), ids[b].cpu().numpy(), 30) for b in range(2)])
    assert float((out.detach().cpu() - ref_out).abs().max()) < 1e-5


# ---- FeCo + EOT inside the fused device loop (sg_pgd_params::feco_ratio / eot_batch) --------------------------------------
@pytest.fixture(scope="module")
def eng_fp32():
    from speakerguard_b200.engine import Engine
    e = Engine("cuda:0", precision="fp32")
    e.load_xv(O.make_xv_params(seed=0))
    return e


def test_fused_feco_step_equals_stagewise_composition(eng_fp32):
    """One PGD step through FeCo inside sg_pgd_run == the same kernels called stage by stage (raw MFCC -> k-means with the
    pass's seed -> cluster means -> CMVN over the k means -> TDNN -> loss -> adjoints -> sign step): the iterate must be
    identical, and so must the final decisions (second pass, pass index 1)."""
    from speakerguard_b200 import _lib
    from speakerguard_b200.engine import make_loss_params
    eng = eng_fp32
    torch.manual_seed(5)
    B, N, seed = 3, 32000, 0x1234567890ABCDEF
    x = ((torch.rand(B, N) * 2 - 1) * 0.5).cuda()
    y = torch.tensor([1, 6, 9]).cuda()
    lp = make_loss_params("Entropy")
    xa = x.clone()
    dec, scores, _ = eng.pgd_run(xa, x, y, max_iter=1, epsilon=0.002, step_size=0.0004, lp=lp, dither_mode=_lib.DITHER_OFF,
                                 seed=seed, feco_ratio=0.5, grad_sign=1.0)
    m = eng.num_frames(N)
    k = int(m * 0.5)

    def defended_forward(xin, pass_):
        raw = eng.mfcc_fwd(xin, _lib.DITHER_OFF, None, ld=32)
        ids = eng.feco_kmeans(raw[:, :, :30].contiguous(), k, seed=seed, pass_=pass_)       # the loop's clustering of that pass
        means, counts = eng.feco_means_fwd(raw[:, :, :30].contiguous(), ids, k, True)
        feat = eng.cmvn(means, ld_out=32)
        emb, ws = eng.embed_fwd(feat)
        sc, d = eng.score_fwd(emb)
        return raw, ids, counts, feat, emb, ws, sc, d

    raw, ids, counts, feat, emb, ws, sc, _ = defended_forward(x, 0)
    _, ds = eng.loss(sc, y, lp)
    dfeat = eng.embed_bwd(eng.score_bwd(emb, ds), ws, B, k)
    dmeans = eng.cmvn(dfeat, ld_out=30, backward=True)
    draw30 = eng.feco_means_bwd(dmeans, ids, counts, m, True)
    draw = torch.zeros(B, m, 32, device="cuda")
    draw[:, :, :30] = draw30
    g = eng.mfcc_bwd(x, draw, _lib.DITHER_OFF)
    x1 = x.clone()
    eng.step_linf(x1, x, g, 0.0004, 1.0, 0.002)
    same = float((x1 == xa).float().mean())
    print(f"fused FeCo step vs stage-wise: {same:.6f} of the samples identical")
    assert same > 0.9999
    *_, sc1, d1 = defended_forward(xa, 1)
    assert torch.equal(d1, dec)
    assert float((sc1 - scores).abs().max()) < 1e-4 * float(scores.abs().max())


@pytest.mark.parametrize("feco", [0.0, 0.5])
def test_eot_copies_as_batch_rows(eng_fp32, feco):
    """eot_batch: E copies of the batch as B * eot_batch rows per pass.  With the dither off (and, for FeCo, E = eot_batch so
    that both runs draw the k-means seeds of the same passes) every way of batching the copies averages the same gradients:
    the iterates agree wherever the gradient is not within rounding of zero, and the graph replay equals launch-by-launch."""
    from speakerguard_b200 import _lib
    from speakerguard_b200.engine import make_loss_params
    eng = eng_fp32
    torch.manual_seed(6)
    B, N = 3, 24000
    x = ((torch.rand(B, N) * 2 - 1) * 0.5).cuda()
    y = torch.tensor([2, 4, 8]).cuda()
    lp = make_loss_params("Entropy")
    out = {}
    for name, kw, graph in [("e1", dict(eot_size=1, eot_batch=1), 0), ("e4b4", dict(eot_size=4, eot_batch=4), 0),
                            ("e4b4_graph", dict(eot_size=4, eot_batch=4), 1), ("e4b2", dict(eot_size=4, eot_batch=2), 0),
                            ("e4b1", dict(eot_size=4, eot_batch=1), 0)]:
        eng.set_option(_lib.OPT_CUDA_GRAPH, graph)
        xa = x.clone()
        dec, sc, _ = eng.pgd_run(xa, x, y, max_iter=3, epsilon=0.002, step_size=0.0004, lp=lp, dither_mode=_lib.DITHER_OFF,
                                 seed=77, grad_sign=1.0, feco_ratio=feco, **kw)
        out[name] = (xa.cpu(), dec.cpu(), sc.cpu())
    eng.set_option(_lib.OPT_CUDA_GRAPH, 1)
    assert torch.equal(out["e4b4"][0], out["e4b4_graph"][0]) and torch.equal(out["e4b4"][1], out["e4b4_graph"][1])
    if feco == 0.0:
        # identical copies: the average of E equal gradients is the gradient (up to the summation order of the copies)
        for name in ("e4b4", "e4b2", "e4b1"):
            same = float((out[name][0] == out["e1"][0]).float().mean())
            print(f"{name} vs e1: {same:.6f} identical")
            assert same > 0.999
            assert torch.equal(out[name][1], out["e1"][1])
    else:
        # with FeCo the copies differ (one clustering per row and pass): check the perturbation is a valid PGD iterate
        for name in ("e4b4", "e4b2", "e4b1", "e1"):
            d = (out[name][0] - x.cpu()).abs()
            assert 0 < float(d.max()) <= 0.002 + 1e-7 and torch.isfinite(out[name][2]).all()


def test_fused_feco_through_the_attack_classes(tmp_path):
    """PGD(defended_model(xv, [[1, FeCoDefense]])) takes the fused loop (FeCo + EOT on the device).  With the dither off the
    clustering is the only randomness; a generic-path defense that is handed the fused loop's per-pass k-means seeds must
    then reproduce the fused attack (same kernels, EOT copies as batch rows in both) up to gradient-rounding sign flips."""
    import itertools
    from speakerguard_b200.attack.FGSM import fused_target
    from speakerguard_b200.attack.PGD import PGD
    from speakerguard_b200.defense.feature_level import FeCo, FeCoDefense
    from speakerguard_b200.model.defended_model import defended_model
    from speakerguard_b200.model.xv_plda import xv_plda
    from speakerguard_b200.synthetic import make_xv_params, state_dict_of, write_xv_model_files
    p = make_xv_params(0)
    f = write_xv_model_files(p, str(tmp_path))
    base = xv_plda(state_dict_of(p), f["plda.txt"], f["mean.vec"], f["transform.txt"], model_file=f["speaker_model"],
                   device="cuda:0", dither="off", precision="fp32")
    dm_fused = defended_model(base, defense=[[1, FeCoDefense("kmeans", 0.5, "L2")]], order="sequential")
    assert fused_target(dm_fused)[0] is base and fused_target(dm_fused)[1] is not None
    assert fused_target(defended_model(base, defense=[[1, lambda feat: FeCo(feat, "kmeans", 0.5, "L2")]])) == (None, None)
    torch.manual_seed(3)
    x = ((torch.rand(4, 1, 32000) * 2 - 1) * 0.5).cuda()
    with torch.no_grad():
        y = base.make_decision(x)[0]
    iters, E = 2, 4
    # the seed the fused loop will be given (xv_plda.fused_dither) and the pass order of its k-means launches:
    # iteration it -> pass it (all E copies in one pass of 4 * E rows), final evaluation -> pass `iters`
    seed = (base.seed + 0x9E3779B97F4A7C15 * base._pass) & 0xFFFFFFFFFFFFFFFF
    adv_f, suc_f = PGD(dm_fused, epsilon=0.002, step_size=0.0004, max_iter=iters, batch_size=4, EOT_size=E, EOT_batch_size=E,
                       verbose=0).attack(x, y)
    counter = itertools.count(0)
    # the generic EOT wrapper tiles the batch like the fused loop does (copies outermost): rows = 4 utterances repeated in
    # the E-copy passes, the plain batch in the final evaluation
    dm_seeded = defended_model(base, order="sequential",
                               defense=[[1, lambda feat: FeCo(feat, "kmeans", 0.5, "L2", seed=seed, pass_=next(counter),
                                                              copy_rows=4 if feat.shape[0] > 4 else 0)]])
    adv_g, suc_g = PGD(dm_seeded, epsilon=0.002, step_size=0.0004, max_iter=iters, batch_size=4, EOT_size=E, EOT_batch_size=E,
                       verbose=0).attack(x, y)
    assert next(counter) == iters + 1                                # one defended forward per iteration + the final evaluation
    same = float((adv_f == adv_g).float().mean())
    print(f"fused FeCo + EOT attack vs generic path with the same k-means seeds: {same:.6f} of the samples identical")
    assert same > 0.999 and suc_f == suc_g
    assert 0 < float((adv_f - x).abs().max()) <= 0.002 + 1e-7
