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
