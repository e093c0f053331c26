"""Shard equivalence: the utterance axis split contiguously over G handles reproduces the unsharded
attack bit for bit, because the in-kernel philox dither is keyed on the GLOBAL utterance index (SG_OPT_UTT_OFFSET).
Two handles on two streams of one GPU stand in for two GPUs; with >= 2 GPUs the second handle lives on cuda:1 (which
also exercises the per-device kernel attributes / __constant__ uploads)."""
import pytest
import torch

from oracle import sg_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_two_way_split_equals_unsharded(prec):
    from speakerguard_b200 import _lib
    from speakerguard_b200.engine import Engine, make_loss_params
    p = O.make_xv_params(seed=0)
    B, N = 6, 32000
    torch.manual_seed(42)
    x = ((torch.rand(B, 1, N) * 2 - 1) * 0.5)[:, 0].contiguous()
    y = torch.randint(0, 10, (B,))
    lp = make_loss_params("Entropy")
    kw = dict(max_iter=3, epsilon=0.002, step_size=0.0004, lp=lp, dither_mode=_lib.DITHER_PHILOX, seed=77, grad_sign=1.0)
    full = Engine("cuda:0", precision=prec)
    full.load_xv(p)
    xa = x.cuda().clone()
    dec, scores, _ = full.pgd_run(xa, x.cuda(), y.cuda(), **kw)
    torch.cuda.synchronize()
    dev1 = "cuda:1" if torch.cuda.device_count() > 1 else "cuda:0"
    parts = []
    for (lo, hi), dev in (((0, 4), "cuda:0"), ((4, 6), dev1)):
        e = Engine(dev, precision=prec)
        e.load_xv(p)
        with torch.cuda.device(dev), torch.cuda.stream(torch.cuda.Stream(device=dev)):
            xs = x[lo:hi].to(dev)
            xo = xs.clone()
            d, s, _ = e.pgd_run(xo, xs, y[lo:hi].to(dev), utt_offset=lo, **kw)
            torch.cuda.synchronize(dev)
        parts.append((xo.cpu(), d.cpu(), s.cpu()))
    assert torch.equal(torch.cat([a for a, _, _ in parts]), xa.cpu())
    assert torch.equal(torch.cat([d for _, d, _ in parts]), dec.cpu())
    assert torch.equal(torch.cat([s for _, _, s in parts]), scores.cpu())
    # and the offset matters: without it the second shard draws utterance 0's noise
    e = Engine("cuda:0", precision=prec)
    e.load_xv(p)
    xo = x[4:6].cuda().clone()
    e.pgd_run(xo, x[4:6].cuda(), y[4:6].cuda(), utt_offset=0, **kw)
    assert not torch.equal(xo.cpu(), xa.cpu()[4:6])


def test_sharded_attack_helper_through_the_public_classes():
    """dist.sharded_attack with world = 2 emulated in one process == PGD.attack on the whole batch."""
    import tempfile
    from speakerguard_b200 import dist as sgd
    from speakerguard_b200.attack.PGD import PGD
    from speakerguard_b200.model.xv_plda import xv_plda
    from speakerguard_b200.synthetic import make_xv_params, state_dict_of, synthetic_batch, write_xv_model_files
    p = make_xv_params(0)

    def model():
        with tempfile.TemporaryDirectory() as tmp:
            f = write_xv_model_files(p, tmp)
            return xv_plda(state_dict_of(p), f["plda.txt"], f["mean.vec"], f["transform.txt"], model_file=f["speaker_model"],
                           device="cuda:0", precision="bf16", dither="philox", seed=5)

    x, y = synthetic_batch(5, 32000, 10)
    x, y = x.cuda(), y.cuda()
    mk = lambda m: PGD(m, epsilon=0.002, step_size=0.0004, max_iter=2, batch_size=8, verbose=0)
    ref_adv, ref_suc = mk(model()).attack(x, y)
    got_adv, got_suc = [], []
    for r in range(2):
        adv, suc, (lo, hi) = sgd.sharded_attack(mk(model()), x, y, r, 2)      # a fresh model per rank, like a fresh process
        got_adv.append(adv)
        got_suc += suc
    assert torch.equal(torch.cat(got_adv), ref_adv)
    assert got_suc == ref_suc


@pytest.mark.parametrize("prec,iters,eot", [("fp32", 3, 1), ("bf16", 4, 1), ("bf16", 5, 1), ("bf16", 2, 3)])
def test_graph_replay_is_bit_identical_to_launch_by_launch(prec, iters, eot):
    """SG_OPT_CUDA_GRAPH: the captured-iteration replay of sg_pgd_run (seed / pass counter read from the device control
    block) gives exactly the iterates, scores and decisions of the launch-by-launch loop - odd and even iteration counts
    (ping-pong parity), EOT > 1 (in-place step), and a second attack with another seed on the same captured graphs."""
    from speakerguard_b200 import _lib
    from speakerguard_b200.engine import Engine, make_loss_params
    p = O.make_xv_params(seed=0)
    B, N = 4, 32000
    torch.manual_seed(9)
    x = ((torch.rand(B, 1, N) * 2 - 1) * 0.5)[:, 0].contiguous().cuda()
    y = torch.randint(0, 10, (B,)).cuda()
    lp = make_loss_params("Entropy")
    eng = Engine("cuda:0", precision=prec)
    eng.load_xv(p)
    ws = eng.pgd_ws(B, N)
    out = {}
    for graph in (0, 1):
        eng.set_option(_lib.OPT_CUDA_GRAPH, graph)
        for seed in (21, 22):
            xa = x.clone()
            n0 = eng.launch_count()
            dec, sc, _ = eng.pgd_run(xa, x, y, max_iter=iters, epsilon=0.002, step_size=0.0004, lp=lp, eot_size=eot,
                                     dither_mode=_lib.DITHER_PHILOX, seed=seed, ws=ws, grad_sign=1.0)
            torch.cuda.synchronize()
            out[(graph, seed)] = (xa.cpu(), dec.cpu(), sc.cpu(), eng.launch_count() - n0)
    for seed in (21, 22):
        a, b = out[(0, seed)], out[(1, seed)]
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
        assert abs(a[3] - b[3]) <= 8 + iters, (a[3], b[3])          # same kernels (+ control-block ticks / copies)
    assert not torch.equal(out[(1, 21)][0], out[(1, 22)][0])


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_sharded_feco_eot_loop_equals_unsharded(prec):
    """FeCo + EOT inside the fused loop with EOT copies as batch rows: the philox dither and the k-means streams of a row are
    keyed by (EOT copy, GLOBAL utterance index), not by the row's position in the pass, so a contiguous split of the batch
    over two handles reproduces the unsharded attack bit for bit although the tiled rows sit at different positions."""
    from speakerguard_b200 import _lib
    from speakerguard_b200.engine import Engine, make_loss_params
    p = O.make_xv_params(seed=0)
    B, N = 6, 32000
    torch.manual_seed(43)
    x = ((torch.rand(B, 1, N) * 2 - 1) * 0.5)[:, 0].contiguous()
    y = torch.randint(0, 10, (B,))
    kw = dict(max_iter=3, epsilon=0.002, step_size=0.0004, lp=make_loss_params("Entropy"), dither_mode=_lib.DITHER_PHILOX, seed=78,
              grad_sign=1.0, eot_size=4, eot_batch=2, feco_ratio=0.5)
    full = Engine("cuda:0", precision=prec)
    full.load_xv(p)
    xa = x.cuda().clone()
    dec, scores, _ = full.pgd_run(xa, x.cuda(), y.cuda(), **kw)
    torch.cuda.synchronize()
    parts = []
    for lo, hi in ((0, 4), (4, 6)):
        e = Engine("cuda:0", precision=prec)
        e.load_xv(p)
        xs = x[lo:hi].cuda()
        xo = xs.clone()
        d, s, _ = e.pgd_run(xo, xs, y[lo:hi].cuda(), utt_offset=lo, **kw)
        torch.cuda.synchronize()
        parts.append((xo.cpu(), d.cpu(), s.cpu()))
    assert torch.equal(torch.cat([a for a, _, _ in parts]), xa.cpu())
    assert torch.equal(torch.cat([d for _, d, _ in parts]), dec.cpu())
    assert torch.equal(torch.cat([s for _, _, s in parts]), scores.cpu())
