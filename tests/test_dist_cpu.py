"""Multi-GPU host logic on CPU: world_size-2 gloo process group (no GPU needed).
The attack path has no data-path collective: ranks shard utterances contiguously and only the
final metric scalars are all-reduced."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, q):
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank),
                       "WORLD_SIZE": str(world), "LOCAL_RANK": str(rank)})
    from speakerguard_b200 import dist as sgd
    r, w, _ = sgd.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    lo, hi = sgd.shard_bounds(n_total, rank, world)
    g = torch.Generator().manual_seed(0)
    x = torch.rand(n_total, 1, 64, generator=g) - 0.5
    adv = x + 0.002 * torch.sign(torch.randn(n_total, 1, 64, generator=g))
    success = [(i % 3) != 0 for i in range(n_total)]
    local = sgd.attack_metrics(x[lo:hi], adv[lo:hi], success[lo:hi])
    red = sgd.reduce_metrics(local)
    mx = sgd.max_over_ranks(float(rank + 1), "cpu")
    sgd.barrier()
    q.put((rank, lo, hi, red, mx))
    dist.destroy_process_group()


def test_shard_bounds_partition():
    from speakerguard_b200.dist import shard_bounds
    for n in (1, 7, 8, 1024, 1025):
        for world in (1, 2, 3, 8):
            cuts = [shard_bounds(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_metric_reduction():
    from speakerguard_b200 import dist as sgd
    world, n_total = 2, 9
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    out.sort()
    assert out[0][1:3] == (0, 5) and out[1][1:3] == (5, 9)
    # single-process ground truth over the whole batch
    g = torch.Generator().manual_seed(0)
    x = torch.rand(n_total, 1, 64, generator=g) - 0.5
    adv = x + 0.002 * torch.sign(torch.randn(n_total, 1, 64, generator=g))
    success = [(i % 3) != 0 for i in range(n_total)]
    ref = sgd.reduce_metrics(sgd.attack_metrics(x, adv, success))
    for _, _, _, red, mx in out:
        assert mx == 2.0
        for k in ref:
            assert abs(red[k] - ref[k]) < 1e-9, (k, red[k], ref[k])
    assert ref["n"] == 9 and abs(ref["success_rate"] - 6 / 9) < 1e-12 and abs(ref["linf"] - 0.002) < 1e-6


class _FakeAttacker:
    """Stands in for PGD on CPU: the 'dither' of utterance g is a function of its GLOBAL index, like the philox key."""
    utt_offset = 0

    def attack(self, x, y):
        idx = torch.arange(x.shape[0]) + self.utt_offset
        noise = torch.sin(idx.double() * 12.9898).view(-1, 1, 1).float()
        return x + 0.001 * noise, [bool((int(i) + int(t)) % 2) for i, t in zip(idx, y)]


def _shard_worker(rank, world, port, q):
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank),
                       "WORLD_SIZE": str(world), "LOCAL_RANK": str(rank)})
    from speakerguard_b200 import dist as sgd
    sgd.init_from_env(backend="gloo")
    g = torch.Generator().manual_seed(3)
    x, y = torch.rand(11, 1, 32, generator=g) - 0.5, torch.randint(0, 10, (11,), generator=g)
    att = _FakeAttacker()
    adv, success, (lo, hi) = sgd.sharded_attack(att, x, y, rank, world)
    assert att.utt_offset == 0                                   # restored
    red = sgd.reduce_metrics(sgd.attack_metrics(x[lo:hi], adv, success))
    q.put((rank, lo, hi, adv, success, red))
    dist.destroy_process_group()


def test_sharded_attack_reproduces_the_unsharded_one():
    """dist.sharded_attack: rank r attacks x[lo:hi] with the attacker's utt_offset = lo; concatenating the shards gives the
    single-process result bit for bit (the GPU form of this check is tests/test_gpu_shard.py)."""
    from speakerguard_b200 import dist as sgd
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_shard_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(3)
    x, y = torch.rand(11, 1, 32, generator=g) - 0.5, torch.randint(0, 10, (11,), generator=g)
    ref_adv, ref_suc = _FakeAttacker().attack(x, y)
    assert torch.equal(torch.cat([o[3] for o in out]), ref_adv)
    assert out[0][4] + out[1][4] == ref_suc
    ref = sgd.reduce_metrics(sgd.attack_metrics(x, ref_adv, ref_suc))
    for o in out:
        for k in ref:
            assert abs(o[5][k] - ref[k]) < 1e-9
