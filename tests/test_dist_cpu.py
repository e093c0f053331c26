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
