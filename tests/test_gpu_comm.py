"""sg_comm_init / sg_allreduce_metrics: libsgb200's own NCCL communicator for the one collective of the path,
the end-of-attack metric reduction.  Needs >= 2 GPUs (NCCL refuses two ranks on one device); single-GPU boxes check that
the entry points fail loudly instead."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank), "WORLD_SIZE": str(world),
                       "LOCAL_RANK": str(rank)})
    import torch.distributed as tdist
    from speakerguard_b200 import dist as sgd
    from speakerguard_b200.engine import Engine
    r, w, local = sgd.init_from_env()
    torch.cuda.set_device(local)
    eng = Engine(f"cuda:{local}")
    sgd.engine_comm_init(eng)
    v = torch.tensor([1.0, rank + 1.0, 10.0 * (rank + 1), 0.5, 0.002], dtype=torch.float64, device=f"cuda:{local}")
    red = sgd.reduce_metrics(v, engine=eng)
    ref = sgd.reduce_metrics(v)                                   # torch.distributed's all-reduce of the same vector
    q.put((rank, red, ref))
    tdist.barrier()
    tdist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_nccl_metric_allreduce_matches_torch_distributed():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for _, red, ref in out:
        assert red == ref
        assert red["n"] == 2.0 and abs(red["success_rate"] - 3.0 / 2.0) < 1e-12


def test_allreduce_without_a_communicator_fails_loudly():
    from speakerguard_b200 import _lib
    from speakerguard_b200.engine import Engine
    eng = Engine("cuda:0")
    v = torch.zeros(5, dtype=torch.float64, device="cuda:0")
    with pytest.raises(_lib.SgError, match="sg_comm_init"):
        eng.allreduce_metrics(v)
    assert eng.lib.sg_comm_nccl_version() >= 20000
