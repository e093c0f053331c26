"""Multi-GPU plumbing: one process per GPU, utterances sharded contiguously, no data-path
collective (every utterance is independent through the whole attack iteration).
NCCL (via torch.distributed) is used only to reduce a handful of metric scalars at the end of an
attack: success count, sum of SNR / L2 / Linf and the utterance count (reference
metric/metric.py:10-42 semantics for SNR / Lp)."""
from __future__ import annotations

import os
from typing import Dict, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: str = None) -> Tuple[int, int, int]:
    """-> (rank, world, local_rank); initialises the default process group when WORLD_SIZE > 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split of n utterances: rank r owns [lo, hi); remainders go to the first ranks."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def sharded_attack(attacker, x: torch.Tensor, y: torch.Tensor, rank: int, world: int, device=None):
    """Strong-scaling form of ``attacker.attack(x, y)``: rank r attacks x[lo:hi] of the *global* batch on its
    own GPU, no data-path collective.  The attacker's ``utt_offset`` is set to ``lo`` so the in-kernel philox dither is keyed
    on the global utterance index: concatenating the ranks' results reproduces the single-GPU attack bit for bit.
    -> (adv [hi-lo,1,N], success list, (lo, hi))"""
    lo, hi = shard_bounds(x.shape[0], rank, world)
    xs, ys = x[lo:hi], y[lo:hi]
    if device is not None:
        xs, ys = xs.to(device, non_blocking=True), ys.to(device, non_blocking=True)
    prev = getattr(attacker, "utt_offset", 0)
    attacker.utt_offset = lo
    try:
        adv, success = attacker.attack(xs, ys)
    finally:
        attacker.utt_offset = prev
    return adv, success, (lo, hi)


def attack_metrics(x: torch.Tensor, adv: torch.Tensor, success) -> torch.Tensor:
    """Local sums [n, n_success, sum SNR(dB), sum L2, sum Linf] as one fp64 vector on x's device."""
    x2, a2 = x.flatten(1).double(), adv.detach().flatten(1).double()
    delta = a2 - x2
    noise = delta.pow(2).sum(1)
    snr = 10.0 * torch.log10(x2.pow(2).sum(1) / noise.clamp_min(1e-300))
    snr = torch.where(noise > 0, snr, torch.zeros_like(snr))
    suc = torch.as_tensor(success, dtype=torch.float64, device=x.device)
    return torch.stack([torch.tensor(float(x.shape[0]), dtype=torch.float64, device=x.device), suc.sum(), snr.sum(),
                        noise.sqrt().sum(), delta.abs().max(1)[0].sum()])


def engine_comm_init(engine) -> None:
    """Give ``engine`` its own NCCL communicator (sg_comm_init) for the metric all-reduce: rank 0's unique id travels over
    the already-initialised torch.distributed group (host side, once per process)."""
    if getattr(engine, "comm_world", 0) or not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return

    def exchange(ident):
        box = [ident]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    engine.comm_init(dist.get_rank(), dist.get_world_size(), exchange)


def reduce_metrics(local: torch.Tensor, engine=None) -> Dict[str, float]:
    """All-reduce (sum) of the metric vector over the job; a no-op without a process group.  With an ``engine`` that has a
    communicator (engine_comm_init) the reduction is libsgb200's own sg_allreduce_metrics; otherwise torch.distributed's."""
    v = local.clone()
    if engine is not None and getattr(engine, "comm_world", 0) > 1 and v.is_cuda:
        engine.allreduce_metrics(v)
    elif dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(v, op=dist.ReduceOp.SUM)
    n = max(float(v[0]), 1.0)
    return {"n": float(v[0]), "success_rate": float(v[1]) / n, "snr_db": float(v[2]) / n, "l2": float(v[3]) / n,
            "linf": float(v[4]) / n}


def max_over_ranks(value: float, device) -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def barrier() -> None:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
