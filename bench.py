#!/usr/bin/env python
"""Headline benchmark: PGD utterance-iterations/s against xv_plda (BASELINE.json configs[1]).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision ...]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one full PGD-100 attack (100 gradient passes + the final evaluation pass,
attack/PGD.py semantics) on one batch of 1024 synthetic 3 s utterances per GPU (weak scaling:
utterances shard independently, no data-path collective).  Prints ONE JSON line on rank 0.

  value        utterance-iterations/s, whole job, inputs resident in HBM, device-timed (CUDA
               events, max over ranks), through the C-ABI (sg_pgd_run).
  e2e          same metric through the public drop-in API ``PGD(model).attack(x, y)`` with HOST
               buffers: pinned H2D of x/y and D2H of the adversarial batch + success inside the
               timed region.
  roofline     the TDNN contraction kernel (forward + dgrad launches of one profiled step):
               algorithmic FLOPs / CUDA-event time vs the measured tensor peak.
  cpu_baseline the oracle port of the reference path on the host cores, bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOAD = "PGD-100 Linf eps=0.002 step=0.0004 CE-untargeted vs xv_plda CSI-E (random-init TDNN+PLDA, L=200, S=10), " \
           "synthetic 3 s 16 kHz utterances, batch 1024 per GPU"
PRECISION_NOTE = {
    "fp32": "FFMA everywhere: parity mode, 1e-4 relative vs the reference (tests/test_gpu_xv.py)",
    "tf32": "TDNN contractions on tcgen05 kind::tf32, fp32 storage (the reference's own GPU default for cuDNN convs); "
            "vs fp32 mode: embeddings/scores < 1e-3, 98.6 % input-gradient signs equal (tests/test_gpu_tc.py)",
    "bf16": "TDNN activations/gradients/weights in bf16, fp32 accumulate (tcgen05 kind::f16), everything else fp32; "
            "vs fp32 mode: embeddings/scores < 1e-3, 97.3 % input-gradient signs equal, cosine 0.996, identical attack "
            "success rate (tests/test_gpu_tc.py); --precision tf32 / fp32 select the higher-precision modes",
}
TDNN = [(30, 512, 5, 1), (512, 512, 5, 2), (512, 512, 7, 3), (512, 512, 1, 1), (512, 1500, 1, 1)]


def tdnn_valid_frames(m: int):
    t, out = m, []
    for _, _, k, d in TDNN:
        t -= (k - 1) * d
        out.append(t)
    return out


def peaks_hbm_gbs() -> float:
    return measured_peaks()[0].get("hbm_gbs", 6548.5)


def tdnn_flops_per_utt(m: int) -> float:
    """Algorithmic FLOPs of the TDNN forward + dgrad for one utterance-iteration:
    2 * sum_l 2*C_in*k*C_out*T_l (valid frames only, no wgrad, unpadded channels)."""
    t, tot = m, 0.0
    for ci, co, k, d in TDNN:
        t -= (k - 1) * d
        tot += 2.0 * ci * k * co * t
    return 2.0 * tot


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "power_w_max": max(pw) if pw else None, "samples": len(self.rows)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


def cpu_reference_throughput(p, N, budget_s=12.0, batch=8, iters=2):
    """utt-iter/s of the oracle port of the reference path (PGD, CE, B=8) on the host cores."""
    from oracle import sg_oracle as O
    from speakerguard_b200.synthetic import synthetic_batch
    torch.set_num_threads(os.cpu_count() or 1)
    x, y = synthetic_batch(batch, N, p["enroll"].shape[0])
    m = O.num_frames(N)
    done, t0 = 0, time.perf_counter()
    while True:
        d = torch.randn(iters + 1, batch, m, 400)
        O.pgd_attack(x[:, 0], y, p, epsilon=0.002, step_size=0.0004, max_iter=iters, dither=d)
        done += iters
        el = time.perf_counter() - t0
        if el >= budget_s:
            break
    return batch * done / el, f"oracle PGD-{done} (incl. {done // iters} evaluation passes), B={batch}, {N / 16000:g} s, " \
                              f"{el:.1f} s of CPU work", torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from speakerguard_b200.synthetic import make_xv_params
    p = make_xv_params(0)
    N = int(args.seconds * 16000)
    for _ in range(args.warmup):
        cpu_reference_throughput(p, N, budget_s=0.0, iters=1)
    t0 = time.perf_counter()
    vals = []
    for _ in range(args.steps):
        v, sample, cores = cpu_reference_throughput(p, N, budget_s=args.ref_budget / max(args.steps, 1), iters=2)
        vals.append(v)
    el = time.perf_counter() - t0
    value = sum(vals) / len(vals)
    out = {"impl": "reference", "metric": "PGD utterance-iterations/s vs xv_plda", "value": value, "unit": "utt-iter/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * el / args.steps,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": WORKLOAD, "note": "reference arm = CPU path of the reference (oracle port, torch-CPU fp32, "
                      "all host threads) on a bounded sample of the same workload: B=8 per step"},
           "cpu_baseline": {"value": value, "unit": "utt-iter/s", "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": "utt-iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def iv_cpu_throughput(p, N, budget_s=15.0, batch=2):
    """utt-iter/s of the oracle port of the reference's iv_plda path (PGD, SV margin loss) on the host cores."""
    from oracle import sg_oracle as O
    from speakerguard_b200.synthetic import synthetic_batch
    torch.set_num_threads(os.cpu_count() or 1)
    x, _ = synthetic_batch(batch, N, 1)
    y = torch.zeros(batch, dtype=torch.int64)
    pp = dict(p)
    pp["threshold"] = 0.0
    m = O.num_frames(N)
    done, t0 = 0, time.perf_counter()
    while True:
        d = torch.randn(2, batch, m, 400)
        O.pgd_attack(x[:, 0], y, pp, epsilon=0.002, step_size=0.0004, max_iter=1, task="SV", dither=d, system="iv")
        done += 1
        el = time.perf_counter() - t0
        if el >= budget_s:
            break
    return batch * done / el, f"oracle PGD-{done} (incl. {done} evaluation passes), B={batch}, {N / 16000:g} s, C=2048, D=400, " \
                              f"{el:.1f} s of CPU work", torch.get_num_threads()


def run_iv(args):
    """BASELINE configs[4]: PGD vs iv_plda (SV task), 5 s utterances, batch 256 per GPU.  Same JSON contract as the
    headline line; the dominant kernel is the UBM log-likelihood contraction and its adjoint (split-TF32 on tcgen05)."""
    from speakerguard_b200 import dist
    from speakerguard_b200.attack.PGD import PGD
    from speakerguard_b200.model.iv_plda import iv_plda
    from speakerguard_b200.synthetic import make_iv_params, synthetic_batch
    rank, world, local = dist.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: speakerguard_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B, N, iters = args.batch, int(args.seconds * 16000), args.iters
    Cn, F, D, L = 2048, 72, 400, 200
    p = make_iv_params(0, C=Cn, F=F, D=D, L=L, S=1)
    prec = "fp32" if args.precision == "fp32" else "tf32x3"
    model = iv_plda(None, None, None, None, None, threshold=0.0, device=dev, params=p, precision=prec, seed=rank)
    eng = model.engine
    m = eng.num_frames(N)
    x_host, _ = synthetic_batch(B, N, 1, seed=1234 + rank)
    y_host = torch.zeros(B, dtype=torch.int64)
    x_host, y_host = x_host.pin_memory(), y_host.pin_memory()
    x_dev, y_dev = x_host.to(dev), y_host.to(dev)
    attacker = PGD(model, task="SV", epsilon=0.002, step_size=0.0004, max_iter=iters, batch_size=B, verbose=0)
    for _ in range(args.warmup):
        attacker.attack(x_dev, y_dev)
    torch.cuda.synchronize()
    dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eng.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        adv, success = attacker.attack(x_dev, y_dev)
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    launches = eng.launch_count()
    clocks = sampler.stop() if rank == 0 else None
    ms_step = dist.max_over_ranks(e0.elapsed_time(e1), dev) / args.steps
    value = world * B * iters / (ms_step / 1000.0)

    adv_host = torch.empty(B, 1, N).pin_memory()

    def e2e_step():
        xd = x_host.to(dev, non_blocking=True)
        yd = y_host.to(dev, non_blocking=True)
        a, s_ = attacker.attack(xd, yd)
        adv_host.copy_(a, non_blocking=True)
        torch.cuda.synchronize()

    e2e_step()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(max(args.e2e_steps, 1)):
        e2e_step()
    e2e_s = dist.max_over_ranks(time.perf_counter() - t0, dev) / max(args.e2e_steps, 1)

    eng.profile(True)
    attacker.attack(x_dev, y_dev)
    prof = eng.profile_read()
    eng.profile(False)
    tot_prof = sum(v[0] for v in prof.values())
    if world > 1:
        import torch.distributed as tdist
        tdist.barrier()
        tdist.destroy_process_group()
    if rank != 0:
        return
    passes = iters + 1
    kq = F + F * (F + 1) // 2
    flops_pass = 2.0 * B * m * kq * Cn                      # one contraction over the valid frames
    flops = flops_pass * (2 * iters + 1)                    # forward + adjoint per iteration, forward of the evaluation pass
    peaks, peak_src = measured_peaks()
    peak = peaks["bf16_tflops_sustained"] / 2.0
    gemm_ms = prof["iv_gemm"][0]
    achieved = flops / (gemm_ms / 1000.0) / 1e12
    out = {
        "metric": "PGD utterance-iterations/s vs iv_plda (SV)", "value": value, "unit": "utt-iter/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32" if prec == "fp32" else "tf32x3", "data": "synthetic",
        "config": {"workload": "PGD-%d L-inf eps=0.002 vs iv_plda SV (UBM C=2048 full-cov F=72, i-vector D=400, LDA/PLDA L=200), "
                               "synthetic %g s utterances, batch %d per GPU (BASELINE configs[4])" % (iters, args.seconds, B),
                   "batch_per_gpu": B, "samples": N, "frames": m, "pgd_iters": iters, "passes_per_step": passes,
                   "dither": "philox", "cache": "inputs larger than L2 (per-pass operands > 8 GB)", "precision": prec},
        "clocks": clocks, "gpu_launches": launches,
        "e2e": {"value": world * B * iters / e2e_s, "unit": "utt-iter/s", "h2d_bytes_per_step": B * N * 4 + B * 8,
                "d2h_bytes_per_step": B * N * 4, "ms_per_step": e2e_s * 1000.0,
                "api": "speakerguard_b200.attack.PGD(iv_plda).attack(x, y) with pinned host buffers"},
        "roofline": {"bound": "tensor", "kernel": "iv_gemm category: UBM log-likelihood contraction + adjoint (and the small statistics / "
                     "extractor contractions timed with it)", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                     "frac": achieved / peak, "traffic": None,
                     "peak_source": f"tf32 = half of bf16 sustained, {peak_src}",
                     "note": "algorithmic flops count every product once; tf32x3 issues 3 tensor-core products per algorithmic product "
                             "(operands split hi+lo for fp32-level accuracy), so the pipe utilisation is ~3x frac",
                     "algorithmic_flops_per_utt_iter": 2 * flops_pass / B,
                     "share_of_step": gemm_ms / tot_prof if tot_prof else None},
        "kernel_ms_per_step": {k: round(v[0], 3) for k, v in prof.items() if v[1]},
        "attack_success_rate": float(sum(success)) / len(success),
    }
    if not args.no_cpu_baseline and world == 1:
        v, sample, cores = iv_cpu_throughput(p, N)
        out["cpu_baseline"] = {"value": v, "unit": "utt-iter/s", "cores": cores, "kind": "port", "sample": sample}
    print(json.dumps(out))
    sys.stdout.flush()


def antrain_cpu_throughput(N, budget_s=12.0, batch=8, pgd_iters=10):
    """utterances/s of one adversarial-training step (PGD-10 on the train-mode model + optimisation step) of the oracle port."""
    from oracle import sg_oracle as O
    import torch.nn.functional as F
    torch.set_num_threads(os.cpu_count() or 1)
    p = O.make_audionet_params(seed=0, num_class=251)
    g = torch.Generator().manual_seed(0)
    x = (torch.rand(batch, N, generator=g) * 2 - 1) * 0.5
    y = torch.randint(0, 251, (batch,), generator=g)
    done, t0 = 0, time.perf_counter()
    while True:
        xa = x.clone()
        for _ in range(pgd_iters):
            xr = xa.clone().requires_grad_(True)
            logits, _ = O.audionet_train_forward(xr, p)
            F.cross_entropy(logits, y, reduction="sum").backward()
            xa = torch.min(torch.max(xa + 0.0004 * xr.grad.sign(), x - 0.002), x + 0.002).clamp(-1, 1)
        O.audionet_train_step(xa, y, p)
        done += 1
        el = time.perf_counter() - t0
        if el >= budget_s:
            break
    return batch * done / el, f"oracle: {done} steps of (PGD-{pgd_iters} + train step), B={batch}, {N / 16000:g} s, {el:.1f} s of CPU work", \
        torch.get_num_threads()


def run_antrain(args):
    """Adversarial training of AudioNet (adver_train.py:183-221): per step PGD-10 on the train-mode model, then forward /
    backward with parameter gradients and Adam.  Metric: training utterances per second."""
    from speakerguard_b200 import dist
    from speakerguard_b200.attack.PGD import PGD
    from speakerguard_b200.model.audionet_csine import audionet_csine
    rank, world, local = dist.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: speakerguard_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B, N, iters = args.batch, int(args.seconds * 16000), args.iters
    torch.manual_seed(rank)
    model = audionet_csine(num_class=251, device=dev)
    model.train()
    opt = torch.optim.Adam(model.parameters())
    attacker = PGD(model, targeted=False, step_size=0.0004, epsilon=0.002, max_iter=iters, batch_size=B, loss="Entropy", verbose=0)
    crit = torch.nn.CrossEntropyLoss()
    g = torch.Generator().manual_seed(100 + rank)
    x_host = ((torch.rand(B, 1, N, generator=g) * 2 - 1) * 0.5).pin_memory()
    y_host = torch.randint(0, 251, (B,), generator=g).pin_memory()
    x_dev, y_dev = x_host.to(dev), y_host.to(dev)

    def step(x, y):
        adv, _ = attacker.attack(x, y)
        out = model(adv)
        loss = crit(out, y)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return loss

    for _ in range(args.warmup):
        step(x_dev, y_dev)
    torch.cuda.synchronize()
    dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    model.engine.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step(x_dev, y_dev)
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    launches = model.engine.launch_count()
    clocks = sampler.stop() if rank == 0 else None
    ms_step = dist.max_over_ranks(e0.elapsed_time(e1), dev) / args.steps
    t0 = time.perf_counter()
    for _ in range(max(args.e2e_steps, 1)):
        l_ = step(x_host.to(dev, non_blocking=True), y_host.to(dev, non_blocking=True))
        float(l_)                                                   # device -> host read of the step's loss
    e2e_s = dist.max_over_ranks(time.perf_counter() - t0, dev) / max(args.e2e_steps, 1)
    model.engine.profile(True)
    step(x_dev, y_dev)
    prof = model.engine.profile_read()
    model.engine.profile(False)
    if world > 1:
        import torch.distributed as tdist
        tdist.barrier()
        tdist.destroy_process_group()
    if rank != 0:
        return
    T = 1 + (N - 1) // 160
    out = {"metric": "adversarial-training utterances/s (AudioNet, PGD-%d inside the step)" % iters, "value": world * B / (ms_step / 1000.0),
           "unit": "utt/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "adver_train.py step: PGD-%d (eps 0.002) on the train-mode AudioNet (251 classes) + CE + Adam, "
                                  "synthetic %g s utterances, batch %d per GPU; replicas are independent (no gradient all-reduce: "
                                  "the reference trains on one device)" % (iters, args.seconds, B), "batch_per_gpu": B, "samples": N,
                      "frames": T},
           "clocks": clocks, "gpu_launches": launches,
           "e2e": {"value": world * B / e2e_s, "unit": "utt/s", "h2d_bytes_per_step": B * N * 4 + B * 8, "d2h_bytes_per_step": 4,
                   "ms_per_step": e2e_s * 1000.0},
           "kernel_ms_per_step": {k: round(v[0], 3) for k, v in prof.items() if v[1]},
           "roofline": None, "final_loss": float(loss)}
    if not args.no_cpu_baseline and world == 1:
        v, sample, cores = antrain_cpu_throughput(N, pgd_iters=iters)
        out["cpu_baseline"] = {"value": v, "unit": "utt/s", "cores": cores, "kind": "port", "sample": sample}
    print(json.dumps(out))
    sys.stdout.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="xv", choices=["xv", "iv", "antrain"],
                    help="xv: the headline (BASELINE configs[1]); iv: configs[4], PGD vs iv_plda (defaults B=256, 5 s, 50 iterations)")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("SGB200_PRECISION", "bf16"), choices=["fp32", "tf32", "bf16"])
    ap.add_argument("--batch", type=int, default=1024, help="utterances per GPU")
    ap.add_argument("--seconds", type=float, default=3.0)
    ap.add_argument("--iters", type=int, default=100, help="PGD iterations per attack")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--ref-budget", type=float, default=30.0, help="CPU seconds for --impl reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.workload == "iv":
        if args.impl == "reference":
            raise SystemExit("--impl reference is defined for the headline workload only")
        defaults = {"batch": 256, "seconds": 5.0, "iters": 50}
        for k, v in defaults.items():
            if getattr(args, k) == ap.get_default(k):
                setattr(args, k, v)
        run_iv(args)
        return
    if args.workload == "antrain":
        if args.impl == "reference":
            raise SystemExit("--impl reference is defined for the headline workload only")
        for k, v in {"batch": 128, "seconds": 3.0, "iters": 10}.items():
            if getattr(args, k) == ap.get_default(k):
                setattr(args, k, v)
        run_antrain(args)
        return
    if args.impl == "reference":
        run_reference(args)
        return

    from speakerguard_b200 import _lib, dist
    from speakerguard_b200.attack.PGD import PGD
    from speakerguard_b200.engine import make_loss_params
    from speakerguard_b200.model.xv_plda import xv_plda
    from speakerguard_b200.synthetic import make_xv_params, state_dict_of, synthetic_batch, write_xv_model_files

    rank, world, local = dist.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: speakerguard_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B, N, iters = args.batch, int(args.seconds * 16000), args.iters
    p = make_xv_params(0)
    tmp = tempfile.mkdtemp(prefix=f"sgb200_bench_{rank}_")
    files = write_xv_model_files(p, tmp)
    model = xv_plda(state_dict_of(p), files["plda.txt"], files["mean.vec"], files["transform.txt"],
                    model_file=files["speaker_model"], device=dev, precision=args.precision, dither="philox", seed=rank)
    eng = model.engine
    m = eng.num_frames(N)
    x_host, y_host = synthetic_batch(B, N, 10, seed=1234 + rank)
    x_host, y_host = x_host.pin_memory(), y_host.pin_memory()
    x_dev, y_dev = x_host.to(dev), y_host.to(dev)
    lp = make_loss_params("Entropy", False, "CSI")
    ws = eng.pgd_ws(B, N)
    xa = torch.empty(B, N, device=dev)

    def step(seed):
        xa.copy_(x_dev[:, 0])
        return eng.pgd_run(xa, x_dev[:, 0], y_dev, max_iter=iters, epsilon=0.002, step_size=0.0004, lp=lp,
                           dither_mode=_lib.DITHER_PHILOX, seed=seed, ws=ws)

    for w in range(args.warmup):
        step(1000 + w)
    torch.cuda.synchronize()
    dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eng.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for k in range(args.steps):
        dec, scores, _ = step(2000 + k)
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    launches = eng.launch_count()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = dist.max_over_ranks(e0.elapsed_time(e1), dev)
    ms_step = ms_total / args.steps
    value = world * B * iters / (ms_step / 1000.0)

    # ---- end-to-end through the public API, host buffers -------------------------------------------
    attacker = PGD(model, task="CSI", epsilon=0.002, step_size=0.0004, max_iter=iters, batch_size=B, verbose=0)
    adv_host = torch.empty(B, 1, N).pin_memory()

    def e2e_step():
        xd = x_host.to(dev, non_blocking=True)
        yd = y_host.to(dev, non_blocking=True)
        adv, success = attacker.attack(xd, yd)
        adv_host.copy_(adv, non_blocking=True)
        torch.cuda.synchronize()
        return adv, success

    adv, success = e2e_step()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        adv, success = e2e_step()
    e2e_s = dist.max_over_ranks(time.perf_counter() - t0, dev) / max(args.e2e_steps, 1)
    if args.e2e_steps == 0:
        e2e_s = float("inf")
    e2e_value = world * B * iters / e2e_s
    metrics = dist.reduce_metrics(dist.attack_metrics(x_dev, adv, success))     # NCCL: metric scalars only

    # ---- per-kernel device time of one profiled step, roofline of the TDNN contraction -------------
    eng.profile(True)
    step(3000)
    prof = eng.profile_read()
    eng.profile(False)
    tot_prof = sum(v[0] for v in prof.values())
    # The roofline kernel is the plain contraction (conv_tc_kernel: 5 forward + 4 or 5 dgrad launches per pass).  With
    # SG_OPT_POOL_FUSION the layer-5 dgrad launch also applies the statistics-pooling adjoint to its A tiles (HBM-side
    # work of the former pool_bwd pass); it is timed in its own category and reported beside the plain launches.
    fused_ms, fused_launches = prof.get("tdnn_dgrad5_pool", (0.0, 0))
    tdnn_ms = prof["tdnn_fwd"][0] + prof["tdnn_dgrad"][0]
    tdnn_launches = prof["tdnn_fwd"][1] + prof["tdnn_dgrad"][1]
    passes = iters + 1
    flops_all = tdnn_flops_per_utt(m) * B * iters + 0.5 * tdnn_flops_per_utt(m) * B   # + forward of the evaluation pass
    flops_l5d = 2.0 * TDNN[4][0] * TDNN[4][1] * TDNN[4][2] * tdnn_valid_frames(m)[4] * B * iters if fused_launches else 0.0
    flops = flops_all - flops_l5d
    achieved = flops / (tdnn_ms / 1000.0) / 1e12
    fused = None
    if fused_launches:
        hbm = peaks_hbm_gbs()
        fused_bytes = (m * B * 1536 * 2 + m * B * 512 * 2) * iters                # r5 read once + dA4 written, bf16
        fused = {"kernel": "conv_tc_kernel<XFORM>: layer-5 dgrad + statistics-pooling adjoint on the staged tiles",
                 "launches": fused_launches, "avg_launch_ms": fused_ms / fused_launches,
                 "tflops": flops_l5d / (fused_ms / 1000.0) / 1e12,
                 "hbm_gbs": fused_bytes / (fused_ms / 1000.0) / 1e9, "hbm_frac": fused_bytes / (fused_ms / 1000.0) / 1e9 / hbm,
                 "note": "replaces pool_bwd (read r5, write dA5: 1.9 GB) + the plain layer-5 dgrad (read dA5): -0.09 ms per pass net"}
    peaks, peak_src = measured_peaks()
    if args.precision == "bf16":
        peak, peak_note = peaks["bf16_tflops_sustained"], f"bf16 sustained, {peak_src}"
    elif args.precision == "tf32":
        peak, peak_note = peaks["bf16_tflops_sustained"] / 2.0, f"tf32 = half of bf16 sustained, {peak_src}"
    else:
        peak, peak_note = peaks["bf16_tflops_sustained"] / 2.0, \
            f"fp32 FFMA parity mode is not on the tensor pipe; quoted against the tf32 tensor peak (half of bf16 sustained, {peak_src})"

    if world > 1:
        import torch.distributed as tdist
        tdist.barrier()
        tdist.destroy_process_group()
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "conv_tc_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(args.precision, {}).get("dram_bytes_per_launch_avg")
    if rank != 0:
        return
    out = {
        "metric": "PGD utterance-iterations/s vs xv_plda", "value": value, "unit": "utt-iter/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": {"fp32": "f32", "tf32": "tf32", "bf16": "bf16"}[args.precision], "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": B, "samples": N, "frames": m, "pgd_iters": iters,
                   "passes_per_step": passes, "dither": "philox (in-kernel N(0,1), fresh per pass)",
                   "cache": "inputs larger than L2 (activations 3.8-7.6 GB per pass)", "precision": args.precision,
                   "precision_note": PRECISION_NOTE[args.precision]},
        "clocks": clocks, "gpu_launches": launches,
        "e2e": {"value": e2e_value, "unit": "utt-iter/s", "h2d_bytes_per_step": B * N * 4 + B * 8,
                "d2h_bytes_per_step": B * N * 4 + B * 8, "ms_per_step": e2e_s * 1000.0,
                "api": "speakerguard_b200.attack.PGD(model).attack(x, y) with pinned host buffers"},
        "roofline": {"bound": "tensor", "kernel": "TDNN conv-as-GEMM, conv_tc_kernel (plain forward + dgrad launches; the layer-5 dgrad with the fused pooling adjoint is listed under fused_l5_dgrad and included in all_tdnn_launches)", "achieved": achieved,
                     "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                     "traffic_note": "dram__bytes_read+write per launch, mean over the TDNN launches of one pass (ncu, profiles/conv_tc_traffic.json)",
                     "peak_source": peak_note,
                     "launches": tdnn_launches, "avg_launch_ms": tdnn_ms / max(tdnn_launches, 1),
                     "algorithmic_flops_per_utt_iter": tdnn_flops_per_utt(m),
                     "share_of_step": tdnn_ms / tot_prof if tot_prof else None,
                     "all_tdnn_launches": {"tflops": flops_all / ((tdnn_ms + fused_ms) / 1000.0) / 1e12,
                                           "frac": flops_all / ((tdnn_ms + fused_ms) / 1000.0) / 1e12 / peak,
                                           "share_of_step": (tdnn_ms + fused_ms) / tot_prof if tot_prof else None},
                     "fused_l5_dgrad": fused},
        "kernel_ms_per_step": {k: round(v[0], 3) for k, v in prof.items()},
        "attack_metrics": metrics,
    }
    if not args.no_cpu_baseline and world == 1:
        v, sample, cores = cpu_reference_throughput(p, N)
        out["cpu_baseline"] = {"value": v, "unit": "utt-iter/s", "cores": cores, "kind": "port", "sample": sample}
    print(json.dumps(out))
    sys.stdout.flush()


if __name__ == "__main__":
    main()
