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

# The ONE JSON line goes to the real stdout; everything else the run prints (the attack classes keep the reference's
# progress prints, attack/FGSM.py:30-34) goes to stderr so that stdout stays machine-readable.
_STDOUT = sys.stdout
sys.stdout = sys.stderr


def _emit(line):
    _STDOUT.write(line + "\n")
    _STDOUT.flush()


WORKLOAD = "PGD-100 Linf eps=0.002 step=0.0004 CE-untargeted vs xv_plda CSI-E (random-init TDNN+PLDA, L=200, S=10), " \
           "synthetic 3 s 16 kHz utterances, batch 1024 per GPU"
PRECISION_NOTE = {
    "fp32": "FFMA everywhere: parity mode, 1e-4 relative vs the reference (tests/test_gpu_xv.py)",
    "tf32": "TDNN contractions on tcgen05 kind::tf32, fp32 storage (the reference's own GPU default for cuDNN convs); "
            "vs the reference goldens: tests/test_gpu_precision.py (embeddings / scores / decisions / PGD-100 outcome)",
    "bf16": "TDNN activations/gradients/weights in bf16, fp32 accumulate (tcgen05 kind::f16), everything else fp32; "
            "vs the reference goldens: tests/test_gpu_precision.py (embeddings / scores / decisions / PGD-100 outcome); the "
            "tf32 and fp32 modes are in precision_ladder",
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


def cpu_port_throughput(p, N, budget_s=12.0, batch=8, iters=2):
    """utt-iter/s of the oracle PORT of the reference path (vectorised torch-CPU restatement; PGD, CE, B=8) on the host cores."""
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
    return batch * done / el, f"oracle port PGD-{done} (incl. {done // iters} evaluation passes), B={batch}, {N / 16000:g} s, " \
                              f"{el:.1f} s of CPU work", torch.get_num_threads()


_REF = {}


def reference_runner(p):
    """The unmodified reference classes (baseline/_ref, copied by tools/install_reference.py) or None when absent."""
    if "r" not in _REF:
        try:
            from oracle.reference_runner import ReferenceXv
            _REF["r"] = ReferenceXv(p)
        except Exception as e:                                   # not installed on this box: fall back to the port, say so
            _REF["r"], _REF["why"] = None, f"{type(e).__name__}: {e}"
    return _REF["r"]


def cpu_reference_baseline(p, N, pgd_iters=10, batch=8):
    """SURVEY 8(d) / BASELINE.md 5: stock ``PGD(model).attack(x, y)`` of the unmodified reference on the host cores, PGD-10 /
    B = 8 / 3 s, plus config 1 exactly (FGSM eps 0.002, B = 8, 2 s).  -> cpu_baseline dict."""
    from speakerguard_b200.synthetic import synthetic_batch
    r = reference_runner(p)
    if r is None:
        v, sample, cores = cpu_port_throughput(p, N)
        return {"value": v, "unit": "utt-iter/s", "cores": cores, "kind": "port", "sample": sample,
                "note": "unmodified reference not installed on this box (%s)" % _REF.get("why", "")}
    x, y = synthetic_batch(batch, N, p["enroll"].shape[0])
    r.attack_seconds("PGD", x, y, epsilon=0.002, step_size=0.0004, max_iter=1)          # warm-up (thread pools, table caches)
    t = r.attack_seconds("PGD", x, y, epsilon=0.002, step_size=0.0004, max_iter=pgd_iters)
    x1, y1 = synthetic_batch(8, 32000, p["enroll"].shape[0])
    t1 = min(r.attack_seconds("FGSM", x1, y1, epsilon=0.002) for _ in range(2))
    return {"value": batch * pgd_iters / t, "unit": "utt-iter/s", "cores": r.threads, "kind": "reference",
            "sample": f"unmodified reference attack/PGD.py PGD-{pgd_iters} (+1 evaluation pass), B={batch}, {N / 16000:g} s: "
                      f"{t:.2f} s wall on {r.threads} threads",
            "config1_fgsm_b8_2s": {"value": 8 / t1, "unit": "utt-iter/s", "wall_s": t1},
            "source": os.path.relpath(r.dir, ROOT) if r.dir.startswith(ROOT) else r.dir}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (unmodified attack/PGD.py over model/xv_plda.py
    from baseline/_ref; the oracle port only if that copy is missing), all host threads, each step a bounded sample of the
    headline workload: PGD-k on B = 8 utterances of 3 s, k sized so the whole run ends within --ref-budget seconds."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from speakerguard_b200.synthetic import make_xv_params, synthetic_batch
    p = make_xv_params(0)
    N = int(args.seconds * 16000)
    r = reference_runner(p)
    batch = 8
    if r is None:
        for _ in range(args.warmup):
            cpu_port_throughput(p, N, budget_s=0.0, iters=1)
        t0 = time.perf_counter()
        vals = [cpu_port_throughput(p, N, budget_s=args.ref_budget / max(args.steps, 1), iters=2) for _ in range(args.steps)]
        el = time.perf_counter() - t0
        value, sample, cores, kind = sum(v[0] for v in vals) / len(vals), vals[-1][1], vals[-1][2], "port"
        note = "unmodified reference not installed (%s): oracle port" % _REF.get("why", "")
    else:
        x, y = synthetic_batch(batch, N, p["enroll"].shape[0])
        kw = dict(epsilon=0.002, step_size=0.0004)
        t2 = r.attack_seconds("PGD", x, y, max_iter=2, **kw)                       # 3 passes: cost of one pass
        per_iter = t2 / 3.0
        k = int(max(2, min(10, args.ref_budget / max(args.steps + args.warmup, 1) / per_iter - 1)))
        for _ in range(args.warmup):
            r.attack_seconds("PGD", x, y, max_iter=k, **kw)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            r.attack_seconds("PGD", x, y, max_iter=k, **kw)
        el = time.perf_counter() - t0
        value, cores, kind = batch * k * args.steps / el, r.threads, "reference"
        sample = f"unmodified reference attack/PGD.py over model/xv_plda.py: {args.steps} x PGD-{k} (+1 evaluation pass each), " \
                 f"B={batch}, {N / 16000:g} s, {el:.1f} s of CPU work on {cores} threads"
        note = "reference arm = the reference's own classes from baseline/_ref (unmodified copy of /root/reference, import shims " \
               "kaldi_io / np.infty only) on a bounded sample of the same workload: B=8, PGD-%d per step" % k
    out = {"impl": "reference", "metric": "PGD utterance-iterations/s vs xv_plda", "value": value, "unit": "utt-iter/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * el / args.steps,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": WORKLOAD, "note": note},
           "cpu_baseline": {"value": value, "unit": "utt-iter/s", "cores": cores, "kind": kind, "sample": sample},
           "e2e": {"value": value, "unit": "utt-iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(json.dumps(out))


def measure_matmul_peak(dev, dtype: str, seconds: float = 2.0):
    """8192^3 matmul through cuBLAS, the method of MEASURED_PEAKS.json: best of 10 (burst) and back to back for `seconds`
    (sustained).  dtype 'tf32' = fp32 tensors with allow_tf32, 'bf16' = bf16 tensors.  -> (burst, sustained) TFLOP/s."""
    n = 8192
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    td = torch.bfloat16 if dtype == "bf16" else torch.float32
    a, b = torch.randn(n, n, device=dev, dtype=td), torch.randn(n, n, device=dev, dtype=td)
    c = torch.empty(n, n, device=dev, dtype=td)
    for _ in range(3):
        torch.matmul(a, b, out=c)
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    reps = max(10, int(seconds * 1000.0 / best))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        torch.matmul(a, b, out=c)
    e1.record()
    e1.synchronize()
    torch.backends.cuda.matmul.allow_tf32 = prev
    fl = 2.0 * n ** 3
    return fl / (best / 1e3) / 1e12, fl * reps / (e0.elapsed_time(e1) / 1e3) / 1e12


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
    _emit(json.dumps(out))
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
    _emit(json.dumps(out))
    sys.stdout.flush()


def cw2_cpu_throughput(N, budget_s=12.0, batch=4, iters=20):
    """utt-iter/s of the oracle port of attack/CW2.py:41-132 against the AudioNet restatement on the host cores."""
    from oracle import sg_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    p = O.make_audionet_params(seed=0, num_class=251)
    g = torch.Generator().manual_seed(0)
    x = (torch.rand(batch, N, generator=g) * 2 - 1) * 0.5
    y = torch.randint(0, 251, (batch,), generator=g)
    done, t0 = 0, time.perf_counter()
    while True:
        O.cw2_attack(x, y, lambda z: O.audionet_forward(z, p), targeted=True, initial_const=1e-3, binary_search_steps=1,
                     max_iter=iters, stop_early=True, stop_early_iter=1000, lr=1e-2)
        done += iters
        el = time.perf_counter() - t0
        if el >= budget_s:
            break
    return batch * done / el, f"oracle port CW2 {done} iterations (1 search step per call, +1 evaluation pass), B={batch}, " \
                              f"{N / 16000:g} s, {el:.1f} s of CPU work", torch.get_num_threads()


def run_cw2(args):
    """BASELINE configs[2]: CW2 targeted, 9 binary-search steps x 1000 iterations (c0 1e-3, lr 1e-2, stop_early_iter 1000) vs
    AudioNet CSI-NE (251 classes, random init, eval), batch 512 x 3 s per GPU, whole attack on the device
    (sg_cw2_audionet_run).  One step = one attack.  HBM-side path: ~1.4 MB of waveform-sized traffic per utterance-iteration."""
    from speakerguard_b200 import dist
    from speakerguard_b200.attack.CW2 import CW2
    from speakerguard_b200.model.audionet_csine import audionet_csine
    from speakerguard_b200.synthetic import synthetic_batch
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from bench_configs import audionet_params
    rank, world, local = dist.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: speakerguard_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B, N, iters, bs = args.batch, int(args.seconds * 16000), args.iters, args.search_steps
    an_prec = "fp32" if args.precision == "fp32" else "tf32"    # AudioNet's convolutions: FFMA parity mode or tcgen05 / TF32
    an = audionet_csine(params=audionet_params(), device=dev, precision=an_prec)
    eng = an.engine
    x_host, _ = synthetic_batch(B, N, 10, seed=1234 + rank)
    x_host = x_host.pin_memory()
    x_dev = x_host.to(dev)
    with torch.no_grad():
        pred = an(x_dev).argmax(1)
    g = torch.Generator().manual_seed(7 + rank)
    y_host = ((pred.cpu() + torch.randint(1, 251, (B,), generator=g)) % 251).pin_memory()      # targets != prediction
    y_dev = y_host.to(dev)
    mk = lambda s, it: CW2(an, targeted=True, initial_const=1e-3, binary_search_steps=s, max_iter=it, stop_early=True,
                           stop_early_iter=1000, lr=1e-2, batch_size=B, verbose=0)
    warm, att = mk(1, 20), mk(bs, iters)
    for _ in range(max(args.warmup, 1)):
        warm.attack(x_dev, y_dev)
    torch.cuda.synchronize()
    dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eng.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        adv, success = att.attack(x_dev, y_dev)
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    launches = eng.launch_count()
    clocks = sampler.stop() if rank == 0 else None
    ms_step = dist.max_over_ranks(e0.elapsed_time(e1), dev) / args.steps
    done_iters = eng.cw2_iterations() or bs * iters          # gradient iterations actually executed by the last attack
    value = world * B * done_iters / (ms_step / 1000.0)
    adv_host = torch.empty(B, 1, N).pin_memory()
    t0 = time.perf_counter()
    a, s_ = att.attack(x_host.to(dev, non_blocking=True), y_host.to(dev, non_blocking=True))
    adv_host.copy_(a, non_blocking=True)
    torch.cuda.synchronize()
    e2e_s = dist.max_over_ranks(time.perf_counter() - t0, dev)
    eng.profile(True)
    mk(1, 50).attack(x_dev, y_dev)
    prof = eng.profile_read()
    eng.profile(False)
    if world > 1:
        import torch.distributed as tdist
        tdist.barrier()
        tdist.destroy_process_group()
    if rank != 0:
        return
    l2 = float((adv - x_dev).pow(2).sum((1, 2)).sqrt().mean())
    peaks, peak_src = measured_peaks()
    hbm = peaks.get("hbm_gbs", 6548.5)
    per_it = 1.4e6 * N / 48000.0
    out = {"metric": "CW2 utterance-iterations/s vs AudioNet", "value": value, "unit": "utt-iter/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32" if an_prec == "fp32" else "tf32", "data": "synthetic",
           "config": {"precision": an_prec,
                      "workload": "CW2 targeted (%d binary-search steps x %d iterations, c0 1e-3, lr 1e-2, stop_early_iter 1000) vs AudioNet "
                                  "CSI-NE (251 classes, random init, eval), synthetic %g s utterances, batch %d per GPU (BASELINE configs[2])"
                                  % (bs, iters, args.seconds, B), "batch_per_gpu": B, "samples": N, "iterations_executed": done_iters,
                      "cache": "inputs larger than L2 (iterate + Adam state 0.5 GB, activations 1.3 GB per pass)"},
           "clocks": clocks, "gpu_launches": launches,
           "e2e": {"value": world * B * done_iters / e2e_s, "unit": "utt-iter/s", "h2d_bytes_per_step": B * N * 4 + B * 8,
                   "d2h_bytes_per_step": B * N * 4, "ms_per_step": e2e_s * 1000.0,
                   "api": "speakerguard_b200.attack.CW2(model).attack(x, y) with pinned host buffers"},
           "roofline": {"bound": "hbm", "kernel": "whole CW2 iteration (tanh / L2 prepare, log-mel fwd, CNN fwd + bwd, log-mel adjoint, "
                        "Adam, best tracking)", "achieved": per_it * B * done_iters / (ms_step / 1000.0) / 1e9, "peak": hbm,
                        "unit": "GB/s", "frac": per_it * B * done_iters / (ms_step / 1000.0) / 1e9 / hbm, "traffic": None,
                        "algorithmic_bytes_per_utt_iter": per_it, "peak_source": peak_src},
           "kernel_ms_profiled_50_iters": {k: round(v[0], 3) for k, v in prof.items() if v[1]},
           "attack_metrics": {"success_rate": float(sum(success)) / len(success), "l2": l2}}
    if not args.no_cpu_baseline and world == 1:
        v, sample, cores = cw2_cpu_throughput(N)
        out["cpu_baseline"] = {"value": v, "unit": "utt-iter/s", "cores": cores, "kind": "port", "sample": sample}
    _emit(json.dumps(out))
    sys.stdout.flush()


FFMA_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12     # 148 SMs x 128 fp32 FMA lanes x 2 flop x max SM clock: 74.5 TFLOP/s nominal


def layer_flops(m: int):
    """Algorithmic forward FLOPs of each TDNN layer for one utterance (valid frames, unpadded channels)."""
    return [2.0 * ci * k * co * t for (ci, co, k, d), t in zip(TDNN, tdnn_valid_frames(m))]


def xv_leg(args, precision, steps, warmup, e2e_steps, rank, world, local, want_metrics=True, clock_sampler=None):
    """One precision mode of the headline workload: device-timed value, end-to-end value through the public API, per-kernel
    profile and the roofline of the TDNN contraction over ALL its launches (with the per-layer split)."""
    from speakerguard_b200 import _lib, dist
    from speakerguard_b200.attack.PGD import PGD
    from speakerguard_b200.engine import make_loss_params
    from speakerguard_b200.model.xv_plda import xv_plda
    from speakerguard_b200.synthetic import make_xv_params, state_dict_of, synthetic_batch, write_xv_model_files
    dev = torch.device("cuda", local)
    N, iters = int(args.seconds * 16000), args.iters
    if args.scaling == "strong":                     # GPU r takes x[r*B/G:(r+1)*B/G] of ONE global batch
        lo, hi = dist.shard_bounds(args.batch, rank, world)
        B, global_B, offset = hi - lo, args.batch, lo
    else:
        B, global_B, offset = args.batch, args.batch * world, rank * args.batch
    p = make_xv_params(0)
    tmp = tempfile.mkdtemp(prefix=f"sgb200_bench_{rank}_")
    files = write_xv_model_files(p, tmp)
    # one seed for the whole job: the philox dither is keyed on the GLOBAL utterance index (utt_offset), so the sharded job
    # draws exactly what a single GPU would draw for the same utterances
    model = xv_plda(state_dict_of(p), files["plda.txt"], files["mean.vec"], files["transform.txt"],
                    model_file=files["speaker_model"], device=dev, precision=precision, dither="philox", seed=0)
    eng = model.engine
    m = eng.num_frames(N)
    xg, yg = synthetic_batch(global_B, N, 10, seed=1234)
    x_host, y_host = xg[offset:offset + B].clone().pin_memory(), yg[offset:offset + B].clone().pin_memory()
    del xg, yg
    x_dev, y_dev = x_host.to(dev), y_host.to(dev)
    lp = make_loss_params("Entropy", False, "CSI")
    ws = eng.pgd_ws(B, N)
    xa = torch.empty(B, N, device=dev)

    def step(seed):
        xa.copy_(x_dev[:, 0])
        return eng.pgd_run(xa, x_dev[:, 0], y_dev, max_iter=iters, epsilon=0.002, step_size=0.0004, lp=lp,
                           dither_mode=_lib.DITHER_PHILOX, seed=seed, ws=ws, grad_sign=1.0, utt_offset=offset)

    for w in range(warmup):
        step(1000 + w)
    torch.cuda.synchronize()
    dist.barrier()
    if clock_sampler is not None:
        clock_sampler.start()
    eng.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for k in range(steps):
        dec, scores, _ = step(2000 + k)
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    launches = eng.launch_count()
    clocks = clock_sampler.stop() if clock_sampler is not None else None
    ms_step = dist.max_over_ranks(e0.elapsed_time(e1), dev) / steps
    value = global_B * iters / (ms_step / 1000.0)

    # ---- end-to-end through the public API, host buffers -------------------------------------------
    attacker = PGD(model, task="CSI", epsilon=0.002, step_size=0.0004, max_iter=iters, batch_size=B, verbose=0)
    attacker.utt_offset = offset
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
    # e2e: every step uploads its batch from pinned host memory and downloads its adversarial batch, through
    # speakerguard_b200.io.attack_stream (the attackMain.py:306-333 loop with the copies of neighbouring batches overlapped
    # with the attack on a side stream; pipeline fill and drain are inside the timed region)
    from speakerguard_b200.io import attack_stream
    adv_hosts = [adv_host, torch.empty(B, 1, N).pin_memory()]
    if e2e_steps:                                                    # warm the side stream's allocator pool and the pipeline
        for _adv_h, _succ in attack_stream(attacker, ((x_host, y_host) for _ in range(2)), device=dev, out=adv_hosts):
            pass
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _adv_h, _succ in attack_stream(attacker, ((x_host, y_host) for _ in range(e2e_steps)), device=dev, out=adv_hosts):
        pass
    torch.cuda.synchronize()
    e2e_s = dist.max_over_ranks(time.perf_counter() - t0, dev) / max(e2e_steps, 1)
    e2e_value = global_B * iters / e2e_s if e2e_steps else None
    # the same steps strictly one after the other (upload, attack, download, sync): what the overlap buys
    t0 = time.perf_counter()
    for _ in range(min(e2e_steps, 1)):
        adv, success = e2e_step()
    e2e_serial_s = dist.max_over_ranks(time.perf_counter() - t0, dev) / max(min(e2e_steps, 1), 1)
    if want_metrics and world > 1:
        dist.engine_comm_init(eng)                                   # libsgb200's own NCCL communicator (sg_comm_init)
    # the path's only collective: five metric scalars, summed over NVLink by sg_allreduce_metrics
    metrics = dist.reduce_metrics(dist.attack_metrics(x_dev, adv, success), engine=eng) if want_metrics else None

    # ---- per-kernel device time of one profiled step, roofline of the TDNN contraction -------------
    eng.profile(True)
    step(3000)
    prof = eng.profile_read()
    launches_prof = eng.profile_dump()
    eng.profile(False)
    tot_prof = sum(v[0] for v in prof.values())
    fused_ms, fused_launches = prof.get("tdnn_dgrad5_pool", (0.0, 0))
    tdnn_ms = prof["tdnn_fwd"][0] + prof["tdnn_dgrad"][0] + fused_ms
    tdnn_launches = prof["tdnn_fwd"][1] + prof["tdnn_dgrad"][1] + fused_launches
    fl_utt = tdnn_flops_per_utt(m)
    flops_all = fl_utt * B * iters + 0.5 * fl_utt * B            # + forward of the evaluation pass
    achieved = flops_all / (tdnn_ms / 1000.0) / 1e12
    peaks, peak_src = measured_peaks()
    if precision == "bf16":
        peak, peak_note = peaks["bf16_tflops_sustained"], f"bf16 sustained, {peak_src}"
    else:
        peak, peak_note = args.tf32_peak["sustained"], args.tf32_peak["how"]
    # per-layer split (every launch is tagged with its layer by the library)
    lf = layer_flops(m)
    per = {}
    for cat, tag, ms in launches_prof:
        if cat in ("tdnn_fwd", "tdnn_dgrad", "tdnn_dgrad5_pool") and tag:
            key = ("fwd" if cat == "tdnn_fwd" else "dgrad") + f"_l{tag}" + ("+pool_adjoint" if cat == "tdnn_dgrad5_pool" else "")
            e = per.setdefault(key, [0.0, 0, tag])
            e[0] += ms
            e[1] += 1
    n_pass_f, n_pass_b = iters + 1, iters
    layers = {}
    for key, (ms, n, tag) in sorted(per.items()):
        npass = n_pass_f if key.startswith("fwd") else n_pass_b
        fl = lf[tag - 1] * B * npass
        layers[key] = {"ms_per_pass": ms / npass, "launches_per_pass": n / npass, "tflops": fl / (ms / 1000.0) / 1e12,
                       "frac": fl / (ms / 1000.0) / 1e12 / peak}
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "conv_tc_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
            traffic = (tj.get(precision + "_round2") or tj.get(precision, {})).get("dram_bytes_per_launch_avg")
    roof = {"bound": "tensor",
            "kernel": "TDNN conv-as-GEMM: ALL forward + dgrad launches of the step (conv_tc_kernel incl. the layer-5 dgrad with the "
                      "fused pooling adjoint and layer 1's tap gather)" if precision != "fp32" else
                      "TDNN conv-as-GEMM in fp32 FFMA parity mode (conv_simt_kernel): not on the tensor pipe",
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
            "traffic_note": "dram__bytes_read+write per launch, mean over the TDNN launches of one pass (ncu, profiles/conv_tc_traffic.json)",
            "peak_source": peak_note, "launches": tdnn_launches, "avg_launch_ms": tdnn_ms / max(tdnn_launches, 1),
            "algorithmic_flops_per_utt_iter": fl_utt, "share_of_step": tdnn_ms / tot_prof if tot_prof else None,
            "step_tflops": flops_all / (ms_step * B / (global_B / world) / 1000.0) / 1e12,
            "step_frac": flops_all / (ms_step / 1000.0) / 1e12 / peak, "per_layer": layers}
    if precision == "fp32":
        roof["ffma_peak"] = FFMA_PEAK_TFLOPS
        roof["frac_of_ffma_peak"] = achieved / FFMA_PEAK_TFLOPS
        roof["peak_source"] += "; ffma_peak = 148 SMs x 128 lanes x 2 x 1.965 GHz (nominal)"
    # feature pass (F1 + F2 + CMVN): HBM roofline on SURVEY 8(d)'s algorithmic bytes
    feat_ms = prof["mfcc_fwd"][0] + prof["mfcc_bwd"][0] + prof["cmvn"][0]
    feat_bytes = (192000.0 + 36000.0) * B * (iters + 1) * (N / 48000.0) + (36000.0 + 3 * 192000.0) * B * iters * (N / 48000.0)
    hbm = peaks.get("hbm_gbs", 6548.5)
    feat = {"bound": "hbm", "kernel": "mfcc_fwd + mfcc_bwd(+fused sign step) + cmvn", "ms_per_pass": feat_ms / (iters + 1),
            "achieved": feat_bytes / (feat_ms / 1000.0) / 1e9, "peak": hbm, "unit": "GB/s",
            "frac": feat_bytes / (feat_ms / 1000.0) / 1e9 / hbm, "algorithmic_bytes_per_utt_iter": 840000.0 * N / 48000.0,
            "share_of_step": feat_ms / tot_prof if tot_prof else None}
    return {"precision": precision, "value": value, "ms_per_step": ms_step, "steps": steps, "warmup": warmup, "B": B,
            "global_B": global_B, "m": m, "N": N, "launches": launches, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "utt-iter/s", "h2d_bytes_per_step": B * N * 4 + B * 8,
                    "d2h_bytes_per_step": B * N * 4 + B * 8, "ms_per_step": e2e_s * 1000.0,
                    "ms_per_step_serial_copies": e2e_serial_s * 1000.0,
                    "api": "speakerguard_b200.attack.PGD(model).attack(x, y) with pinned host buffers"},
            "roofline": roof, "feature_roofline": feat, "kernel_ms_per_step": {k: round(v[0], 3) for k, v in prof.items() if v[1]},
            "attack_metrics": metrics, "params": p}


def run_torch_gpu(args):
    """--impl torch_gpu: the batched pure-PyTorch restatement on the same GPU (cuFFT + cuDNN + cuBLAS; tools/torch_gpu_baseline.py)
    on the headline workload - the existing-library bar.  Same JSON contract, "impl": "torch_gpu"."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from torch_gpu_baseline import TorchGpuXv
    from speakerguard_b200 import dist
    from speakerguard_b200.synthetic import make_xv_params, synthetic_batch
    rank, world, local = dist.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B, N, iters = args.batch, int(args.seconds * 16000), args.iters
    m = TorchGpuXv(make_xv_params(0), dev, args.precision)
    x, y = synthetic_batch(B, N, 10, seed=1234 + rank)
    x_dev, y_dev = x[:, 0].to(dev), y.to(dev)
    for _ in range(max(args.warmup, 1)):
        m.pgd(x_dev, y_dev, min(iters, 3))
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        xa, dec = m.pgd(x_dev, y_dev, iters)
    e1.record()
    torch.cuda.synchronize()
    ms_step = dist.max_over_ranks(e0.elapsed_time(e1), dev) / args.steps
    if world > 1:
        import torch.distributed as tdist
        tdist.barrier()
        tdist.destroy_process_group()
    if rank != 0:
        return
    _emit(json.dumps({"impl": "torch_gpu", "metric": "PGD utterance-iterations/s vs xv_plda", "value": world * B * iters / (ms_step / 1e3),
                      "unit": "utt-iter/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
                      "config": {"workload": WORKLOAD, "batch_per_gpu": B, "samples": N, "pgd_iters": iters,
                                 "note": "batched pure-torch restatement (cuFFT rfft, cuDNN conv1d + autograd input gradients, no wgrad: "
                                         "weights frozen), torch.randn dither; tools/torch_gpu_baseline.py"},
                      "success_rate": float((dec != y_dev).float().mean())}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="xv", choices=["xv", "iv", "antrain", "cw2"],
                    help="xv: the headline (BASELINE configs[1]); iv: configs[4], PGD vs iv_plda (defaults B=256, 5 s, 50 iterations); "
                         "cw2: configs[2], CW2 vs AudioNet (defaults B=512, 3 s)")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch_gpu"])
    ap.add_argument("--precision", default=os.environ.get("SGB200_PRECISION", "bf16"), choices=["fp32", "tf32", "bf16"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --batch utterances per GPU; strong: --batch utterances in total, split contiguously over the GPUs")
    ap.add_argument("--batch", type=int, default=1024, help="utterances per GPU (weak) or in total (strong)")
    ap.add_argument("--seconds", type=float, default=3.0)
    ap.add_argument("--iters", type=int, default=100, help="PGD iterations per attack")
    ap.add_argument("--search-steps", type=int, default=9, help="cw2: binary-search steps")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--ref-budget", type=float, default=120.0, help="CPU seconds for --impl reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-peak", action="store_true", help="do not re-measure the TF32 / bf16 matmul peaks (profiling runs only)")
    ap.add_argument("--no-ladder", action="store_true", help="skip the tf32 / fp32 legs of the precision ladder (N = 1 only)")
    args = ap.parse_args()
    if args.workload in ("iv", "antrain", "cw2"):
        if args.impl != "ours":
            raise SystemExit("--impl reference / torch_gpu are defined for the headline workload only")
        defaults = {"iv": {"batch": 256, "seconds": 5.0, "iters": 50}, "antrain": {"batch": 128, "seconds": 3.0, "iters": 10},
                    "cw2": {"batch": 512, "seconds": 3.0, "iters": 1000, "steps": 1, "warmup": 1}}[args.workload]
        for k, v in defaults.items():
            if getattr(args, k) == ap.get_default(k):
                setattr(args, k, v)
        {"iv": run_iv, "antrain": run_antrain, "cw2": run_cw2}[args.workload](args)
        return
    if args.impl == "reference":
        run_reference(args)
        return
    if args.impl == "torch_gpu":
        run_torch_gpu(args)
        return

    from speakerguard_b200 import dist
    rank, world, local = dist.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: speakerguard_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # TF32 tensor peak measured here with MEASURED_PEAKS.json's own method, beside a bf16 re-measurement
    if args.no_peak:       # profiling runs (ncu launch lists): skip the 2 x 2 s of library matmuls, assume tf32 = bf16 / 2
        _pk = measured_peaks()[0]
        bf_b, bf_s = _pk.get("bf16_tflops", 1631.4), _pk.get("bf16_tflops_sustained", 1376.4)
        tf_b, tf_s = bf_b / 2, bf_s / 2
    else:
        tf_b, tf_s = measure_matmul_peak(dev, "tf32")
        bf_b, bf_s = measure_matmul_peak(dev, "bf16")
    args.tf32_peak = {"burst": tf_b, "sustained": tf_s, "bf16_burst_same_run": bf_b, "bf16_sustained_same_run": bf_s,
                      "how": "tf32 sustained: torch.matmul fp32 8192^3 with allow_tf32, back to back for 2 s, CUDA events, measured "
                             "in this run (burst = best of 10)"}
    sampler = ClockSampler(local) if rank == 0 else None
    leg = xv_leg(args, args.precision, args.steps, args.warmup, args.e2e_steps, rank, world, local, clock_sampler=sampler)
    ladder = {}
    if world == 1 and not args.no_ladder:
        for prec, st, wu in (("tf32", 3, 3), ("fp32", 2, 1)):
            if prec == args.precision:
                continue
            torch.cuda.empty_cache()
            lg = xv_leg(args, prec, st, wu, 1, rank, world, local, want_metrics=True)
            ladder[prec] = {"value": lg["value"], "unit": "utt-iter/s", "ms_per_step": lg["ms_per_step"], "steps": st, "warmup": wu,
                            "e2e": lg["e2e"], "roofline": lg["roofline"], "feature_roofline": lg["feature_roofline"],
                            "kernel_ms_per_step": lg["kernel_ms_per_step"], "attack_metrics": lg["attack_metrics"],
                            "precision_note": PRECISION_NOTE[prec]}
    if world > 1:
        import torch.distributed as tdist
        tdist.barrier()
        tdist.destroy_process_group()
    if rank != 0:
        return
    B, N, m, iters = leg["B"], leg["N"], leg["m"], args.iters
    workload = WORKLOAD if args.scaling == "weak" else WORKLOAD.replace("batch 1024 per GPU", f"global batch {args.batch} split over the GPUs")
    out = {
        "metric": "PGD utterance-iterations/s vs xv_plda", "value": leg["value"], "unit": "utt-iter/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": leg["ms_per_step"], "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": {"fp32": "f32", "tf32": "tf32", "bf16": "bf16"}[args.precision],
        "data": "synthetic",
        "config": {"workload": workload, "batch_per_gpu": B, "global_batch": leg["global_B"], "samples": N, "frames": m,
                   "pgd_iters": iters, "passes_per_step": iters + 1,
                   "dither": "philox (in-kernel N(0,1), fresh per pass, keyed on the global utterance index)",
                   "cache": "inputs larger than L2 (activations 3.8-7.6 GB per pass)", "precision": args.precision,
                   "precision_note": PRECISION_NOTE[args.precision]},
        "clocks": leg["clocks"], "gpu_launches": leg["launches"], "e2e": leg["e2e"], "roofline": leg["roofline"],
        "feature_roofline": leg["feature_roofline"], "kernel_ms_per_step": leg["kernel_ms_per_step"],
        "attack_metrics": leg["attack_metrics"], "tf32_peak_measured": {k: v for k, v in args.tf32_peak.items()},
    }
    if ladder:
        out["precision_ladder"] = ladder
    if not args.no_cpu_baseline and world == 1:
        out["cpu_baseline"] = cpu_reference_baseline(leg["params"], N)
        v, sample, cores = cpu_port_throughput(leg["params"], N, budget_s=6.0)
        out["cpu_baseline_port"] = {"value": v, "unit": "utt-iter/s", "cores": cores, "kind": "port", "sample": sample}
    _emit(json.dumps(out))
    sys.stdout.flush()


if __name__ == "__main__":
    main()
