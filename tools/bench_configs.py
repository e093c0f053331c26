#!/usr/bin/env python
"""Secondary measurements for the other BASELINE.json configs (not the driver's headline):
config 1 (FGSM, B=8, 2 s), config 3 (CW2 vs AudioNet, B=512, 3 s) and config 4 (EOT-PGD vs a
FeCo-defended xv_plda).  Prints one JSON line per config; wall-clock with a final synchronise."""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def xv_model(precision, dither="philox"):
    from speakerguard_b200.model.xv_plda import xv_plda
    from speakerguard_b200.synthetic import make_xv_params, state_dict_of, write_xv_model_files
    p = make_xv_params(0)
    f = write_xv_model_files(p, tempfile.mkdtemp(prefix="sgb200_cfg_"))
    return xv_plda(state_dict_of(p), f["plda.txt"], f["mean.vec"], f["transform.txt"], model_file=f["speaker_model"],
                   device="cuda:0", precision=precision, dither=dither)


def audionet_params(seed=0, num_class=251):
    """Random-init AudioNet in torch's default initialisation (engine naming)."""
    torch.manual_seed(seed)
    p = {}
    c1 = torch.nn.Conv2d(1, 1, kernel_size=[5, 5], padding=[2, 2])
    p["conv1.weight"], p["conv1.bias"] = c1.weight.detach(), c1.bias.detach()
    spec = [("conv2", 32, 64), ("conv3", 64, 128), ("conv4", 128, 128), ("conv5", 128, 128), ("conv6", 128, 128),
            ("conv7", 128, 64), ("conv8", 64, 32)]
    for n, ci, co in spec:
        c = torch.nn.Conv1d(ci, co, 3)
        p[f"{n}.weight"], p[f"{n}.bias"] = c.weight.detach(), c.bias.detach()
    fc = torch.nn.Linear(32, num_class)
    p["fc.weight"], p["fc.bias"] = fc.weight.detach(), fc.bias.detach()
    for n, co in [("conv1", 1)] + [(n, co) for n, _, co in spec]:
        p[f"{n}.bn_mean"], p[f"{n}.bn_var"] = torch.zeros(co), torch.ones(co)
        p[f"{n}.bn_gamma"], p[f"{n}.bn_beta"] = torch.ones(co), torch.zeros(co)
    return p


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, out


def main():
    from speakerguard_b200.attack.CW2 import CW2
    from speakerguard_b200.attack.FGSM import FGSM
    from speakerguard_b200.attack.PGD import PGD
    from speakerguard_b200.defense.feature_level import FeCo
    from speakerguard_b200.model.audionet_csine import audionet_csine
    from speakerguard_b200.model.defended_model import defended_model
    from speakerguard_b200.synthetic import synthetic_batch
    prec = os.environ.get("SGB200_PRECISION", "bf16")

    # config 1: FGSM eps 0.002, B=8, 2 s (the reference's CPU-runnable case)
    model = xv_model(prec)
    x, y = synthetic_batch(8, 32000)
    x, y = x.cuda(), y.cuda()
    att = FGSM(model, epsilon=0.002, batch_size=8, verbose=0)
    t, _ = timed(lambda: att.attack(x, y), reps=10)
    print(json.dumps({"config": "1: FGSM B=8 2 s vs xv_plda", "precision": prec, "s_per_attack": t, "utt_iter_per_s": 8 / t}))

    # config 4: PGD-10, EOT_size 50 (batch 50), FeCo kmeans 0.5 at the raw-feature level: the fused device loop (FeCoDefense
    # object: FeCo + EOT inside sg_pgd_run, EOT copies as batch rows) and the generic autograd path (plain lambda)
    from speakerguard_b200.defense.feature_level import FeCoDefense
    exact = bool(os.environ.get("SGB200_EXACT"))       # BASELINE.json sizes: C4 B=256 / PGD-10, C3 targeted 9 x 1000
    B4 = int(os.environ.get("SGB200_CFG4_B", "256" if exact else "32"))
    x4, y4 = synthetic_batch(B4, 48000)
    x4, y4 = x4.cuda(), y4.cuda()
    it4 = 10 if exact else 2
    for path, defense in (("fused", FeCoDefense("kmeans", 0.5, "L2")), ("generic", lambda f: FeCo(f, "kmeans", 0.5, "L2"))):
        dm = defended_model(model, defense=[[1, defense]], order="sequential")
        att4 = PGD(dm, epsilon=0.002, step_size=0.0004, max_iter=it4, batch_size=B4, EOT_size=50, EOT_batch_size=50, verbose=0)
        t, _ = timed(lambda: att4.attack(x4, y4), reps=2 if path == "fused" else 1)
        print(json.dumps({"config": f"4: EOT-PGD-{it4} (EOT 50) vs FeCo(kmeans 0.5)-defended xv_plda, B={B4}, 3 s", "path": path,
                          "precision": prec, "s_per_iteration": t / it4, "utt_iter_per_s": B4 * it4 / t,
                          "eot_utt_passes_per_s": B4 * it4 * 50 / t}), flush=True)
    del att4, dm, model
    torch.cuda.empty_cache()

    # config 3: CW2 vs AudioNet (fused device loop), B=512, 3 s, 1 search step x 100 iterations
    an = audionet_csine(params=audionet_params(), device="cuda:0")
    x3, _ = synthetic_batch(512, 48000)
    x3 = x3.cuda()
    with torch.no_grad():
        y3 = an(x3).argmax(1)
    if exact:
        g = torch.Generator().manual_seed(7)
        tgt = (y3.cpu() + torch.randint(1, 251, (512,), generator=g)) % 251          # targets != prediction
        att3 = CW2(an, targeted=True, initial_const=1e-3, binary_search_steps=9, max_iter=1000, stop_early=True,
                   stop_early_iter=1000, lr=1e-2, batch_size=512, verbose=0)
        t0 = time.perf_counter()
        adv, suc = att3.attack(x3, tgt.cuda())
        torch.cuda.synchronize()
        t = time.perf_counter() - t0
        print(json.dumps({"config": "3: CW2 targeted (9 x 1000 iters, c0 1e-3) vs AudioNet C=251, B=512, 3 s", "s_per_attack": t,
                          "utt_iter_per_s_upper": 512 * 9000 / t, "note": "early stop may end search steps before 1000 iterations",
                          "success_rate": sum(suc) / len(suc)}))
        return
    att3 = CW2(an, targeted=False, initial_const=1e2, binary_search_steps=1, max_iter=100, stop_early=True,
               stop_early_iter=1000, lr=1e-2, batch_size=512, verbose=0)
    t, (adv, suc) = timed(lambda: att3.attack(x3, y3), reps=2)
    print(json.dumps({"config": "3: CW2 (1 x 100 iters) vs AudioNet C=251, B=512, 3 s", "s_per_attack": t,
                      "utt_iter_per_s": 512 * 100 / t, "success_rate": sum(suc) / len(suc)}))


def config5():
    """config 5: PGD vs iv_plda, SV task, 5 s utterances, B=256 (C=2048 UBM, 400-dim i-vectors, LDA 200)."""
    from speakerguard_b200.attack.PGD import PGD
    from speakerguard_b200.model.iv_plda import iv_plda
    from speakerguard_b200.synthetic import make_iv_params, synthetic_batch
    B = int(os.environ.get("SGB200_CFG5_B", "256"))
    iters = int(os.environ.get("SGB200_CFG5_ITERS", "5"))
    t0 = time.perf_counter()
    p = make_iv_params(0, C=2048, F=72, D=400, L=200, S=1)
    t1 = time.perf_counter()
    model = iv_plda(None, None, None, None, None, threshold=0.0, device="cuda:0", params=p,
                    precision=os.environ.get("SGB200_IV_PRECISION", "tf32x3"))
    t2 = time.perf_counter()
    x, _ = synthetic_batch(B, 80000)
    x = x.cuda()
    y = torch.zeros(B, dtype=torch.int64, device="cuda")
    att = PGD(model, task="SV", epsilon=0.002, step_size=0.0004, max_iter=iters, batch_size=B, verbose=0)
    model.engine.profile(True)
    t, _ = timed(lambda: att.attack(x, y), reps=1)
    prof = model.engine.profile_read()
    model.engine.profile(False)
    tot = sum(v[0] for v in prof.values())
    print(json.dumps({"config": f"5: PGD-{iters} vs iv_plda SV, B={B}, 5 s, C=2048 D=400",
                      "precision": os.environ.get("SGB200_IV_PRECISION", "tf32x3"), "s_per_iteration": t / (iters + 1),
                      "utt_iter_per_s": B * (iters + 1) / t, "params_s": t1 - t0, "load_s": t2 - t1,
                      "profile_ms": {k: round(v[0], 2) for k, v in prof.items() if v[1]},
                      "profile_total_ms": round(tot, 2)}))


if __name__ == "__main__":
    if os.environ.get("SGB200_ONLY_CFG5"):
        config5()
        sys.exit(0)
    main()
