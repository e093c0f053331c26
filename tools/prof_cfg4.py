"""Per-category device time of the fused EOT-PGD-through-FeCo loop (BASELINE configs[3]: B = 256, EOT 50, k-means 0.5)."""
import sys, tempfile, torch, json
sys.path.insert(0, '.')
sys.path.insert(0, 'tools')
from bench_configs import xv_model
from speakerguard_b200 import _lib
from speakerguard_b200.engine import make_loss_params
from speakerguard_b200.synthetic import synthetic_batch
B, E, Eb, iters = int(sys.argv[1]) if len(sys.argv) > 1 else 256, 50, int(sys.argv[2]) if len(sys.argv) > 2 else 5, 2
model = xv_model("bf16")
eng = model.engine
x, y = synthetic_batch(B, 48000)
x = x[:, 0].cuda().contiguous(); y = y.cuda()
lp = make_loss_params("Entropy")
def run():
    xa = x.clone()
    eng.pgd_run(xa, x, y, max_iter=iters, epsilon=0.002, step_size=0.0004, lp=lp, dither_mode=_lib.DITHER_PHILOX, seed=1,
                eot_size=E, eot_batch=Eb, feco_ratio=0.5, grad_sign=1.0)
    torch.cuda.synchronize()
run()
eng.profile(True)
run()
prof = eng.profile_read()
eng.profile(False)
tot = sum(v[0] if isinstance(v, (list, tuple)) else v for v in prof.values())
print(json.dumps({"B": B, "eot_batch": Eb, "rows_per_pass": B * Eb, "ms_per_iteration": tot / iters, "per_category_ms_per_iteration": {k: round((v[0] if isinstance(v, (list, tuple)) else v) / iters, 3) for k, v in prof.items() if (v[0] if isinstance(v, (list, tuple)) else v) > 0}}))
