"""Time the FeCo k-means kernel alone (B utterances x n frames x 30 dims, k = n / 2) and print clustering statistics."""
import sys, time, torch
sys.path.insert(0, '.')
from oracle import sg_oracle as O
from speakerguard_b200.engine import Engine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1600
n = int(sys.argv[2]) if len(sys.argv) > 2 else 300
e = Engine("cuda:0", precision="bf16"); e.load_xv(O.make_xv_params(seed=0))
x = ((torch.rand(min(B, 64), 48000) * 2 - 1) * 0.5).cuda()
from speakerguard_b200 import _lib
raw = e.mfcc_fwd(x, _lib.DITHER_PHILOX, None, seed=1, pass_=0, ld=32)[:, :n, :30].contiguous()
feat = raw.repeat((B + raw.shape[0] - 1) // raw.shape[0], 1, 1)[:B].contiguous()
k = n // 2
for _ in range(2):
    ids = e.feco_kmeans(feat, k, seed=5)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(5):
    ids = e.feco_kmeans(feat, k, seed=7 + i)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
out, counts = e.feco_means_fwd(feat, ids, k, True)
cen = out
inertia = float(((feat - torch.gather(cen, 1, ids.long().unsqueeze(-1).expand(-1, -1, 30))) ** 2).sum((1, 2)).mean())
print(f"feco_kmeans B={B} n={n} k={k}: {dt*1e3:.3f} ms per launch, {dt/B*1e6:.2f} us per utterance; mean inertia {inertia:.1f}, empty clusters {float((counts==0).float().mean()):.4f}")
