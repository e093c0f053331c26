import sys, time, torch
sys.path.insert(0, '.')
from oracle import sg_oracle as O
from speakerguard_b200.engine import Engine
from speakerguard_b200 import _lib
B, n = 1280, 300
e = Engine("cuda:0", precision="bf16"); e.load_xv(O.make_xv_params(seed=0))
x = ((torch.rand(64, 48000) * 2 - 1) * 0.5).cuda()
raw = e.mfcc_fwd(x, _lib.DITHER_PHILOX, None, seed=1, pass_=0, ld=32)[:, :n, :30].contiguous()
feat = raw.repeat(B // 64, 1, 1).contiguous()
k = n // 2
for mi in (1, 2, 5, 10, 20, 50, 100):
    for _ in range(2): ids = e.feco_kmeans(feat, k, seed=5, max_iter=mi)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(5): ids = e.feco_kmeans(feat, k, seed=7 + i, max_iter=mi)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print(f"max_iter {mi}: {dt*1e3:.3f} ms")
