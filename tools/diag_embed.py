import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import sg_oracle as O
from speakerguard_b200.engine import Engine
p = O.make_xv_params(seed=0)
eng = Engine("cuda:0"); eng.load_xv(p)
for (B, T) in [(1, 200), (2, 200), (3, 200), (3, 300), (1, 111), (2, 111), (1, 500), (2, 300), (4,200)]:
    g = torch.Generator().manual_seed(T)
    feat = (torch.randn(B, T, 30, generator=g) * 3).requires_grad_(True)
    emb_ref = O.process_emb(O.xvector(feat, p), p)
    w = torch.randn(B, 200, generator=g)
    (emb_ref * w).sum().backward()
    f32 = torch.zeros(B, T, 32); f32[:, :, :30] = feat.detach()
    emb, ws = eng.embed_fwd(f32.cuda())
    dfeat = eng.embed_bwd(w.cuda(), ws, B, T).cpu()[:, :, :30]
    err = (dfeat - feat.grad).abs()
    per_bt = err.max(2)[0]                      # [B,T]
    scale = feat.grad.abs().max()
    print(f"B={B} T={T}: max rel {float(err.max()/scale):.3e}")
    for b in range(B):
        bad = (per_bt[b] > 1e-4 * scale).nonzero().flatten().tolist()
        print(f"   utt {b}: bad frames ({len(bad)}): {bad[:12]} ... {bad[-6:]}  maxerr/scale {float(per_bt[b].max()/scale):.3e}")
