import sys, torch
sys.path.insert(0, '.')
from oracle import sg_oracle as O
from speakerguard_b200.engine import Engine
prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 200
e = Engine("cuda:0", precision=prec); e.load_xv(O.make_xv_params(seed=0))
feat = torch.randn(3, T, 32, device="cuda")
emb, ws = e.embed_fwd(feat)
torch.cuda.synchronize()
print("fwd ok", emb.abs().max().item())
d = e.embed_bwd(torch.randn_like(emb), ws, 3, T)
torch.cuda.synchronize()
print("bwd ok", d.abs().max().item())
