import sys, tempfile, torch
sys.path.insert(0, '.')
from speakerguard_b200.attack.PGD import PGD
from speakerguard_b200.defense.feature_level import FeCo, FeCoDefense
from speakerguard_b200.model.defended_model import defended_model
from speakerguard_b200.model.xv_plda import xv_plda
from speakerguard_b200.synthetic import make_xv_params, state_dict_of, write_xv_model_files, synthetic_batch
p = make_xv_params(0)
f = write_xv_model_files(p, tempfile.mkdtemp())
base = xv_plda(state_dict_of(p), f["plda.txt"], f["mean.vec"], f["transform.txt"], model_file=f["speaker_model"], device="cuda:0", dither="philox")
dm_fused = defended_model(base, defense=[[1, FeCoDefense("kmeans", 0.5, "L2")]], order="sequential")
dm_generic = defended_model(base, defense=[[1, lambda feat: FeCo(feat, "kmeans", 0.5, "L2")]], order="sequential")
x, y = synthetic_batch(8, 32000)
x, y = x.cuda(), y.cuda()
def ce(dm, xx):
    with torch.no_grad():
        s = torch.stack([dm.score(xx) for _ in range(8)]).mean(0)
    return torch.nn.functional.cross_entropy(s, y, reduction='none'), s
l0, s0 = ce(dm_generic, x)
print("labels", y.tolist(), "clean CE", l0.tolist())
for name, dm in (("fused", dm_fused), ("generic", dm_generic), ("undefended-fused", base)):
    for E in (1, 4):
        att = PGD(dm, epsilon=0.002, step_size=0.0004, max_iter=20, batch_size=8, EOT_size=E, EOT_batch_size=E, verbose=0)
        adv, success = att.attack(x, y)
        l1, s1 = ce(dm_generic, adv)
        print(name, "E", E, "CE after (defended eval)", [round(v, 3) for v in l1.tolist()], "mean gain", float((l1 - l0).mean()), "succ", sum(success))
