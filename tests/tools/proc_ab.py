"""Run procrustes_forward with both Jacobi paths in one process and diff the outputs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import basd_b200.synthetic as syn
from basd_b200 import _engine as eng
from tests import _cases as cs

work = cs.workload("c2", 4)
logits, targets, st, te, at = syn.make_inputs(work, seed=0)
dev = "cuda"
layers = sorted(st)
students = [st[l].to(dev).contiguous() for l in layers]
teachers = [te[k].to(dev).contiguous() for k in sorted(te)]
attns = [at[k].to(dev).contiguous() for k in sorted(at)]
stats = eng.attention_only_stats(teachers, attns, True)
e, l = len(students), len(teachers)
weights = torch.full((e, l), 1.0 / l, device=dev)
res = {}
for name in ("grouped", "legacy", "grouped2"):
    if name == "legacy":
        os.environ["BASD_JACOBI_LEGACY"] = "1"
    else:
        os.environ.pop("BASD_JACOBI_LEGACY", None)
    cap = {}
    orig_fin = eng.call
    pro = eng.procrustes_forward(students, teachers, stats, weights, work.n_student, True)
    torch.cuda.synchronize()
    res[name] = dict(f=pro.f.clone(), gw=pro.gw.clone(), m_a=pro.m_a.clone(), m_b=pro.m_b.clone(),
                     sweeps=pro.sweeps.clone())
    print(name, "geo", float(pro.geo), "f[:4]", pro.f[:4].tolist(), "sweeps", pro.sweeps.tolist())
for a, b in (("grouped", "legacy"), ("grouped", "grouped2")):
    for k in ("f", "gw", "m_a", "m_b"):
        x, y = res[a][k].double(), res[b][k].double()
        print(f"{a} vs {b} {k}: rel diff {float((x - y).norm() / y.norm()):.3e} max abs {float((x - y).abs().max()):.3e}")
    d = (res[a]["m_a"] - res[b]["m_a"]).flatten(1).norm(dim=1) / res[b]["m_a"].flatten(1).norm(dim=1)
    print("   per-problem m_a rel diff:", [f"{v:.2e}" for v in d.tolist()])
