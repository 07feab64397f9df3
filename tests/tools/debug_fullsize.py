"""GPU-box debugging aid: C2 at full size (B = 256) -- where does the student-gradient error against the
oracle come from?  Splits the CUDA gradient into its direct (Procrustes) and selector parts and compares
per layer and per sample.   python tests/tools/debug_fullsize.py [batch] [seed]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import basd_b200.synthetic as syn
from basd_b200 import _engine as eng
from tests import _cases as cs

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 7
dev = "cuda"
work = syn.scaled(syn.WORKLOADS["c2"], batch)
temps = [0.3, 0.6, 0.9, 1.2]
dev_inputs = syn.make_inputs_fast(work, seed=seed, device=dev)
inputs = tuple(x.cpu() if not isinstance(x, dict) else {k: v.cpu() for k, v in x.items()} for x in dev_inputs)
del dev_inputs
ref = cs.run_oracle(work, inputs, temps)
logits, targets, st, te, at = inputs
proj_s, proj_t, _ = cs.selector_state(work)
logt = torch.tensor(temps)
layers = sorted(st)
students = [st[l].to(dev).contiguous() for l in layers]
teachers = [te[k].to(dev).contiguous() for k in sorted(te)]
attns = [at[k].to(dev).contiguous() for k in sorted(at)]
stats, _ = eng.statistics(students, teachers, attns, work.has_cls)
b, n_s, _ = students[0].shape
sel = eng.selector_forward(stats, b * n_s, b * n_s, proj_s.to(dev), proj_t.to(dev), logt.to(dev))
print("ranks equal", sel.ranks.tolist() == ref["ranks"], "weights maxabs", float((sel.weights.cpu() - ref["weights"]).abs().max()))
pro = eng.procrustes_forward(students, teachers, stats, sel.weights, work.n_student, True)
ce, geo = float(ref["ce"]), float(ref["geo"])
share = (1 / geo) / (1 / ce + 1 / geo)
print("geo", float(pro.geo), geo, "share", share, "sweeps mean", float(pro.sweeps.float().mean()), "max", int(pro.sweeps.max()))
go = torch.tensor(share, device=dev)
gdir, dw, _ = eng.procrustes_backward(students, teachers, stats, pro, go, work.n_student)
gsel, dlogt = eng.selector_backward(students, sel, proj_s.to(dev), logt.to(dev), dw, None, 1)
print("d_logt", dlogt.tolist(), "oracle", ref["grad_log_temps"].tolist())
for i, l in enumerate(layers):
    r = ref["grad_students"][l]
    d = gdir[i].float().cpu()
    s = gsel[i].float().cpu()
    tot = d + s
    resid = r - d                                   # the oracle's selector share if the direct part is right
    print(f"layer {l}: cos(total) {cs.cosine(tot, r):.6f}  |ref| {float(r.norm()):.4e} |direct| {float(d.norm()):.4e} "
          f"|sel| {float(s.norm()):.4e} |ref-direct| {float(resid.norm()):.4e}  cos(sel, ref-direct) {cs.cosine(s, resid):.5f}  "
          f"cos(direct, ref) {cs.cosine(d, r):.6f}")
    per = torch.nn.functional.cosine_similarity(tot.flatten(1).double(), r.flatten(1).double(), dim=1)
    worst = per.argsort()[:5]
    print("   per-sample cos: min", float(per.min()), "median", float(per.median()), "worst samples", worst.tolist(),
          [round(float(per[j]), 5) for j in worst])
    # error energy: along the selector direction or not
    err = tot - r
    print("   |err|/|ref|", float(err.norm() / r.norm()), " err projected on sel:", float((err.flatten() @ s.flatten()) / (s.norm() * err.norm())))
    # column-space view: error per feature dimension (selector acts as Z_c W': a right-multiplication)
    ecol = err.flatten(0, 1).norm(dim=0)
    print("   error by feature column: top", ecol.topk(5).values.tolist(), "median", float(ecol.median()))
