"""Capture the Procrustes X^T matrices of a real step and A/B the two Jacobi paths on them."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import basd_b200.synthetic as syn
from basd_b200 import _engine as eng
from basd_b200._native import call, ptr, stream
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
captured = {}
orig = eng.jacobi_rows
def spy(g, dims=None, sweeps_out=None):
    captured["g"] = g.clone()
    return orig(g, dims=dims, sweeps_out=sweeps_out)
eng.jacobi_rows = spy
pro = eng.procrustes_forward(students, teachers, stats, weights, work.n_student, True)
torch.cuda.synchronize()
g0 = captured["g"]
p, n, _ = g0.shape
print("captured", g0.shape, "finite", bool(torch.isfinite(g0).all()), "absmax", float(g0.abs().max()))
ref = torch.linalg.svdvals(g0.double())
print("cond (first problem): smax", float(ref[0, 0]), "s[-5:]", ref[0, -5:].tolist())
for legacy in (False, True):
    if legacy:
        os.environ["BASD_JACOBI_LEGACY"] = "1"
    else:
        os.environ.pop("BASD_JACOBI_LEGACY", None)
    work_g = g0.clone()
    sweeps = torch.zeros(p, dtype=torch.int32, device=dev)
    call("basd_jacobi_rows", ptr(work_g), n, n, n, n * n, p, None, 1e-6, 18, ptr(sweeps), stream())
    torch.cuda.synchronize()
    w = work_g.double()
    gram0 = g0.double().transpose(1, 2) @ g0.double()
    gram1 = w.transpose(1, 2) @ w
    inv = (gram1 - gram0).abs().amax(dim=(1, 2)) / gram0.abs().amax(dim=(1, 2))
    rr = w @ w.transpose(1, 2)
    nrm = rr.diagonal(dim1=1, dim2=2).clamp(min=0).sqrt()
    cosm = rr / (nrm.unsqueeze(1) * nrm.unsqueeze(2)).clamp(min=1e-300)
    big = nrm > 1e-5 * nrm.max(dim=1, keepdim=True).values
    mask = big.unsqueeze(1) & big.unsqueeze(2) & ~torch.eye(n, dtype=torch.bool, device=dev)
    off = (cosm.abs() * mask).amax(dim=(1, 2))
    sv = nrm.sort(dim=1, descending=True).values
    sverr = (sv - ref).abs().max(dim=1).values / ref[:, 0]
    print(f"legacy={legacy}: invariance max {float(inv.max()):.2e} off-cos max {float(off.max()):.2e} "
          f"sv err max {float(sverr.max()):.2e} sweeps {sweeps.tolist()}")
    print("   nan rows:", int((~torch.isfinite(work_g)).sum()), " sum sv", float(sv[0].sum()), "ref", float(ref[0].sum()))
    # unmasked check, per problem: any nonzero row counts
    nz = nrm > 0
    mask2 = nz.unsqueeze(1) & nz.unsqueeze(2) & ~torch.eye(n, dtype=torch.bool, device=dev)
    off2 = (cosm.abs() * mask2).amax(dim=(1, 2))
    small = [(nrm[i] / nrm[i].max()).sort().values[:4].tolist() for i in range(p)]
    print("   unmasked off-cos per problem:", [f"{v:.1e}" for v in off2.tolist()])
    print("   smallest relative row norms, problem 2:", small[2], " problem 3:", small[3])
    i = 2
    worst = (cosm[i].abs() * mask2[i]).flatten().argmax()
    r, c = int(worst // n), int(worst % n)
    print(f"   problem 2 worst pair rows ({r},{c}) rel norms {float(nrm[i, r] / nrm[i].max()):.2e} {float(nrm[i, c] / nrm[i].max()):.2e} cos {float(cosm[i, r, c]):.3e}")
