"""GPU-box debugging aid: precision of the selector share of the student gradient at full size.
For fixed dL/dweights (taken from the CUDA path) the selector gradient is  d/dS sum_il dW[i,l] w[i,l](S);
it is computed (a) by the CUDA closed form, (b) by autograd through the reference algorithm in fp32 and
(c) the same in fp64.  If (b) is as far from (c) as (a) is, the discrepancy is conditioning of the
reference's own definition (1 / (s_i^2 - s_j^2) across the rank boundary), not the kernels.
    python tests/tools/debug_selector_grad.py [batch] [seed]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.nn.functional as F
import basd_b200.synthetic as syn
from basd_b200 import _engine as eng
from oracle import ref_port as rp
from tests import _cases as cs

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 7
if len(sys.argv) > 3:
    eng.EIG_STOP_COS = float(sys.argv[3])
print("EIG_STOP_COS", eng.EIG_STOP_COS)
dev = "cuda"
work = syn.scaled(syn.WORKLOADS["c2"], batch)
temps = [0.3, 0.6, 0.9, 1.2]
dev_inputs = syn.make_inputs_fast(work, seed=seed, device=dev)
logits, targets, st, te, at = tuple(x.cpu() if not isinstance(x, dict) else {k: v.cpu() for k, v in x.items()}
                                    for x in dev_inputs)
del dev_inputs
proj_s, proj_t, _ = cs.selector_state(work)
logt = torch.tensor(temps)
layers = sorted(st)
students = [st[l].to(dev).contiguous() for l in layers]
teachers = [te[k].to(dev).contiguous() for k in sorted(te)]
attns = [at[k].to(dev).contiguous() for k in sorted(at)]
stats, _ = eng.statistics(students, teachers, attns, work.has_cls)
b, n_s, _ = students[0].shape
sel = eng.selector_forward(stats, b * n_s, b * n_s, proj_s.to(dev), proj_t.to(dev), logt.to(dev))
print("eig sweeps", sel.sweeps["eig"].tolist())
torch.manual_seed(0)
dw = torch.randn(len(layers), len(teachers), device=dev) * 0.1
gsel, dlogt = eng.selector_backward(students, sel, proj_s.to(dev), logt.to(dev), dw, None, 1)
ranks = sel.ranks.tolist()
lam = sel.lam_s.cpu()
for i in range(len(layers)):
    gaps = [(float(lam[i, k - 1] - lam[i, k]) / float(lam[i, k - 1])) for k in sorted(set(ranks))]
    print(f"student {i}: lam max {float(lam[i,0]):.4e} min {float(lam[i,-1]):.4e}; relative gap at the rank boundaries "
          f"{[round(g, 5) for g in gaps]}; smallest relative gap anywhere {float(((lam[i,:-1]-lam[i,1:])/lam[i,:-1]).min()):.2e}")


def reference_weights(student, dtype):
    """weights[i, :] of one extraction point through the reference algorithm (layer_selector.py:86-108)."""
    ps, pt = proj_s.to(dtype), proj_t.to(dtype)
    out = []
    bases = []
    with torch.no_grad():
        for key in sorted(te):
            z = te[key].to(dtype).reshape(-1, work.d_teacher) @ pt.T
            z = z - z.mean(0, keepdim=True)
            _, s_, vt = torch.linalg.svd(z, full_matrices=False)
            bases.append((vt, s_))
    return bases


t_bases = {}
for dtype in ((torch.float32, torch.float64) if os.environ.get("WITH_FP64") else (torch.float32,)):
    t_bases[dtype] = reference_weights(None, dtype)
    for i, l in enumerate(layers):
        s = st[l].to(dtype).clone().requires_grad_(True)
        ps = proj_s.to(dtype)
        zs = s.reshape(-1, work.d_student) @ ps.T
        zs = zs - zs.mean(0, keepdim=True)
        _, _, vt_s = torch.linalg.svd(zs, full_matrices=False)
        dist = []
        for j, (vt_t, s_t) in enumerate(t_bases[dtype]):
            k = ranks[j]
            cosv = torch.linalg.svdvals(vt_s[:k] @ vt_t[:k].T)
            ang = torch.acos(cosv.clamp(max=1.0 - torch.finfo(torch.float32).eps))
            sw = s_t[:k]
            dist.append((sw * ang.pow(2)).sum() / sw.sum())
        dist = torch.stack(dist)
        w = F.softmax(-dist / F.softplus(logt[i].to(dtype)), dim=0)
        (w * dw[i].cpu().to(dtype)).sum().backward()
        g = s.grad.float()
        if dtype == torch.float32:
            g32 = globals().setdefault("g32", {})
            g32[l] = g
            print(f"layer {l}: fp32 autograd |g| {float(g.norm()):.4e}  cuda |g| {float(gsel[i].float().norm()):.4e}  "
                  f"cos(cuda, fp32 autograd) {cs.cosine(gsel[i].float().cpu(), g):.5f}")
        else:
            print(f"layer {l}: fp64 autograd |g| {float(g.norm()):.4e}  cos(cuda, fp64) {cs.cosine(gsel[i].float().cpu(), g):.5f}  "
                  f"cos(fp32 autograd, fp64) {cs.cosine(g32[l], g):.5f}  weights diff vs cuda "
                  f"{float((w.detach().float() - sel.weights[i].cpu()).abs().max()):.2e}")
