"""GPU-box debugging aid: accuracy of the selector's statistics and eigen-decompositions at full size
against fp64 computed from the same tokens.   python tests/tools/debug_selector_stats.py [batch] [seed]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import basd_b200.synthetic as syn
from basd_b200 import _engine as eng
from basd_b200._native import call, ptr, stream
from tests import _cases as cs

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 7
if os.environ.get("NO_COMPLETE"):
    eng.complete_null_space = lambda vt: None
dev = "cuda"
work = syn.scaled(syn.WORKLOADS["c2"], batch)
logits, targets, st, te, at = syn.make_inputs_fast(work, seed=seed, device=dev)
proj_s, proj_t, _ = cs.selector_state(work)
proj_s, proj_t = proj_s.to(dev), proj_t.to(dev)
layers = sorted(st)
students = [st[l].contiguous() for l in layers]
teachers = [te[k].contiguous() for k in sorted(te)]
attns = [at[k].contiguous() for k in sorted(at)]
stats, _ = eng.statistics(students, teachers, attns, work.has_cls)
b, n_s, d_s = students[0].shape
m = b * n_s
sel = eng.selector_forward(stats, m, m, proj_s, proj_t, torch.tensor([0.3, 0.6, 0.9, 1.2], device=dev))
ranks = sorted(set(sel.ranks.tolist()))
# the matrices the eigen-solver was actually given (same calls as selector_forward)
e_ = len(students)
k32 = torch.empty(e_, d_s, d_s, device=dev)
c32 = torch.empty(e_, d_s, device=dev)
from basd_b200 import _native as nat
ws = torch.empty(nat.load().basd_rotate_stats_f64_workspace_bytes(d_s, d_s, e_) // 8, dtype=torch.float64, device=dev)
call("basd_rotate_stats_f64", ptr(proj_s), d_s, d_s, ptr(stats.gram_s), ptr(stats.col_s), e_, 1.0 / m, ptr(ws),
     ptr(k32), ptr(c32), stream())
torch.cuda.synchronize()
print("ranks", sel.ranks.tolist())
for i, s in enumerate(students):
    x = s.double().reshape(-1, d_s)
    g64 = x.T @ x
    print(f"student {i}: token Gram rel err {float((stats.gram_s[i].double() - g64).norm() / g64.norm()):.2e} "
          f"max abs / max {float((stats.gram_s[i].double() - g64).abs().max() / g64.abs().max()):.2e}; "
          f"colsum rel {float((stats.col_s[i].double() - x.sum(0)).norm() / x.sum(0).norm()):.2e}")
    z = x @ proj_s.double().T
    z = z - z.mean(0, keepdim=True)
    k64 = z.T @ z / m
    lam64, v64 = torch.linalg.eigh(k64)
    lam64, v64 = lam64.flip(0), v64.flip(1)
    lam = sel.lam_s[i].double() / m
    vt = sel.vt_s[i].double()                       # rows = eigenvectors
    norms = vt.norm(dim=1)
    print(f"   zero rows: {int((norms < 0.5).sum())}; first zero row index {int((norms < 0.5).float().argmax()) if (norms < 0.5).any() else -1}; "
          f"lam tail {[float(x) for x in lam[-4:]]} lam64 tail {[float(x) for x in lam64[-4:]]}")
    # the Gram the kernels decomposed: rebuild from their own factors
    krec = vt.T @ torch.diag(lam) @ vt
    print(f"   K (from V lam V^T) vs fp64 K: rel {float((krec - k64).norm() / k64.norm()):.2e}, in units of the "
          f"boundary gap: |dK|_2 / gap = {float(torch.linalg.matrix_norm(krec - k64, ord=2) / (lam64[ranks[0]-1] - lam64[ranks[0]])):.3f}")
    print(f"   lam rel err (top) {float(((lam - lam64) / lam64)[:8].abs().max()):.2e} at boundary "
          f"{[float((lam[k-1]-lam64[k-1])/lam64[k-1]) for k in ranks]} lam64[k-1],[k]: "
          f"{[(float(lam64[k-1]), float(lam64[k])) for k in ranks]}")
    print(f"   V orthogonality {float((vt @ vt.T - torch.eye(d_s, device=dev, dtype=torch.float64)).abs().max()):.2e}; "
          f"residual |K64 v - lam v| / lam_max max {float(((k64 @ vt.T) - vt.T * lam).norm(dim=0).max() / lam64[0]):.2e}")
    kk = k32[i].double()
    scale = float((kk * k64).sum() / (k64 * k64).sum())           # the kernels' normalisation of K vs 1/M
    kk = kk / scale
    lam32, v32 = torch.linalg.eigh(kk)
    lam32, v32 = lam32.flip(0), v32.flip(1)
    gap = float(lam64[ranks[0] - 1] - lam64[ranks[0]])
    print(f"   K given to the solver vs fp64 K: scale {scale:.4g}, |dK|_2 {float(torch.linalg.matrix_norm(kk - k64, ord=2)):.3e} "
          f"(boundary gap {gap:.3e}, lam_min64 {float(lam64[-1]):.3e}); eigenvalues of the given K: min {float(lam32[-1]):.3e} "
          f"negative count {int((lam32 < 0).sum())}")
    mu = z.new_tensor(0.0)
    for k in ranks:
        s_stat = torch.linalg.svdvals(v32[:, :k].T @ v64[:, :k])
        s_solv = torch.linalg.svdvals(vt[:k] @ v32[:, :k])
        print(f"   k={k}: sin(angle) statistics error (eigh64 of given K vs fp64 K) {float((1 - s_stat.min() ** 2).clamp(min=0).sqrt()):.4f}; "
              f"solver error (kernel V vs eigh64 of given K) {float((1 - s_solv.min() ** 2).clamp(min=0).sqrt()):.4f}")
    for k in ranks:
        sv = torch.linalg.svdvals(vt[:k] @ v64[:, :k])
        print(f"   k={k}: sin(largest principal angle between top-k subspaces) {float((1 - sv.min() ** 2).clamp(min=0).sqrt()):.4f}")
