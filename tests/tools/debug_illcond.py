"""GPU-box debugging aid: the stage-by-stage comparison of debug_stages.py on ill-conditioned tokens
(features of random-init backbones, rotated, geometric spectrum, two massive columns)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import basd_b200.synthetic as syn
from basd_b200 import backbone_features as bf
from oracle import ref_port as rp
from tests import _cases as cs
from tests.tools import debug_stages

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 50
decay = float(sys.argv[2]) if len(sys.argv) > 2 else 2
work = cs.workload("c2", 4)
work.token_dtype = torch.float32
layers = rp.extraction_layers(work.student_depth, work.num_points)
logits, targets, st, te, at = bf.backbone_inputs("deit_small", "deit_base", layers, 4, seed=0, device="cuda")


def distort(x):
    d = x.shape[-1]
    q, _ = torch.linalg.qr(torch.randn(d, d, generator=torch.Generator().manual_seed(d)))
    y = (x.cpu().float() @ q) * torch.logspace(0, -decay, d)
    y[..., 7] *= scale
    y[..., 100] *= scale
    return y


inputs = (logits.cpu(), targets.cpu(), {k: distort(v) for k, v in st.items()},
          {k: distort(v) for k, v in te.items()}, {k: v.cpu() for k, v in at.items()})
syn.make_inputs = lambda *a, **k: inputs
debug_stages.main(work=work, temps=None)

# hypothesis: dL/dweights = <z, T_l> is polluted by the (exactly zero) column sums of z times the large
# column means of T_l; removing z's column means must restore it
from basd_b200 import _engine as eng
from oracle import kernel_model as km
dev = "cuda"
logits, targets, st_, te_, at_ = inputs
proj_s, proj_t, logt = cs.selector_state(work)
layers = sorted(st_)
model = km.full_step_model(logits, targets, st_, te_, at_, layers=layers, proj_s=proj_s, proj_t=proj_t,
                           log_temps=logt, n_student=work.n_student, has_cls=work.has_cls,
                           criterion=cs.criterion(work))
students = [st_[l].to(dev).contiguous() for l in layers]
teachers = [te_[k].to(dev).contiguous() for k in sorted(te_)]
attns = [at_[k].to(dev).contiguous() for k in sorted(at_)]
stats, _ = eng.statistics(students, teachers, attns, work.has_cls)
b, n_s, _ = students[0].shape
sel = eng.selector_forward(stats, b * n_s, b * n_s, proj_s.to(dev), proj_t.to(dev), logt.to(dev))
pro = eng.procrustes_forward(students, teachers, stats, sel.weights, work.n_student, True)
ce, geo = model["ce"], model["geo"]
share = (1 / geo) / (1 / ce + 1 / geo)
go = torch.tensor(float(share), device=dev)
gdir, dw, z = eng.procrustes_backward(students, teachers, stats, pro, go, work.n_student, want_teacher_grad=True)
zbar = z.double().mean(dim=2)                                   # (E, B, Dt)
print("|z| per entry", float(z.abs().mean()), "|column mean of z|", float(zbar.abs().mean()),
      "max", float(zbar.abs().max()))
tsum = torch.stack([t.double().sum(dim=1) for t in teachers])   # (L, B, Dt)
corr = torch.einsum("ebd,lbd->el", zbar, tsum)
print("dw cuda     ", dw[0].tolist())
print("dw corrected", (dw.double() - corr)[0].tolist())
print("dw model    ", model["d_weights"][0].tolist())

# piece by piece for extraction point 0: teacher-token gradient and importance gradient
t_stack = torch.stack([te_[k] for k in sorted(te_)])
rows = torch.stack([km.attn_rows(at_[k], work.has_cls) for k in sorted(at_)])
w_mix = model["weights"][0]
aligned = km.mix_and_align(w_mix, t_stack, work.n_student)
imp, total = km.mix_importance(w_mix, rows, work.n_student)
gt_m, gw_m, cw_m, f_m = [], [], [], []
for bi in range(b):
    f, a_, b_, c_ = km.procrustes_sample(st_[layers[0]][bi].float(), aligned[bi].float(), imp[bi])
    gt_m.append(b_); gw_m.append((c_ - f) / total[bi]); cw_m.append(c_); f_m.append(f)
gt_m, gw_m = torch.stack(gt_m), torch.stack(gw_m)
scale = float(share) / (len(layers) * b)
zc = z[0].cpu() / scale
print("teacher-token grad: cos", cs.cosine(zc, gt_m), "rel", float((zc - gt_m).norm() / gt_m.norm()))
print("importance grad gw: cos", cs.cosine(pro.gw[0].cpu(), gw_m), "rel",
      float((pro.gw[0].cpu() - gw_m).norm() / gw_m.norm()), "norms", float(pro.gw[0].norm()), float(gw_m.norm()))
print("  f cuda", pro.f[:b].tolist(), "model", [float(x) for x in f_m])
tok_term = torch.stack([(gt_m * t_stack[j].float()).sum() for j in range(t_stack.shape[0])]) * scale
row_term = torch.stack([(gw_m * rows[j]).sum() for j in range(t_stack.shape[0])]) * scale
print("model token term", tok_term[:4].tolist(), "row term", row_term[:4].tolist())
tok_c = torch.stack([(z[0].cpu().double() * t_stack[j].double()).sum() for j in range(t_stack.shape[0])])
row_c = torch.stack([(pro.gw[0].cpu().double() * rows[j].double()).sum() for j in range(t_stack.shape[0])]) * scale
print("cuda  token term", tok_c[:4].tolist(), "row term", row_c[:4].tolist())
print("aligned rel", float((pro.aligned[0].float().cpu() - aligned).norm() / aligned.norm()),
      "w rel", float((pro.w[0].cpu() - imp).norm() / imp.norm()))
