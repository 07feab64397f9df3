"""GPU-box debugging aid: runs the CUDA phases one by one and prints how far each
intermediate is from the CPU kernel model (oracle/kernel_model.py)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import basd_b200.synthetic as syn
from basd_b200 import _engine as eng
from oracle import kernel_model as km
from tests import _cases as cs


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp(min=1e-300))


def main(key="c1", batch=16, seed=5, temps=(0.2, 0.6, 1.0, 1.4), work=None, uniform_attn=None):
    work = work if work is not None else cs.workload(key, batch)
    kw = {} if uniform_attn is None else dict(uniform_attn=uniform_attn)
    logits, targets, st, te, at = syn.make_inputs(work, seed=seed, **kw)
    proj_s, proj_t, logt0 = cs.selector_state(work)
    logt = torch.tensor(temps[:work.num_points]) if temps else logt0
    layers = sorted(st)
    model = km.full_step_model(logits, targets, st, te, at, layers=layers, proj_s=proj_s, proj_t=proj_t,
                               log_temps=logt, n_student=work.n_student, has_cls=work.has_cls,
                               criterion=cs.criterion(work))
    dev = "cuda"
    students = [st[l].to(dev).contiguous() for l in layers]
    teachers = [te[k].to(dev).contiguous() for k in sorted(te)]
    attns = [at[k].to(dev).contiguous() for k in sorted(at)]
    stats, _ = eng.statistics(students, teachers, attns, work.has_cls)
    b, n_s, _ = students[0].shape
    n_t = teachers[0].shape[1]
    sel = eng.selector_forward(stats, b * n_s, b * n_t, proj_s.to(dev), proj_t.to(dev), logt.to(dev))
    print("ranks", sel.ranks.tolist(), "model", model["ranks"])
    print("dist rel", rel(sel.dist, model["dist"]), "weights maxabs",
          float((sel.weights.cpu() - model["weights"]).abs().max()))
    print("kxk sweeps", sel.sweeps["kxk"].tolist())
    msel = model["sel"]
    for i in range(len(layers)):
        print(f" student {i}: lam rel", rel(sel.lam_s[i], msel["saved"][i]["lam"]),
              "min gap/lam (model)", float(((msel['saved'][i]['lam'][:-1] - msel['saved'][i]['lam'][1:]) / msel['saved'][i]['lam'][:-1]).min()))
        v = sel.vt_s[i].cpu().T
        vm = msel["saved"][i]["vec"]
        print("   |V^T V_model| diag min", float((v.T @ vm).diagonal().abs().min()),
              "orth err", float((v.T @ v - torch.eye(v.shape[0])).abs().max()))
    pro = eng.procrustes_forward(students, teachers, stats, sel.weights, work.n_student, True)
    print("geo", float(pro.geo), "model", float(model["geo"]), "procrustes sweeps max", int(pro.sweeps.max()),
          "mean", float(pro.sweeps.float().mean()), "hist", torch.bincount(pro.sweeps.cpu()).tolist())
    # backward with dL/dgeo = share_geo from the model
    ce, geo = model["ce"], model["geo"]
    share = (1 / geo) / (1 / ce + 1 / geo)
    go = torch.tensor(float(share), device=dev)
    gdir, dw, _ = eng.procrustes_backward(students, teachers, stats, pro, go, work.n_student)
    print("d_weights rel", rel(dw, model["d_weights"]))
    print("  cuda ", dw.flatten()[:6].tolist())
    print("  model", model["d_weights"].flatten()[:6].tolist())
    for i, l in enumerate(layers):
        print(f" direct grad layer {l}: cos", cs.cosine(gdir[i].float().cpu(), model["grad_direct"][l]),
              "rel", rel(gdir[i].float(), model["grad_direct"][l]))
    gsel, dlogt = eng.selector_backward(students, sel, proj_s.to(dev), logt.to(dev), dw, None, 1)
    print("d_logt", dlogt.tolist(), "model", model["grad_log_temps"].tolist())
    for i, l in enumerate(layers):
        msel_grad = model["grad_students"][l] - model["grad_direct"][l]
        print(f" selector grad layer {l}: cos", cs.cosine(gsel[i].float().cpu(), msel_grad),
              "norm", float(gsel[i].float().norm()), "model", float(msel_grad.norm()),
              "| d_dist rel", rel(torch.zeros(1), torch.ones(1)))
    # finer: recompute pieces of selector backward
    e, l_ = sel.weights.shape
    d_s = proj_s.shape[0]
    d_dist = torch.empty(e, l_, device=dev); d_logt2 = torch.empty(e, device=dev)
    from basd_b200._native import call, ptr, stream
    call("basd_mix_weights_bwd", ptr(dw), ptr(sel.weights), ptr(sel.dist), ptr(logt.to(dev)), e, l_, 1.0,
         ptr(d_dist), ptr(d_logt2), stream())
    for i in range(e):
        print(f" d_dist[{i}] rel", rel(d_dist[i], msel["saved"][i]["d_dist"]))
    uxt = sel.uxt.clone()
    call("basd_scale_rows_dsigma", ptr(uxt), ptr(sel.sig), ptr(sel.lam_t), ptr(sel.ranks), ptr(d_dist), d_s, e, l_, stream())
    dd = d_s * d_s
    t1 = torch.empty(e * l_, d_s, d_s, device=dev)
    eng.sgemm(0, 1, d_s, d_s, d_s, sel.wfull, d_s, dd, sel.vxt, d_s, dd, t1, d_s, dd, e * l_)
    blk = torch.empty(e * l_, d_s, d_s, device=dev)
    eng.sgemm(0, 0, d_s, d_s, d_s, t1, d_s, dd, uxt, d_s, dd, blk, d_s, dd, e * l_)
    omega = torch.empty(e, d_s, d_s, device=dev)
    call("basd_omega_accumulate", ptr(blk), ptr(sel.lam_s), ptr(sel.ranks), d_s, e, l_, ptr(omega), stream())
    for i in range(e):
        # compare in a gauge-free way: V Omega V^T
        vm = msel["saved"][i]["vec"]
        ref = vm @ msel["saved"][i]["omega"] @ vm.T
        v = sel.vt_s[i].cpu().T
        got = v @ omega[i].cpu() @ v.T
        print(f" dGram[{i}] rel", rel(got, ref), "omega absmax", float(omega[i].abs().max()),
              "model", float(msel['saved'][i]['omega'].abs().max()))
        # per teacher layer block check
        for j in range(l_):
            k = int(sel.ranks[j])
            p = msel["saved"][i]["per"][j]
            bm = (p["full"][k:] @ p["vx"])  # (D-k, k) before dsig
            if j < 2:
                sig_c = sel.sig[i * l_ + j, :k].cpu()
                print(f"    pair ({i},{j}) k={k} sig rel", rel(sig_c, p["sig"]), "sig max", float(sig_c.max()),
                      "min", float(sig_c.min()), "blk absmax", float(blk[i * l_ + j].abs().max()))


if __name__ == "__main__":
    args = sys.argv[1:]
    if args:
        main(args[0], int(args[1]), int(args[2]), None if len(args) < 4 else tuple(float(x) for x in args[3].split(",")))
    else:
        main()


def module_path(key="c1", batch=16, seed=5, temps=(0.2, 0.6, 1.0, 1.4)):
    work = cs.workload(key, batch)
    inputs = syn.make_inputs(work, seed=seed)
    logits, targets, st, te, at = inputs
    proj_s, proj_t, logt0 = cs.selector_state(work)
    logt = torch.tensor(temps[:work.num_points]) if temps else logt0
    layers = sorted(st)
    model = km.full_step_model(logits, targets, st, te, at, layers=layers, proj_s=proj_s, proj_t=proj_t,
                               log_temps=logt, n_student=work.n_student, has_cls=work.has_cls,
                               criterion=cs.criterion(work))
    for trial in range(2):
        got = cs.run_cuda(work, inputs, list(temps) if temps else None)
        print(f"[module trial {trial}] loss", float(got["loss"]), "model", float(model["loss"]))
        print("  grad_log_temps", got["grad_log_temps"].tolist(), "model", model["grad_log_temps"].tolist())
        for l in layers:
            g = got["grad_students"][l]
            print(f"  layer {l}: cos(total)", cs.cosine(g, model["grad_students"][l]),
                  "norm", float(g.norm()), "model", float(model["grad_students"][l].norm()),
                  "cos(g - direct_model, selector_model)",
                  cs.cosine(g - model["grad_direct"][l], model["grad_students"][l] - model["grad_direct"][l]))


if __name__ == "__main__" and len(sys.argv) > 1:
    args = sys.argv[1:]
    module_path(args[0], int(args[1]), int(args[2]), None if len(args) < 4 else tuple(float(x) for x in args[3].split(",")))
