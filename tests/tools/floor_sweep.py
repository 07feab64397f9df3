"""GPU-box aid: sweep the Procrustes singular-value floor (Gram sides) and report the parity margins
on the synthetic workloads, the random-init backbone features and ill-conditioned per-sample tokens
(backbone tokens rotated, given a geometric spectrum and two "massive activation" columns).
Usage: python tests/tools/floor_sweep.py [floor ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import basd_b200.synthetic as syn
from basd_b200 import _engine as eng
from basd_b200 import backbone_features as bf
from basd_b200.losses import geometric_relational_loss
from oracle import ref_port as rp
from tests import _cases as cs


def distort(x, scale, decay):
    d = x.shape[-1]
    q, _ = torch.linalg.qr(torch.randn(d, d, generator=torch.Generator().manual_seed(d)))
    y = (x.cpu().float() @ q) * torch.logspace(0, -decay, d)
    y[..., 7] *= scale
    y[..., 100] *= scale
    return y


def procrustes_case(s, t, attn, has_cls=True):
    sg = s.clone().requires_grad_(True)
    ref = rp.procrustes_loss(sg, t, attn, has_cls)
    ref.backward()
    sd = s.cuda().requires_grad_(True)
    got = geometric_relational_loss(sd, t.cuda(), attn.cuda(), has_cls_token=has_cls)
    got.backward()
    return abs(float(got) - float(ref)) / abs(float(ref)), cs.cosine(sd.grad.cpu(), sg.grad)


def full_case(work, inputs):
    ref = cs.run_oracle(work, inputs)
    got = cs.run_cuda(work, inputs)
    rel = abs(float(got["loss"]) - float(ref["loss"])) / abs(float(ref["loss"]))
    cos = min(cs.cosine(got["grad_students"][l], ref["grad_students"][l]) for l in ref["layers"])
    return rel, cos


if __name__ == "__main__":
    floors = [float(a) for a in sys.argv[1:]] or [2.5e-4, 1e-4, 5e-5, 2e-5]
    layers = [0, 4, 7, 11]
    bb = bf.backbone_inputs("deit_small", "deit_base", layers, 4, seed=0, device="cuda")
    bb = tuple(x.cpu() if torch.is_tensor(x) else {k: v.cpu() for k, v in x.items()} for x in bb)
    c3 = bf.backbone_inputs("deit_small", "resnet50", layers, 4, seed=0, device="cuda")
    c3 = tuple(x.cpu() if torch.is_tensor(x) else {k: v.cpu() for k, v in x.items()} for x in c3)
    for fl in floors:
        eng.PROC_SV_FLOOR = fl
        out = [f"floor {fl:.1e}:"]
        for key, batch, seed in (("c2", 8, 2), ("c3", 8, 3)):
            work = cs.workload(key, batch)
            rel, cos = full_case(work, syn.make_inputs(work, seed=seed))
            out.append(f"{key} spectral {rel:.1e}/{cos:.6f}")
        rel, cos = full_case(cs.workload("c2", 4), bb)
        out.append(f"c2 backbone {rel:.1e}/{cos:.6f}")
        rel, cos = full_case(cs.workload("c3", 4), c3)
        out.append(f"c3 backbone {rel:.1e}/{cos:.6f}")
        for scale, decay in ((10, 1), (50, 2), (100, 3)):
            rel, cos = procrustes_case(distort(bb[2][4], scale, decay), distort(bb[3][6], scale, decay), bb[4][6])
            out.append(f"illcond({scale},{decay}) {rel:.1e}/{cos:.6f}")
        print("  ".join(out), flush=True)
