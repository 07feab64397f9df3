"""Shared helpers: seeded inputs, the oracle run and the CUDA run on identical data."""
import os
import types

import torch

import basd_b200.synthetic as syn
from oracle import ref_port as rp

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SELECTOR_SEED = 1234          # must match tests/golden/make_golden.py
PROBE = 512


def golden(name):
    return torch.load(os.path.join(GOLDEN_DIR, f"{name}.pt"), weights_only=False)


def workload(key, batch):
    return syn.scaled(syn.WORKLOADS[key], batch)


def probe_indices(numel, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, numel, (PROBE,), generator=g)


def fingerprint(tensors):
    return torch.tensor([float(t.detach().double().sum()) for t in tensors], dtype=torch.float64)


def selector_state(work):
    torch.manual_seed(SELECTOR_SEED)
    return rp.make_selector_state(work.num_points, work.d_student, work.d_teacher)


def criterion(work):
    return torch.nn.CrossEntropyLoss(label_smoothing=1.0 / work.num_classes)


def run_oracle(work, inputs, log_temps=None, ranks_override=None):
    """fp32 oracle forward+backward on CPU (bf16 tokens upcast exactly). Returns dict."""
    logits, targets, st, te, at = inputs
    proj_s, proj_t, logt = selector_state(work)
    if log_temps is not None:
        logt = torch.as_tensor(log_temps, dtype=torch.float32).clone()
    logt.requires_grad_(True)
    st32 = {k: v.detach().float().cpu().requires_grad_(True) for k, v in st.items()}
    te32 = {k: v.detach().float().cpu() for k, v in te.items()}
    at32 = {k: v.detach().float().cpu() for k, v in at.items()}
    lg = logits.detach().float().cpu().requires_grad_(True)
    layers = rp.extraction_layers(work.student_depth, work.num_points)
    loss, diag = rp.basd_forward(lg, targets.cpu(), st32, te32, at32, layers=layers, proj_s=proj_s,
                                 proj_t=proj_t, log_temps=logt, n_student_tokens=work.n_student,
                                 has_cls=work.has_cls, criterion=criterion(work),
                                 ranks_override=ranks_override)
    loss.backward()
    return dict(loss=loss.detach(), ce=diag.ce, geo=diag.geo, geo_terms=diag.geo_terms,
                ranks=[diag.ranks[k] for k in sorted(diag.ranks)],
                weights=torch.stack([diag.weights[l] for l in layers]),
                dist=torch.stack([diag.dist[l] for l in layers]),
                grad_students={l: st32[l].grad for l in layers}, grad_log_temps=logt.grad,
                grad_logits=lg.grad, layers=layers)


def build_cuda_module(work, log_temps=None, device="cuda", **kw):
    from basd_b200.losses import BASDLoss
    torch.manual_seed(SELECTOR_SEED)
    mod = BASDLoss(criterion(work), work.d_student, work.d_teacher, work.student_depth,
                   work.n_student, config=types.SimpleNamespace(num_extraction_points=work.num_points),
                   teacher_has_cls_token=work.has_cls, **kw)
    if log_temps is not None:
        with torch.no_grad():
            mod.layer_selector.log_temperatures.copy_(torch.as_tensor(log_temps))
    return mod.to(device)


def run_cuda(work, inputs, log_temps=None, device="cuda"):
    logits, targets, st, te, at = inputs
    mod = build_cuda_module(work, log_temps, device)
    st_d = {k: v.to(device).requires_grad_(True) for k, v in st.items()}
    te_d = {k: v.to(device) for k, v in te.items()}
    at_d = {k: v.to(device) for k, v in at.items()}
    lg = logits.to(device).requires_grad_(True)
    loss = mod(lg, targets.to(device), st_d, te_d, at_d)
    loss.backward()
    torch.cuda.synchronize()
    sel = mod.layer_selector
    layers = mod.token_layers
    return dict(loss=loss.detach().cpu(), ce=mod.last["ce"].cpu(), geo=mod.last["geo"].cpu(),
                geo_terms=mod.last["geo_terms"].cpu(),
                ranks=[sel.subspace_ranks[k] for k in sorted(sel.subspace_ranks.keys())],
                weights=mod.last["weights"].cpu(), dist=sel.last_state.dist.cpu(),
                grad_students={l: st_d[l].grad.float().cpu() for l in layers},
                grad_log_temps=sel.log_temperatures.grad.cpu(), grad_logits=lg.grad.cpu(),
                layers=layers, module=mod)


def cosine(a, b):
    return float(torch.nn.functional.cosine_similarity(a.flatten().double(), b.flatten().double(), dim=0))
