"""Generates tests/golden/*.pt by running the UNMODIFIED reference (/root/reference/src/losses)
on seeded synthetic inputs.  Run in the build container only (the reference does not travel
to the GPU box); the fixtures it writes are committed.

    python tests/golden/make_golden.py

Inputs are not stored: they are re-drawn from the same seeds by
``basd_b200.synthetic.make_inputs`` (CPU generator, deterministic for a given torch build).
bf16 token configurations are handed to the reference upcast to fp32 -- the reference itself
cannot take bf16 tokens (layer_selector.py:72 multiplies them with fp32 buffers).
A fingerprint of the inputs is stored so a silent generator change is caught.
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from src.losses.combined import BASDLoss  # noqa: E402  (the real reference)

import basd_b200.synthetic as syn  # noqa: E402

CASES = {
    # name: (workload key, batch, seed, log_temperatures or None)
    "c1_b16_seed0": ("c1", 16, 0, None),
    "c1_b16_seed1_temps": ("c1", 16, 1, [0.3, 0.541, 0.8, 1.2]),
    "c2_b4_seed0": ("c2", 4, 0, None),
    "c3_b8_seed0": ("c3", 8, 0, None),
}
PROBE = 512


def fingerprint(tensors):
    return torch.tensor([float(t.double().sum()) for t in tensors], dtype=torch.float64)


def probe_indices(numel, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, numel, (PROBE,), generator=g)


def run_case(name, key, batch, seed, temps):
    work = syn.scaled(syn.WORKLOADS[key], batch)
    torch.manual_seed(1234)                      # selector buffers (orthogonal_ draws)
    crit = torch.nn.CrossEntropyLoss(label_smoothing=1.0 / work.num_classes)
    ref = BASDLoss(crit, work.d_student, work.d_teacher, work.student_depth, work.n_student,
                   config=types.SimpleNamespace(num_extraction_points=work.num_points),
                   teacher_has_cls_token=work.has_cls)
    if temps is not None:
        with torch.no_grad():
            ref.layer_selector.log_temperatures.copy_(torch.tensor(temps))
    logits, targets, st, te, at = syn.make_inputs(work, seed=seed)
    st32 = {k: v.float().requires_grad_(True) for k, v in st.items()}
    te32 = {k: v.float() for k, v in te.items()}
    lg = logits.clone().requires_grad_(True)
    # mixing weights are not an output of the reference: capture them from the softmax
    captured = []
    orig_softmax = torch.nn.functional.softmax

    def spy(x, dim=None, **kw):
        out = orig_softmax(x, dim=dim, **kw)
        captured.append(out.detach().clone())
        return out

    import src.losses.layer_selector as ls
    ls.F.softmax = spy
    try:
        loss = ref(lg, targets, st32, te32, at)
    finally:
        ls.F.softmax = orig_softmax
    loss.backward()
    sel = ref.layer_selector
    out = {
        "workload": key, "batch": batch, "seed": seed,
        "log_temperatures": sel.log_temperatures.detach().clone(),
        "token_layers": list(ref.token_layers),
        "loss": loss.detach().clone(),
        "ranks": torch.tensor([sel.subspace_ranks[k] for k in sorted(sel.subspace_ranks)]),
        "weights": torch.stack(captured),
        "grad_log_temperatures": sel.log_temperatures.grad.clone(),
        "grad_logits_probe": lg.grad.flatten()[probe_indices(lg.numel(), 5)].clone(),
        "input_fingerprint": fingerprint([logits] + [st[k] for k in sorted(st)] +
                                         [te[k] for k in sorted(te)] + [at[k] for k in sorted(at)]),
        "grad_student": {},
    }
    for layer in ref.token_layers:
        g = st32[layer].grad
        idx = probe_indices(g.numel(), 100 + layer)
        direction = torch.randn(g.numel(), generator=torch.Generator().manual_seed(200 + layer))
        out["grad_student"][layer] = {
            "norm": g.norm().clone(), "probe": g.flatten()[idx].clone(),
            "dot": (g.flatten() * direction).sum().clone(),
        }
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), f"{name}.pt")
    torch.save(out, path)
    print(name, "loss", float(loss), "ranks", out["ranks"].tolist(), "->", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    for name, (key, batch, seed, temps) in CASES.items():
        run_case(name, key, batch, seed, temps)
