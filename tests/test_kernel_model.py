"""The reformulated algorithm (oracle/kernel_model.py: token-space Grams, pivoted Cholesky,
N x N Procrustes, closed-form gradients) against the real reference's golden outputs, at
the tolerances BASELINE.json states: weights 1e-4 abs, loss 1e-3 rel, grad cosine 0.999."""
import pytest
import torch

import basd_b200.synthetic as syn
from oracle import kernel_model as km
from tests import _cases as cs


@pytest.mark.parametrize("name", ["c1_b16_seed1_temps", "c2_b4_seed0", "c3_b8_seed0"])
def test_model_matches_reference(name):
    gold = cs.golden(name)
    work = cs.workload(gold["workload"], gold["batch"])
    logits, targets, st, te, at = syn.make_inputs(work, seed=gold["seed"])
    proj_s, proj_t, _ = cs.selector_state(work)
    out = km.full_step_model(logits, targets, st, te, at, layers=gold["token_layers"], proj_s=proj_s,
                             proj_t=proj_t, log_temps=gold["log_temperatures"],
                             n_student=work.n_student, has_cls=work.has_cls,
                             criterion=cs.criterion(work))
    assert out["ranks"] == gold["ranks"].tolist()
    assert (out["weights"] - gold["weights"]).abs().max() < 1e-4
    assert abs(float(out["loss"]) - float(gold["loss"])) / abs(float(gold["loss"])) < 1e-3
    for layer in gold["token_layers"]:
        g = out["grad_students"][layer]
        ref = gold["grad_student"][layer]
        idx = cs.probe_indices(g.numel(), 100 + layer)
        assert cs.cosine(g.flatten()[idx], ref["probe"]) > 0.999
        assert abs(float(g.norm()) - float(ref["norm"])) / float(ref["norm"]) < 2e-2
    if work.teacher_layers > 1:
        assert cs.cosine(out["grad_log_temps"], gold["grad_log_temperatures"]) > 0.999


def test_model_matches_the_oracle_at_c4_shapes():
    """DeiT-B <- ViT-L/16 (D_s = 768, D_t = 1024, 24 teacher layers): no golden of this size is
    committed, so the reformulated algorithm is compared with the (golden-pinned) oracle port on a
    live draw."""
    work = cs.workload("c4", 8)
    inputs = syn.make_inputs(work, seed=11)
    ref = cs.run_oracle(work, inputs)
    logits, targets, st, te, at = inputs
    proj_s, proj_t, logt = cs.selector_state(work)
    out = km.full_step_model(logits, targets, st, te, at, layers=ref["layers"], proj_s=proj_s, proj_t=proj_t,
                             log_temps=logt, n_student=work.n_student, has_cls=work.has_cls,
                             criterion=cs.criterion(work))
    assert out["ranks"] == ref["ranks"]
    assert (out["weights"] - ref["weights"]).abs().max() < 1e-4
    assert abs(float(out["loss"]) - float(ref["loss"])) / abs(float(ref["loss"])) < 1e-3
    for layer in ref["layers"]:
        assert cs.cosine(out["grad_students"][layer], ref["grad_students"][layer]) > 0.999
    assert cs.cosine(out["grad_log_temps"], ref["grad_log_temps"]) > 0.999


def test_pivoted_cholesky_rank_deficient():
    torch.manual_seed(0)
    a = torch.randn(40, 12)
    k = a @ a.T                                    # rank 12 of 40
    low = km.pivoted_cholesky(k)
    assert int((low.abs().sum(0) > 0).sum()) == 12
    assert (low @ low.T - k).abs().max() < 1e-4 * k.abs().max()


def test_interp_taps_match_torch():
    x = torch.randn(2, 7, 49)
    ref = torch.nn.functional.interpolate(x, size=196, mode="linear", align_corners=False)
    lo, hi, frac = km.interp_taps(49, 196)
    mine = x[..., lo] * (1 - frac) + x[..., hi] * frac
    assert torch.allclose(mine, ref, atol=1e-6)
    ref = torch.nn.functional.interpolate(torch.randn(1, 3, 256), size=196, mode="linear")
    lo, hi, frac = km.interp_taps(256, 196)
    assert lo.max() < 256 and hi.max() < 256
