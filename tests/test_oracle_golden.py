"""Pins the oracle port (oracle/ref_port.py) to outputs of the real reference
(tests/golden/*.pt, written by tests/golden/make_golden.py from /root/reference)."""
import pytest
import torch

import basd_b200.synthetic as syn
from tests import _cases as cs

CASES = ["c1_b16_seed0", "c1_b16_seed1_temps", "c2_b4_seed0", "c3_b8_seed0"]


@pytest.mark.parametrize("name", CASES)
def test_port_reproduces_reference(name):
    gold = cs.golden(name)
    work = cs.workload(gold["workload"], gold["batch"])
    inputs = syn.make_inputs(work, seed=gold["seed"])
    logits, _, st, te, at = inputs
    fp = cs.fingerprint([logits] + [st[k] for k in sorted(st)] + [te[k] for k in sorted(te)] +
                        [at[k] for k in sorted(at)])
    assert torch.allclose(fp, gold["input_fingerprint"], rtol=1e-9), "synthetic generator drifted"
    out = cs.run_oracle(work, inputs, gold["log_temperatures"])
    assert out["layers"] == gold["token_layers"]
    assert out["ranks"] == gold["ranks"].tolist()
    # same LAPACK, same order of operations: agreement should be at rounding level
    assert torch.allclose(out["loss"], gold["loss"], rtol=1e-6)
    assert torch.allclose(out["weights"], gold["weights"], atol=1e-6)
    assert torch.allclose(out["grad_log_temps"], gold["grad_log_temperatures"], rtol=1e-3, atol=1e-7)
    for layer in out["layers"]:
        g = out["grad_students"][layer]
        ref = gold["grad_student"][layer]
        idx = cs.probe_indices(g.numel(), 100 + layer)
        assert torch.allclose(g.norm(), ref["norm"], rtol=1e-4)
        assert cs.cosine(g.flatten()[idx], ref["probe"]) > 0.99999
    idx = cs.probe_indices(out["grad_logits"].numel(), 5)
    assert torch.allclose(out["grad_logits"].flatten()[idx], gold["grad_logits_probe"], atol=1e-7)


def test_mp_rank_median_and_strictness():
    # torch.median is the LOWER middle element and the comparison is strict (layer_selector.py:17-19)
    from oracle import kernel_model as km
    lam = torch.tensor([8.0, 4.0, 2.0, 1.0])          # descending
    # median (lower middle of ascending [1,2,4,8]) = 2; edge = 2*(1+sqrt(4/400))^2 = 2.42
    assert km.mp_rank_from_spectrum(lam, 400, 3) == 2
    assert km.mp_rank_from_spectrum(lam, 400, 1) == 1  # cap D_s-1


def test_extraction_layers_match_reference_rule():
    from oracle import ref_port as rp
    assert rp.extraction_layers(12, 4) == [0, 4, 7, 11]
    assert rp.extraction_layers(12, 1) == [11]
    assert rp.extraction_layers(24, 4) == [0, 8, 15, 23]
