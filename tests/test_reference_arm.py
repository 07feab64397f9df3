"""CPU-side contracts of the reference arm: (1) where /root/reference exists (the build
container), the oracle port is compared LIVE with the unmodified reference on a fresh random
draw that is not among the committed goldens; (2) `bench.py --impl reference` prints the JSON
line the driver expects.  Neither needs a GPU."""
import json
import os
import subprocess
import sys
import types

import pytest
import torch

import basd_b200.synthetic as syn
from tests import _cases as cs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "src", "losses")),
                    reason="the reference tree only exists in the build container")
def test_port_matches_the_live_reference_on_a_fresh_draw():
    sys.path.insert(0, REFERENCE)
    try:
        from src.losses.combined import BASDLoss as RefLoss
    finally:
        sys.path.remove(REFERENCE)
    work = cs.workload("c1", 8)
    inputs = syn.make_inputs(work, seed=77)                      # not a golden seed
    temps = [0.9, 0.2, 0.541, 1.4]
    logits, targets, st, te, at = inputs
    torch.manual_seed(cs.SELECTOR_SEED)
    ref = RefLoss(cs.criterion(work), work.d_student, work.d_teacher, work.student_depth, work.n_student,
                  config=types.SimpleNamespace(num_extraction_points=work.num_points),
                  teacher_has_cls_token=work.has_cls)
    with torch.no_grad():
        ref.layer_selector.log_temperatures.copy_(torch.tensor(temps))
    st32 = {k: v.float().requires_grad_(True) for k, v in st.items()}
    lg = logits.clone().requires_grad_(True)
    loss = ref(lg, targets, st32, {k: v.float() for k, v in te.items()}, {k: v.float() for k, v in at.items()})
    loss.backward()
    out = cs.run_oracle(work, inputs, temps)
    assert ref.token_layers == out["layers"]
    assert [ref.layer_selector.subspace_ranks[k] for k in sorted(ref.layer_selector.subspace_ranks)] == out["ranks"]
    assert torch.allclose(out["loss"], loss.detach(), rtol=1e-6)
    assert torch.allclose(out["grad_log_temps"], ref.layer_selector.log_temperatures.grad, rtol=1e-3, atol=1e-7)
    for layer in out["layers"]:
        assert cs.cosine(out["grad_students"][layer], st32[layer].grad) > 0.99999
    assert torch.allclose(out["grad_logits"], lg.grad, atol=1e-7)


def test_bench_reference_arm_prints_the_contract_line():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1",
                          "--batch", "16", "--cpu-batch", "4", "--steps", "1", "--warmup", "0",
                          "--features", "spectral"],
                         capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, res.stdout
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "basd_loss_fwd_bwd_samples_per_sec"
    assert line["unit"] == "samples/s" and line["higher_is_better"] is True and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["config"]["workload"].startswith("c1")


def test_bench_reference_arm_is_rank0_only():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--workload", "c1", "--batch", "16", "--cpu-batch", "4", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    assert not [ln for ln in res.stdout.splitlines() if ln.startswith("{")]


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "src", "losses")),
                    reason="the reference tree only exists in the build container")
def test_reference_checkpoint_state_loads_into_the_replacement(tmp_path):
    """`accelerator.register_for_checkpointing(basd_loss)` (trainer.py:84) saves the module's
    state_dict; a checkpoint written by the reference must load into the replacement (same keys, shapes,
    dtypes), reproduce the reference's RNG-drawn projections and round-trip back into the reference."""
    sys.path.insert(0, REFERENCE)
    try:
        from src.losses.combined import BASDLoss as RefLoss
    finally:
        sys.path.remove(REFERENCE)
    from basd_b200.losses import BASDLoss
    cfg = types.SimpleNamespace(num_extraction_points=4)
    crit = torch.nn.CrossEntropyLoss()
    torch.manual_seed(7)
    ref = RefLoss(crit, 192, 384, 12, 64, config=cfg, teacher_has_cls_token=True)
    with torch.no_grad():
        ref.layer_selector.log_temperatures.copy_(torch.tensor([0.1, 0.5, 0.9, 1.3]))
    path = tmp_path / "custom_checkpoint_0.pkl"
    torch.save(ref.state_dict(), path)
    torch.manual_seed(99)                                   # different draws: loading must overwrite them
    mine = BASDLoss(crit, 192, 384, 12, 64, config=cfg, teacher_has_cls_token=True)
    assert list(mine.state_dict().keys()) == list(ref.state_dict().keys())
    missing, unexpected = mine.load_state_dict(torch.load(path), strict=True)
    assert not missing and not unexpected
    for key, value in ref.state_dict().items():
        got = mine.state_dict()[key]
        assert got.dtype == value.dtype and got.shape == value.shape and torch.equal(got, value), key
    assert mine.token_layers == ref.token_layers
    assert [n for n, _ in mine.named_parameters()] == [n for n, _ in ref.named_parameters()]
    ref.load_state_dict(mine.state_dict(), strict=True)     # and back
    # same seed -> same projections as the reference draws them (orthogonal_ order, layer_selector.py:51-56)
    torch.manual_seed(7)
    again = BASDLoss(crit, 192, 384, 12, 64, config=cfg, teacher_has_cls_token=True)
    assert torch.equal(again.layer_selector.proj_s, ref.layer_selector.proj_s)
    assert torch.equal(again.layer_selector.proj_t, ref.layer_selector.proj_t)
