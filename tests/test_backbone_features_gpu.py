"""Parity on features captured from random-init backbones (BASELINE.json: "synthetic
ImageNet-shaped features from random-init DeiT/ViT/ResNet backbones"): random 224 x 224 images run
through timm-shaped random-init ViTs / torchvision's ResNet-50 (basd_b200/backbone_features.py), tokens and
attention captured with ``basd_b200.capture`` as the reference's trainer does, then the CUDA loss
against the live CPU oracle on the same tensors.  These token matrices are far from the hand-made
spectra of ``basd_b200.synthetic`` (LayerNorm-free residual streams, rank-deficient centred tokens,
near-uniform attention at initialisation), so they exercise the rank decisions of the path.

Tolerances: BASELINE.json's (weights 1e-4 abs, loss 1e-3 rel, gradient cosine 0.999).
"""
import pytest
import torch

from oracle import ref_port as rp
from basd_b200 import backbone_features as bb
from tests import _cases as cs
from tests.test_loss_parity_gpu import COS_TOL, LOSS_TOL, W_TOL, _check_selector_at_kernel_rank, _ranks_ok

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("key,student,teacher,batch,rows", [
    ("c1", "deit_tiny", "deit_small", 4, False),      # DeiT-T <- DeiT-S (direct student side)
    ("c2", "deit_small", "deit_base", 4, False),      # DeiT-S <- DeiT-B, full attention maps
    ("c2", "deit_small", "deit_base", 4, True),       # same, importance rows from the capture hook
    ("c3", "deit_small", "resnet50", 4, False),       # CNN teacher: 49 tokens resampled to 196
])
def test_random_init_backbones_against_live_oracle(key, student, teacher, batch, rows):
    work = cs.workload(key, batch)
    work.n_student = 196                               # 224 x 224 images, patch 16
    work.n_teacher = 49 if teacher == "resnet50" else 196
    layers = rp.extraction_layers(work.student_depth, work.num_points)
    full = bb.backbone_inputs(student, teacher, layers, batch, seed=3, device="cuda",
                              classes=work.num_classes, importance_rows=False)
    logits, targets, st, te, at = full
    assert st[layers[0]].shape == (batch, 196, work.d_student)
    assert te[0].shape == (batch, work.n_teacher, work.d_teacher)
    # bf16 workloads: both paths see the same bf16-rounded tokens (the oracle upcasts them exactly)
    st = {k: v.to(work.token_dtype) for k, v in st.items()}
    te = {k: v.to(work.token_dtype) for k, v in te.items()}
    ref = cs.run_oracle(work, (logits, targets, st, te, at))
    if rows:   # the (B, N) importance rows the capture hook emits instead of the full maps
        at = bb.backbone_inputs(student, teacher, layers, batch, seed=3, device="cuda",
                                classes=work.num_classes, importance_rows=True)[4]
        assert at[0].shape == (batch, work.n_teacher)
    inputs = (logits, targets, st, te, at)
    got = cs.run_cuda(work, inputs)
    print(key, student, teacher, "loss", float(got["loss"]), float(ref["loss"]), "geo", float(got["geo"]),
          float(ref["geo"]), "ranks", got["ranks"], ref["ranks"])
    assert _ranks_ok(got["ranks"], ref["ranks"], got["module"])
    if got["ranks"] == ref["ranks"]:
        assert (got["weights"] - ref["weights"]).abs().max() < W_TOL
    else:
        _check_selector_at_kernel_rank(work, inputs, None, got)
    assert abs(float(got["geo"]) - float(ref["geo"])) / abs(float(ref["geo"])) < LOSS_TOL
    assert abs(float(got["loss"]) - float(ref["loss"])) / abs(float(ref["loss"])) < LOSS_TOL
    for layer in ref["layers"]:
        c = cs.cosine(got["grad_students"][layer], ref["grad_students"][layer])
        print("  layer", layer, "grad cosine", c)
        assert c > COS_TOL
    if work.teacher_layers > 1 and got["ranks"] == ref["ranks"]:
        assert cs.cosine(got["grad_log_temps"], ref["grad_log_temps"]) > COS_TOL
    assert torch.allclose(got["grad_logits"], ref["grad_logits"], atol=1e-6, rtol=1e-3)
