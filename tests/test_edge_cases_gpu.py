"""Edge cases of the reference behaviour (SURVEY.md §8 a-rows) on the CUDA path vs the live oracle."""
import types

import pytest
import torch

import basd_b200.synthetic as syn
from tests import _cases as cs

pytestmark = pytest.mark.gpu


def _work(**kw):
    base = dict(name="edge", batch=8, n_student=64, n_teacher=64, d_student=128, d_teacher=256,
                teacher_layers=3, heads=2, has_cls=True, token_dtype=torch.float32, student_depth=12,
                num_points=4, num_classes=10)
    base.update(kw)
    return syn.Workload(**base)


def _compare(work, seed=0, temps=None, cos_tol=0.999):
    inputs = syn.make_inputs(work, seed=seed, uniform_attn=False)
    ref = cs.run_oracle(work, inputs, temps)
    got = cs.run_cuda(work, inputs, temps)
    assert got["ranks"] == ref["ranks"], (got["ranks"], ref["ranks"])
    assert (got["weights"] - ref["weights"]).abs().max() < 1e-4
    assert abs(float(got["loss"]) - float(ref["loss"])) / abs(float(ref["loss"])) < 1e-3
    for layer in ref["layers"]:
        c = cs.cosine(got["grad_students"][layer], ref["grad_students"][layer])
        assert c > cos_tol, (layer, c)
    if work.teacher_layers > 1:
        assert cs.cosine(got["grad_log_temps"], ref["grad_log_temps"]) > cos_tol
    return got, ref


def test_no_cls_multilayer_downsampled_teacher_two_points():
    # ViT teacher without CLS token, more teacher tokens than student tokens (256 -> 196), E = 2.
    # D_s = 192 <= N = 196 < D_t = 384: the student side is direct and, because D_t <= 2N, the
    # teacher side is taken direct as well (X = A^T B, no Gram): with a Gram teacher side the
    # ~10 singular directions below sqrt(eps) were lost to both gradients (cosine 0.9989).
    _compare(_work(n_student=196, n_teacher=256, d_student=192, d_teacher=384, has_cls=False,
                   num_points=2, batch=4), temps=[0.4, 0.9])


def test_direct_student_side_with_gram_teacher_side():
    # D_s = 128 <= N = 196, D_t = 512 > 2N: mixed factors (direct student, pivoted-Cholesky teacher)
    _compare(_work(n_student=196, n_teacher=196, d_student=128, d_teacher=512, batch=4), cos_tol=0.998)


def test_single_extraction_point_and_bf16_tokens():
    _compare(_work(num_points=1, token_dtype=torch.bfloat16, d_student=128, d_teacher=256, batch=8),
             temps=[0.7])


def test_eight_extraction_points():
    _compare(_work(num_points=8, student_depth=24, batch=8),
             temps=[0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0])


def test_token_count_beyond_the_shared_memory_kernels():
    # N = 256 > 224: per-sample factorisations leave shared memory (cluster / global paths)
    _compare(_work(n_student=256, n_teacher=256, d_student=128, d_teacher=128, batch=4, teacher_layers=2))


def test_upsampled_teacher_with_cls():
    _compare(_work(n_student=196, n_teacher=49, d_student=192, d_teacher=320, batch=4))


def test_fewer_rows_than_dims_uses_the_small_gram_population():
    # M = 64 token rows < D_s = 128: the reference takes the spectrum of the M x M Gram
    # (layer_selector.py:14-15) -- the M largest eigenvalues of the D x D one
    # (seed 1: every teacher layer keeps its nearest eigenvalue >= 2 % away from the MP edge; at
    # seed 0 one sits 9e-4 from it and fp32 noise decides the rank -- see the tie flag in
    # test_loss_parity_gpu._ranks_ok)
    # With M < D_s the centred student Gram has a null space; its basis is completed (selector.cu:
    # projector + pivoted Cholesky) so that the (I - V V^T) term of the thin-SVD backward is present.
    _compare(_work(batch=1, n_student=64, n_teacher=64, d_student=128), seed=1)
    from basd_b200.losses import marchenko_pastur_rank
    from oracle import ref_port as rp
    torch.manual_seed(5)
    feats = torch.randn(50, 96) * torch.logspace(0, -2, 96)
    assert marchenko_pastur_rank(feats.cuda()) == rp.mp_rank(feats)


def test_degenerate_teacher_layer_gives_nan_like_the_reference():
    # an all-zero teacher layer has MP rank 0: sum(sw) over an empty set -> 0/0 (layer_selector.py:105).
    # The reference's NaN mixing weights then make LAPACK's SVD raise on CPU; the CUDA path has no
    # host synchronisation to raise from and returns the NaN loss.
    work = _work(batch=8)
    logits, targets, st, te, at = syn.make_inputs(work, seed=1, uniform_attn=False)
    te[1] = torch.zeros_like(te[1])
    try:
        ref = cs.run_oracle(work, (logits, targets, st, te, at), None)
        assert torch.isnan(ref["loss"])
    except torch.linalg.LinAlgError as err:
        assert "non-finite" in str(err)
    mod = cs.build_cuda_module(work)
    dev = "cuda"
    with torch.no_grad():
        loss = mod(logits.to(dev), targets.to(dev), {k: v.to(dev) for k, v in st.items()},
                   {k: v.to(dev) for k, v in te.items()}, {k: v.to(dev) for k, v in at.items()})
    assert torch.isnan(loss)


def test_importance_rows_instead_of_attention_maps():
    # the pre-reduced (B, N_t) importance row is accepted in place of the full map
    work = _work(batch=8)
    inputs = syn.make_inputs(work, seed=2, uniform_attn=False)
    full = cs.run_cuda(work, inputs, None)
    logits, targets, st, te, at = inputs
    rows = {k: v[:, :, 0, 1:].mean(dim=1) for k, v in at.items()}
    cls_only = {k: v[:, :, 0, :].contiguous() for k, v in at.items()}
    for variant in (rows, cls_only):
        got = cs.run_cuda(work, (logits, targets, st, te, variant), None)
        assert abs(float(got["loss"]) - float(full["loss"])) < 1e-5 * abs(float(full["loss"]))


@pytest.mark.parametrize("scale,decay,cos_tol", [(10, 1, 0.9999), (50, 2, 0.9999), (100, 3, 0.9995)])
def test_ill_conditioned_per_sample_tokens(scale, decay, cos_tol):
    """Per-sample token matrices with condition numbers 1e2 .. 2e4 (a geometric feature spectrum
    plus two "massive activation" columns, as trained ViTs show) through the Gram-side Procrustes:
    the factors square the condition number, so this pins the floor / rank-cut choices
    (2.5e-4 as the singular-value floor gave cosine 0.9991 at scale 50; tests/tools/floor_sweep.py)."""
    from basd_b200.losses import geometric_relational_loss
    from oracle import ref_port as rp
    gen = torch.Generator().manual_seed(17)

    def tokens(d):
        q, _ = torch.linalg.qr(torch.randn(d, d, generator=gen))
        y = (torch.randn(4, 196, d, generator=gen) @ q) * torch.logspace(0, -decay, d)
        y[..., 7] *= scale
        y[..., 100] *= scale
        return y + 0.3 * torch.randn(1, 1, d, generator=gen)

    s, t = tokens(384), tokens(768)
    attn = torch.softmax(torch.randn(4, 6, 197, 197, generator=gen), dim=-1)
    sg = s.clone().requires_grad_(True)
    ref = rp.procrustes_loss(sg, t, attn, True)
    ref.backward()
    sd = s.cuda().requires_grad_(True)
    got = geometric_relational_loss(sd, t.cuda(), attn.cuda(), has_cls_token=True)
    got.backward()
    cos = cs.cosine(sd.grad.cpu(), sg.grad)
    print("ill-conditioned", scale, decay, "loss", float(got), float(ref), "cosine", cos)
    assert abs(float(got) - float(ref)) / abs(float(ref)) < 1e-3
    assert cos > cos_tol


@pytest.mark.parametrize("scale,decay,cos_tol", [(10, 1, 0.9999), (50, 2, 0.9995)])
def test_ill_conditioned_tokens_through_the_whole_loss(scale, decay, cos_tol):
    """The same distortion on features of random-init backbones through the WHOLE loss: at (50, 2) the
    global statistics span cond^2 = 4e6 and the selector's gradient divides by eigenvalue gaps.
    Measured on B200 at (50, 2): cosine -0.98 before (a) the fp64 projection of the statistics
    (rotate_f64.cu; 0.9989 -> 0.99999 in the CPU model) and (b) the higher singular-value floor for the
    derived Procrustes vectors (the teacher-token gradient was 41x too large, which flipped the sign of
    dL/dweights); 0.99994 after.  Round 2: the student backbone takes the reference's fan-in initialisation
    (train.py:19-32), whose features are harder here: 0.99899 with round 1's kernels, 0.99986 with the
    mean-shifted statistics and the diagonally shifted Cholesky (the north-star bound is 0.999).  Documented limit: at (100, 3) the statistics span
    cond^2 = 3e8 > 1/eps, the Gram is numerically rank deficient in fp32 and the selector share of the
    gradient is lost (cosine 0.06 .. 0.59) -- the reference resolves it because it takes the SVD of the
    tokens themselves."""
    from basd_b200 import backbone_features as bf
    from oracle import ref_port as rp
    work = cs.workload("c2", 4)
    layers = rp.extraction_layers(work.student_depth, work.num_points)
    logits, targets, st, te, at = bf.backbone_inputs("deit_small", "deit_base", layers, 4, seed=0,
                                                     device="cuda")

    def distort(x):
        d = x.shape[-1]
        q, _ = torch.linalg.qr(torch.randn(d, d, generator=torch.Generator().manual_seed(d)))
        y = (x.cpu().float() @ q) * torch.logspace(0, -decay, d)
        y[..., 7] *= scale
        y[..., 100] *= scale
        return y

    work.token_dtype = torch.float32
    inputs = (logits.cpu(), targets.cpu(), {k: distort(v) for k, v in st.items()},
              {k: distort(v) for k, v in te.items()}, {k: v.cpu() for k, v in at.items()})
    ref = cs.run_oracle(work, inputs)
    got = cs.run_cuda(work, inputs)
    cosines = [cs.cosine(got["grad_students"][l], ref["grad_students"][l]) for l in ref["layers"]]
    print("ill-conditioned whole loss", scale, decay, "ranks", got["ranks"], ref["ranks"], "weights",
          float((got["weights"] - ref["weights"]).abs().max()), "cosines", cosines)
    assert abs(float(got["loss"]) - float(ref["loss"])) / abs(float(ref["loss"])) < 1e-3
    if got["ranks"] == ref["ranks"]:
        assert (got["weights"] - ref["weights"]).abs().max() < 1e-4
    assert min(cosines) > cos_tol


@pytest.mark.parametrize("n,ds,dt,batch", [(49, 128, 256, 8), (225, 256, 384, 4), (50, 64, 128, 8)])
def test_token_counts_that_are_not_multiples_of_four(n, ds, dt, batch):
    """7 x 7, 15 x 15 grids: the N x N factor matrices get a padded pitch (128-bit row accesses);
    results must not depend on it."""
    _compare(_work(n_student=n, n_teacher=n, d_student=ds, d_teacher=dt, batch=batch))
