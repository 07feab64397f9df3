"""Synthetic inputs with the shapes the BASD loss sees (token mode).

There is no network for datasets or checkpoints, so the benchmark and the parity tests
feed the loss tokens/attention maps generated here.  Tokens get a decaying spectrum
(``randn * logspace(0,-2,D) @ Q``) so that the Marchenko-Pastur rank is non-trivial
(iid Gaussian features give rank ~1), teacher layers differ from one another (scale,
a layer-specific partial rotation, attention sharpness) so the mixing weights are not
uniform, and the attention maps are row-stochastic like a post-softmax map
(reference: src/models/teacher.py:35-37).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import torch


@dataclass
class Workload:
    """One named shape row of SURVEY.md §8 (C1..C5)."""
    name: str
    batch: int
    n_student: int
    n_teacher: int
    d_student: int
    d_teacher: int
    teacher_layers: int
    heads: int
    has_cls: bool
    token_dtype: torch.dtype
    student_depth: int = 12
    num_points: int = 4
    num_classes: int = 1000
    attn_dtype: torch.dtype = torch.float32
    extra: dict = field(default_factory=dict)


WORKLOADS = {
    # C1: DeiT-Tiny <- DeiT-Small, 32x32 / patch 4, CPU-runnable oracle case
    "c1": Workload("c1_deit_t_from_deit_s_b16", 16, 64, 64, 192, 384, 12, 6, True,
                   torch.float32, num_classes=100),
    # C2: DeiT-Small <- DeiT-Base, 224/16, batch 256, bf16 tokens (the headline config)
    "c2": Workload("c2_deit_s_from_deit_b_b256", 256, 196, 196, 384, 768, 12, 12, True,
                   torch.bfloat16),
    # C3: DeiT-Small <- ResNet-50 (single layer, no CLS, uniform attention)
    "c3": Workload("c3_deit_s_from_resnet50_b256", 256, 196, 49, 384, 2048, 1, 1, False,
                   torch.bfloat16),
    # C4: DeiT-Base <- ViT-L/16 (24 teacher layers)
    "c4": Workload("c4_deit_b_from_vit_l_b256", 256, 196, 196, 768, 1024, 24, 16, True,
                   torch.bfloat16),
}


def scaled(work: Workload, batch: int) -> Workload:
    out = Workload(**{**work.__dict__})
    out.batch = batch
    out.name = f"{work.name.rsplit('_b', 1)[0]}_b{batch}"
    return out


def _orthogonal(dim: int, gen: torch.Generator, device) -> torch.Tensor:
    q, _ = torch.linalg.qr(torch.randn(dim, dim, generator=gen, device=device))
    return q


def make_inputs(work: Workload, *, seed: int = 0, device="cpu", batch_offset: int = 0,
                uniform_attn: bool | None = None):
    """Returns (logits, targets, student_tokens, teacher_tokens, teacher_attns).

    ``batch_offset`` lets a data-parallel rank draw the samples
    ``[offset, offset+batch)`` of a larger virtual batch: sample ``b`` always comes from
    its own generator seeded with ``(seed, b)``, so concatenating ranks' shards equals a
    single-process draw of the whole batch (used by the N-GPU parity tests).
    """
    device = torch.device(device)
    gen = torch.Generator(device=device)
    gen.manual_seed(1000 + seed)
    layers = _extraction_layers(work.student_depth, work.num_points)
    d_s, d_t = work.d_student, work.d_teacher
    spec_s = torch.logspace(0, -2, d_s, device=device)
    spec_t = torch.logspace(0, -2, d_t, device=device)
    q_s = _orthogonal(d_s, gen, device)
    q_t = _orthogonal(d_t, gen, device)
    # layer-specific structure (shared across ranks: drawn before any per-sample draw)
    s_rot = {l: _orthogonal(d_s, gen, device) for l in layers}
    s_off = {l: torch.randn(d_s, generator=gen, device=device) * 0.5 for l in layers}
    t_rot = [_orthogonal(d_t, gen, device) for _ in range(work.teacher_layers)]
    t_off = [torch.randn(d_t, generator=gen, device=device) * 0.5
             for _ in range(work.teacher_layers)]

    b = work.batch
    sgen = torch.Generator(device=device)

    def per_sample(shape_tail, salt):
        rows = []
        for i in range(b):
            sgen.manual_seed((seed * 7919 + salt) * 1_000_003 + batch_offset + i)
            rows.append(torch.randn(*shape_tail, generator=sgen, device=device))
        return torch.stack(rows)

    students = {}
    for n, l in enumerate(layers):
        base = per_sample((work.n_student, d_s), 11 + n) * spec_s
        mix = 0.25 + 0.5 * n / max(1, len(layers) - 1)
        x = (1 - mix) * (base @ q_s) + mix * (base @ s_rot[l]) + s_off[l]
        students[l] = x.to(work.token_dtype)

    teachers, attns = {}, {}
    uniform = (not work.has_cls) if uniform_attn is None else uniform_attn
    for l in range(work.teacher_layers):
        base = per_sample((work.n_teacher, d_t), 101 + l) * spec_t
        mix = l / max(1, work.teacher_layers - 1)
        x = (1 - 0.6 * mix) * (base @ q_t) + 0.6 * mix * (base @ t_rot[l]) + t_off[l]
        teachers[l] = ((1 + 0.2 * l) * x).to(work.token_dtype)
        side = work.n_teacher + (1 if work.has_cls else 0)
        if uniform:
            # CNN teacher: ones/N (reference: src/models/teacher.py:184-191)
            a = torch.full((b, work.heads, side, side), 1.0 / side, device=device)
        else:
            a = torch.softmax((1 + 0.3 * l) * per_sample((work.heads, side, side), 301 + l),
                              dim=-1)
        attns[l] = a.to(work.attn_dtype)

    logits = per_sample((work.num_classes,), 7)
    targets = (per_sample((1,), 9)[:, 0].abs() * 7919.0).long() % work.num_classes
    return logits, targets, students, teachers, attns


def _extraction_layers(depth: int, points: int) -> list[int]:
    # same rule as the reference (src/losses/combined.py:34-40)
    if points == 1:
        return [depth - 1]
    return [round(i * (depth - 1) / (points - 1)) for i in range(points)]


def make_inputs_fast(work: Workload, *, seed: int = 0, device="cuda"):
    """Bulk (non per-sample-seeded) variant for full-size benchmark shapes: same
    distribution, one generator, a few large ``randn`` calls on the target device."""
    device = torch.device(device)
    gen = torch.Generator(device=device)
    gen.manual_seed(2000 + seed)
    layers = _extraction_layers(work.student_depth, work.num_points)
    d_s, d_t, b = work.d_student, work.d_teacher, work.batch
    spec_s = torch.logspace(0, -2, d_s, device=device)
    spec_t = torch.logspace(0, -2, d_t, device=device)
    q_s = _orthogonal(d_s, gen, device)
    q_t = _orthogonal(d_t, gen, device)
    students, teachers, attns = {}, {}, {}
    for n, l in enumerate(layers):
        base = torch.randn(b, work.n_student, d_s, generator=gen, device=device) * spec_s
        mix = 0.25 + 0.5 * n / max(1, len(layers) - 1)
        rot = _orthogonal(d_s, gen, device)
        off = torch.randn(d_s, generator=gen, device=device) * 0.5
        students[l] = ((1 - mix) * (base @ q_s) + mix * (base @ rot) + off).to(work.token_dtype)
    side = work.n_teacher + (1 if work.has_cls else 0)
    for l in range(work.teacher_layers):
        base = torch.randn(b, work.n_teacher, d_t, generator=gen, device=device) * spec_t
        mix = l / max(1, work.teacher_layers - 1)
        rot = _orthogonal(d_t, gen, device)
        off = torch.randn(d_t, generator=gen, device=device) * 0.5
        x = (1 - 0.6 * mix) * (base @ q_t) + 0.6 * mix * (base @ rot) + off
        teachers[l] = ((1 + 0.2 * l) * x).to(work.token_dtype)
        if not work.has_cls:
            a = torch.full((b, work.heads, side, side), 1.0 / side, device=device)
        else:
            a = torch.softmax((1 + 0.3 * l) * torch.randn(b, work.heads, side, side,
                                                          generator=gen, device=device), dim=-1)
        attns[l] = a.to(work.attn_dtype)
    logits = torch.randn(b, work.num_classes, generator=gen, device=device)
    targets = torch.randint(0, work.num_classes, (b,), generator=gen, device=device)
    return logits, targets, students, teachers, attns
