"""Attention-weighted Procrustes loss on the B200 kernels.

Interface mirror of the reference's ``src/losses/relational.py:5-50``.
"""
from __future__ import annotations

import torch

from .._autograd import ProcrustesGeo, StepContext


def geometric_relational_loss(student_tokens: torch.Tensor, teacher_tokens: torch.Tensor,
                              teacher_attn: torch.Tensor, *, has_cls_token: bool) -> torch.Tensor:
    """mean_b( tr_s + tr_t - 2 ||s_w^T t_w||_* ), differentiable w.r.t. ``student_tokens``.

    ``teacher_tokens`` is (B, N_s, D_t) (already aligned, as in combined.py:69-75);
    ``teacher_attn`` is (B, H, N_t+1, N_t+1) with CLS or (B, H, N_t, N_t) without; a
    pre-reduced importance row (B, N_t) is accepted as well."""
    if student_tokens.shape[1] != teacher_tokens.shape[1]:
        raise ValueError("teacher tokens must already be aligned to the student token count")
    step = StepContext(teachers=[teacher_tokens], attns=[teacher_attn], has_cls=has_cls_token,
                       n_student=student_tokens.shape[1], proj_s=None, proj_t=None)
    one = torch.ones(1, 1, dtype=torch.float32, device=student_tokens.device)
    return ProcrustesGeo.apply(one, step, student_tokens)
