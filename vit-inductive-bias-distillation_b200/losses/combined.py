"""BASD loss = CE + layer-selector-mixed, attention-weighted Procrustes, UW-SO weighted.

Interface mirror of the reference's ``src/losses/combined.py`` (``BASDLoss`` at :17-85,
``_align_token_count`` at :9-14): same constructor and ``forward`` signature, same
``token_layers`` / ``layer_selector`` attributes and state-dict keys
(``layer_selector.proj_s``, ``layer_selector.proj_t``, ``layer_selector.log_temperatures``).
"""
from __future__ import annotations

import torch
import torch.distributed as dist
import torch.nn as nn

from .._autograd import ProcrustesGeo
from .._native import call, dtype_code, ptr, stream
from .layer_selector import GrassmannianLayerSelector


def _align_token_count(tokens: torch.Tensor, target_n: int) -> torch.Tensor:
    """1-D linear resampling of the token axis (reference :9-14) on the mix kernel
    (a single "layer" with weight 1)."""
    if tokens.shape[1] == target_n:
        return tokens
    b, n_src, d = tokens.shape
    x = tokens.contiguous()
    out = torch.empty(1, b, target_n, d, dtype=torch.float32, device=x.device)
    one = torch.ones(1, 1, dtype=torch.float32, device=x.device)
    import ctypes
    ptrs = (ctypes.c_void_p * 1)(x.data_ptr())
    call("basd_mix_interp", ptrs, 1, 1, ptr(one), dtype_code(x), b, n_src, target_n, d, ptr(out),
         0, stream())
    return out[0].to(tokens.dtype)


class BASDLoss(nn.Module):
    def __init__(self, base_criterion: nn.Module, student_dim: int, teacher_dim: int,
                 student_depth: int, num_student_tokens: int, *, config,
                 teacher_has_cls_token: bool, process_group=None, sync_stats: bool = True):
        super().__init__()
        self.base_criterion = base_criterion
        self.teacher_has_cls_token = teacher_has_cls_token
        self.num_student_tokens = num_student_tokens
        points = config.num_extraction_points
        if points == 1:                                     # reference :34-40
            self.token_layers = [student_depth - 1]
        else:
            self.token_layers = [round(i * (student_depth - 1) / (points - 1))
                                 for i in range(points)]
        self.layer_selector = GrassmannianLayerSelector(
            num_extraction_points=len(self.token_layers), student_dim=student_dim,
            teacher_dim=teacher_dim)
        self.layer_selector.process_group = process_group
        self.layer_selector.sync_stats = sync_stats
        self.last = {}

    def forward(self, student_output, targets, student_intermediates, all_teacher_tokens,
                all_teacher_attns):
        ce_loss = self.base_criterion(student_output, targets)                  # :56
        sel = self.layer_selector
        weights, step = sel.mixing_weights(                                     # :58-61
            student_intermediates, all_teacher_tokens, all_teacher_attns, self.token_layers,
            has_cls=self.teacher_has_cls_token, n_student=self.num_student_tokens)
        students = [student_intermediates[l] for l in self.token_layers]
        geo_loss = ProcrustesGeo.apply(weights, step, *students)                # :63-76
        sel.last_state.sweeps["procrustes"] = step.procrustes.sweeps
        # UW-SO (reference :78-85): w_i = (1/L_i) / sum_j (1/L_j), weights detached.
        vals = torch.stack([ce_loss.detach().float(), geo_loss.detach()])
        if step.world > 1:       # weights from the global (concatenated-batch) losses
            vals = vals.clone()
            dist.all_reduce(vals, group=step.group)
            vals = vals / step.world
        eps = torch.finfo(ce_loss.dtype).eps
        inv = 1.0 / vals.clamp(min=eps)
        share = inv / inv.sum()
        self.last = {"ce": ce_loss.detach(), "geo": geo_loss.detach(), "weights": weights.detach(),
                     "geo_terms": step.procrustes.geo_terms, "share": share,
                     "global_loss": (share * vals).sum()}
        return share[0] * ce_loss + share[1] * geo_loss
