"""The two hand-written autograd stages of the BASD loss.

  MixingWeights : (log_temperatures, student tokens...) -> softmax mixing weights (E, L)
                  forward  = statistics + all-reduce + selector kernels
                  backward = closed-form SVD/eigh backward (selector_backward)
  ProcrustesGeo : (mixing weights, student tokens...)   -> mean Procrustes loss (scalar)
                  forward  = mix+align, per-sample N x N Procrustes kernels
                  backward = closed-form nuclear-norm gradients + one pass over the teacher
                             stack for dL/dweights

No torch op computes anything on these paths; torch.autograd only chains the two stages
(dL/dweights flows from the second into the first) and sums the two student-token
gradients.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as dist

from . import _engine as eng


@dataclass
class StepContext:
    """Per-call inputs that are not differentiated plus the state shared by both stages."""
    teachers: list
    attns: list
    has_cls: bool
    n_student: int
    proj_s: torch.Tensor
    proj_t: torch.Tensor
    group: object = None
    world: int = 1
    stats: eng.Stats | None = None
    selector: eng.SelectorState | None = None
    procrustes: eng.ProcrustesState | None = None


def world_size(group) -> int:
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group)
    return 1


def _check_inputs(students, step: StepContext):
    for t in list(students) + list(step.teachers):
        if not t.is_cuda:
            raise RuntimeError("BASD loss kernels run on CUDA tensors only (no CPU fallback)")
        if t.dim() != 3:
            raise ValueError(f"token tensors must be (B, N, D), got {tuple(t.shape)}")
    b = students[0].shape[0]
    if any(t.shape[0] != b for t in step.teachers):
        raise ValueError("student and teacher batch sizes differ")


class MixingWeights(torch.autograd.Function):
    @staticmethod
    def forward(ctx, log_temps, step: StepContext, *students):
        students = [s.detach().contiguous() for s in students]
        step.teachers = [t.detach().contiguous() for t in step.teachers]
        _check_inputs(students, step)
        logt = log_temps.detach().to(torch.float32).contiguous()
        stats, flat = eng.statistics(students, step.teachers, step.attns, step.has_cls)
        if step.world > 1:
            eng._all_reduce(flat, step.group)
        b, n_s, _ = students[0].shape
        n_t = step.teachers[0].shape[1]
        step.stats = stats
        step.selector = eng.selector_forward(stats, b * n_s * step.world, b * n_t * step.world,
                                             step.proj_s, step.proj_t, logt, step.group, step.world)
        ctx.step = step
        ctx.students = students
        ctx.logt = logt
        return step.selector.weights.clone()

    @staticmethod
    def backward(ctx, d_weights):
        step = ctx.step
        grads, d_logt = eng.selector_backward(ctx.students, step.selector, step.proj_s, ctx.logt,
                                              d_weights, step.group, step.world)
        need = ctx.needs_input_grad
        out = [d_logt if need[0] else None, None]
        for i, g in enumerate(grads):
            out.append(g if need[2 + i] else None)
        return tuple(out)


class ProcrustesGeo(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weights, step: StepContext, *students):
        students = [s.detach().contiguous() for s in students]
        w = weights.detach().to(torch.float32).contiguous()
        if step.stats is None:       # stand-alone use: only the importance rows are needed
            step.teachers = [t.detach().contiguous() for t in step.teachers]
            _check_inputs(students, step)
            step.stats = eng.attention_only_stats(step.teachers, step.attns, step.has_cls)
        with_grad = any(ctx.needs_input_grad)
        pro = eng.procrustes_forward(students, step.teachers, step.stats, w, step.n_student,
                                     with_grad)
        step.procrustes = pro
        ctx.step = step
        ctx.students = students
        ctx.pro = pro
        return pro.geo.clone()

    @staticmethod
    def backward(ctx, grad_geo):
        step = ctx.step
        grads, d_weights, _ = eng.procrustes_backward(ctx.students, step.teachers, step.stats,
                                                      ctx.pro, grad_geo, step.n_student)
        need = ctx.needs_input_grad
        out = [d_weights if need[0] else None, None]
        for i, g in enumerate(grads):
            out.append(g if need[2 + i] else None)
        return tuple(out)
