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
    contig: dict | None = None       # contiguous copies of strided student tokens, shared by both stages


def _view_key(t: torch.Tensor):
    return (t.data_ptr(), tuple(t.shape), tuple(t.stride()), t.dtype, t._version)


def _contiguous_students(step: StepContext, students):
    """Detached, contiguous student tokens.  CLS-sliced hook outputs are strided views: the copy is
    made once per step and reused by the second stage (it used to be made by both)."""
    if step.contig is None:
        step.contig = {}
    out = []
    for s in students:
        s = s.detach()
        if s.is_contiguous():
            out.append(s)
            continue
        key = _view_key(s)
        if key not in step.contig:
            step.contig[key] = s.contiguous()
        out.append(step.contig[key])
    return out


def world_size(group) -> int:
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group)
    return 1


_TOKEN_DTYPES = (torch.float32, torch.bfloat16)
MAX_POINTS, MAX_TEACHER_LAYERS = 8, 64        # MAX_E / MAX_L of csrc/mix.cu


def _check_inputs(students, step: StepContext):
    """Shape / dtype / device validation before any raw pointer reaches a kernel.  The reference
    fails in torch.stack / mm / bmm with a shape error on the same mistakes (layer_selector.py:72,
    128-129, relational.py:47); here the kernels would read or write out of bounds instead."""
    if not students or not step.teachers:
        raise ValueError("BASD loss needs at least one student and one teacher token tensor")
    if len(students) > MAX_POINTS:
        raise ValueError(f"at most {MAX_POINTS} extraction points are supported, got {len(students)}")
    if len(step.teachers) > MAX_TEACHER_LAYERS:
        raise ValueError(f"at most {MAX_TEACHER_LAYERS} teacher layers are supported, got {len(step.teachers)}")
    if len(step.attns) != len(step.teachers):
        raise ValueError(f"{len(step.teachers)} teacher token tensors but {len(step.attns)} attention maps")
    dev = students[0].device
    for name, group in (("student", students), ("teacher", step.teachers)):
        ref = group[0]
        for t in group:
            if not t.is_cuda:
                raise RuntimeError("BASD loss kernels run on CUDA tensors only (no CPU fallback)")
            if t.dim() != 3:
                raise ValueError(f"{name} token tensors must be (B, N, D), got {tuple(t.shape)}")
            if t.dtype not in _TOKEN_DTYPES:
                raise TypeError(f"{name} tokens must be float32 or bfloat16, got {t.dtype}")
            if t.device != dev:
                raise ValueError(f"{name} tokens live on {t.device}, expected {dev}")
            if t.shape != ref.shape or t.dtype != ref.dtype:
                raise ValueError(f"all {name} token tensors must share one shape and dtype: "
                                 f"{tuple(ref.shape)} {ref.dtype} vs {tuple(t.shape)} {t.dtype}")
    b, n_s, d_s = students[0].shape
    bt, n_t, d_t = step.teachers[0].shape
    if bt != b:
        raise ValueError(f"student batch {b} and teacher batch {bt} differ")
    if n_s != step.n_student:
        raise ValueError(f"student tokens have {n_s} tokens but num_student_tokens is {step.n_student} "
                         "(was the CLS token stripped?)")
    if d_s % 8 or d_t % 8:
        raise ValueError(f"token dimensions must be multiples of 8 (128-bit loads), got {d_s} and {d_t}")
    if min(b, n_s, n_t) < 1:
        raise ValueError("empty batch or token axis")
    for a in step.attns:
        if not a.is_cuda or a.device != dev:
            raise ValueError("attention maps must live on the tokens' CUDA device")
        if a.dtype not in _TOKEN_DTYPES:
            raise TypeError(f"attention maps must be float32 or bfloat16, got {a.dtype}")
    if step.proj_s is not None and tuple(step.proj_s.shape) != (d_s, d_s):
        raise ValueError(f"proj_s is {tuple(step.proj_s.shape)}, student tokens have dimension {d_s}")
    if step.proj_t is not None and tuple(step.proj_t.shape) != (d_s, d_t):
        raise ValueError(f"proj_t is {tuple(step.proj_t.shape)}, expected ({d_s}, {d_t})")


class MixingWeights(torch.autograd.Function):
    @staticmethod
    def forward(ctx, log_temps, step: StepContext, *students):
        students = _contiguous_students(step, students)
        step.teachers = [t.detach().contiguous() for t in step.teachers]
        _check_inputs(students, step)
        logt = log_temps.detach().to(torch.float32).contiguous()
        stats, _ = eng.statistics(students, step.teachers, step.attns, step.has_cls, step.group, step.world)
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
        students = _contiguous_students(step, students)
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
