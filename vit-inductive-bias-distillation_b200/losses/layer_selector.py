"""Grassmannian layer selector on the B200 kernels.

Interface mirror of the reference's ``src/losses/layer_selector.py`` (class at :40-152,
``marchenko_pastur_rank`` at :8-20): same constructor, buffers (``proj_s``, ``proj_t``),
parameter (``log_temperatures``), ``temperatures`` property, ``subspace_ranks`` side effect
and ``forward`` signature.  The arithmetic runs in libbasd_b200.so.
"""
from __future__ import annotations

import ctypes
import math
import types

import torch
import torch.nn as nn

from .. import _engine as eng
from .._autograd import MixingWeights, StepContext, world_size
from .._native import call, dtype_code, ptr, stream


@torch.no_grad()
def marchenko_pastur_rank(features: torch.Tensor) -> int:
    """MP rank of an (M, D) feature matrix on the device kernels (reference :8-20).

    Returns a python int like the reference, which costs one host sync here (the loss's
    own forward keeps the ranks on the device and never syncs)."""
    if features.dim() != 2:
        raise ValueError("features must be (M, D)")
    rows, dim = features.shape
    x = features.contiguous()
    gram = torch.empty(1, dim, dim, dtype=torch.float32, device=x.device)
    col = torch.empty(dim, dtype=torch.float32, device=x.device)
    eng.token_gram(x, gram[0], col)
    k = torch.empty_like(gram)
    call("basd_center_gram", ptr(gram), None, dim, 0.0, ptr(k), 1, stream())
    lam, _ = eng.sym_eig(k)
    rank = torch.empty(1, dtype=torch.int32, device=x.device)
    call("basd_mp_rank", ptr(lam), dim, rows, dim, ptr(rank), None, 1, stream())
    return int(rank.item())


class _LazyRanks(dict):
    """``subspace_ranks`` with the reference's dict interface; the values stay on the GPU
    until somebody reads them (the reference syncs 2 x L_t times per step to fill it)."""

    def __init__(self):
        super().__init__()
        self._pending = None

    def _stage(self, keys, ranks):
        self._pending = (list(keys), ranks)

    def _sync(self):
        if self._pending is not None:
            keys, ranks = self._pending
            self._pending = None
            for k, v in zip(keys, ranks.tolist()):
                super().__setitem__(k, int(v))

    def __getitem__(self, k):
        self._sync()
        return super().__getitem__(k)

    def __iter__(self):
        self._sync()
        return super().__iter__()

    def __len__(self):
        self._sync()
        return super().__len__()

    def __contains__(self, k):
        self._sync()
        return super().__contains__(k)

    def __eq__(self, other):
        self._sync()
        return dict(self.items()) == other

    def __repr__(self):
        self._sync()
        return super().__repr__()

    def keys(self):
        self._sync()
        return super().keys()

    def values(self):
        self._sync()
        return super().values()

    def items(self):
        self._sync()
        return super().items()

    def get(self, k, default=None):
        self._sync()
        return super().get(k, default)


class GrassmannianLayerSelector(nn.Module):
    def __init__(self, num_extraction_points: int, student_dim: int, teacher_dim: int):
        super().__init__()
        self.student_dim = student_dim
        self.subspace_ranks = _LazyRanks()
        # same RNG consumption order as the reference (:51-56)
        proj_s = torch.empty(student_dim, student_dim)
        proj_t = torch.empty(student_dim, teacher_dim)
        nn.init.orthogonal_(proj_s)
        nn.init.orthogonal_(proj_t)
        self.register_buffer("proj_s", proj_s)
        self.register_buffer("proj_t", proj_t)
        self.log_temperatures = nn.Parameter(
            torch.full((num_extraction_points,), math.log(math.exp(1.0) - 1)))
        self.process_group = None
        self.sync_stats = True
        self.last_state = None           # small diagnostics of the last step (ranks, edges, dist, weights, ...)

    @property
    def temperatures(self) -> torch.Tensor:
        return torch.nn.functional.softplus(self.log_temperatures)

    def _step(self, teacher_tokens, teacher_attns, has_cls, n_student) -> StepContext:
        keys = sorted(teacher_tokens.keys())
        if set(teacher_attns.keys()) != set(keys):
            raise ValueError(f"teacher token layers {keys} and attention maps "
                             f"{sorted(teacher_attns.keys())} have different keys")
        world = world_size(self.process_group) if self.sync_stats else 1
        return StepContext(
            teachers=[teacher_tokens[k] for k in keys], attns=[teacher_attns[k] for k in keys],
            has_cls=has_cls, n_student=n_student, proj_s=self.proj_s.float().contiguous(),
            proj_t=self.proj_t.float().contiguous(), group=self.process_group, world=world)

    def mixing_weights(self, student_tokens_per_layer, all_teacher_tokens, all_teacher_attns,
                       extraction_indices, *, has_cls=True, n_student=None):
        """(E, L_t) softmax mixing weights, differentiable w.r.t. the student tokens and
        ``log_temperatures``; also returns the StepContext the Procrustes stage reuses."""
        students = [student_tokens_per_layer[l] for l in extraction_indices]
        if n_student is None:
            n_student = students[0].shape[1]
        step = self._step(all_teacher_tokens, all_teacher_attns, has_cls, n_student)
        weights = MixingWeights.apply(self.log_temperatures, step, *students)
        self.subspace_ranks._stage(sorted(all_teacher_tokens.keys()), step.selector.ranks)
        # keep only the small diagnostics alive between steps: the step context (contiguous teacher
        # copies, Procrustes operators, selector factors: ~3 GB at C2) dies with the autograd graph
        sel = step.selector
        self.last_state = types.SimpleNamespace(
            ranks=sel.ranks, edges=sel.edges, dist=sel.dist, weights=sel.weights, temps=sel.temps,
            sweeps=dict(sel.sweeps))
        return weights, step

    @torch.no_grad()
    def mixed_importance(self, weights, step, n_out=None):
        """The importance-row form of the mixed attention (what relational.py:22-34 reduces the mixed map
        to): (E, B, n_out) rows, sum_l w[i,l] row_l resampled to n_out tokens and normalised per sample."""
        rows = step.stats.rows
        l, b, n_rows = rows.shape
        w = weights.detach().float().contiguous()
        e = w.shape[0]
        n_out = n_rows if n_out is None else n_out
        out = torch.empty(e, b, n_out, dtype=torch.float32, device=rows.device)
        totals = torch.empty(e, b, dtype=torch.float32, device=rows.device)
        call("basd_mix_rows", ptr(rows), ptr(w), e, l, b, n_rows, n_out, ptr(out), ptr(totals), stream())
        return out

    def forward(self, student_tokens_per_layer, all_teacher_tokens, all_teacher_attns,
                extraction_indices):
        """Reference-shaped entry (:116-152): dicts of mixed teacher tokens (B,N_t,D_t) and
        mixed attention maps keyed by student layer.  The weights come from the kernels;
        materialising the full mixed maps is only done here, for API compatibility --
        BASDLoss never does it (it mixes the (B, N) importance rows: `mixed_importance`)."""
        # the mixing weights do not depend on the attention maps; the flag only tells the statistics
        # stage how to read them: (B,H,N_t+1,N_t+1) maps carry a CLS row, (B,1,N_t,N_t) maps do not
        keys = sorted(all_teacher_tokens.keys())
        a0, n_t = all_teacher_attns[keys[0]], all_teacher_tokens[keys[0]].shape[1]
        has_cls = not (a0.dim() == 4 and a0.shape[-1] == n_t)
        weights, step = self.mixing_weights(student_tokens_per_layer, all_teacher_tokens,
                                            all_teacher_attns, extraction_indices, has_cls=has_cls)
        # one pass over the teacher stack / the attention stack for all E outputs (mix.cu); gradients
        # flow through `weights` only inside BASDLoss (ProcrustesGeo), these dicts are detached
        w = weights.detach().float().contiguous()
        e, l = w.shape
        toks = step.teachers
        b, n_t, d_t = toks[0].shape
        mixed = torch.empty(e, b, n_t, d_t, dtype=toks[0].dtype, device=w.device)
        tptrs = (ctypes.c_void_p * l)(*[t.data_ptr() for t in toks])
        call("basd_mix_interp", tptrs, l, e, ptr(w), dtype_code(toks[0]), b, n_t, n_t, d_t, ptr(mixed),
             dtype_code(mixed), stream())
        atts = [all_teacher_attns[k].contiguous() for k in keys]
        if any(a.shape != atts[0].shape or a.dtype != atts[0].dtype for a in atts):
            raise ValueError("teacher attention maps must share one shape and dtype")
        mixed_a = torch.empty((e,) + tuple(atts[0].shape), dtype=atts[0].dtype, device=w.device)
        aptrs = (ctypes.c_void_p * l)(*[a.data_ptr() for a in atts])
        call("basd_mix_flat", aptrs, l, e, ptr(w), dtype_code(atts[0]), atts[0].numel(), ptr(mixed_a), stream())
        mixed_tok = {layer: mixed[i] for i, layer in enumerate(extraction_indices)}
        mixed_att = {layer: mixed_a[i] for i, layer in enumerate(extraction_indices)}
        return mixed_tok, mixed_att
