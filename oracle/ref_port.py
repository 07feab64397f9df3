"""CPU oracle: a restatement of the reference BASD loss path (TEST INFRASTRUCTURE ONLY).

This file is the *checker*, never the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  The shipped package (``vit-inductive-bias-distillation_b200``)
never imports anything from ``oracle/``.

It re-expresses, function by function, what the reference's ``src/losses`` computes
(the algorithm, including the LAPACK SVD / eigvalsh calls the reference makes through
``torch.linalg``), so that on the GPU box -- where ``/root/reference`` does not exist --
the CUDA path still has something to be compared with.

Pinning: the reference ships no tests, golden vectors or fixtures.
The port is therefore pinned against *outputs of the reference itself run in the build
container*: ``tests/golden/make_golden.py`` imports ``/root/reference/src/losses`` and
stores inputs seeds + outputs (ranks, mixing weights, loss, gradients) under
``tests/golden/*.pt``; ``tests/test_oracle_golden.py`` replays them through this port.

All arithmetic is fp32 with autocast off (SURVEY.md §5 "Autocast note"); tensors given
in bf16 are upcast exactly where the reference upcasts them.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# a1  marchenko_pastur_rank            reference: src/losses/layer_selector.py:8-20
# --------------------------------------------------------------------------------------
@torch.no_grad()
def mp_rank(features: torch.Tensor) -> int:
    rows, dim = features.shape
    aspect = dim / rows                                   # :11
    if rows >= dim:                                       # :12-15 (uncentred covariance)
        second_moment = features.T @ features / rows
    else:
        second_moment = features @ features.T / rows
    spectrum = torch.linalg.eigvalsh(second_moment)       # :16 ascending
    noise_level = spectrum.median().item()                # :17 torch.median = lower middle
    edge = noise_level * (1.0 + aspect ** 0.5) ** 2       # :18 MP upper edge
    return int((spectrum > edge).sum().item())            # :19 strict >


# --------------------------------------------------------------------------------------
# a2  _grassmann_subspace              reference: src/losses/layer_selector.py:23-37
# --------------------------------------------------------------------------------------
def top_subspace(z_flat: torch.Tensor, k: int):
    zc = z_flat.float()                                   # :34
    zc = zc - zc.mean(dim=0, keepdim=True)                # :35
    _, svals, vt = torch.linalg.svd(zc, full_matrices=False)   # :36
    return vt[:k].T, svals[:k]                            # :37  (D,k), (k,)


def initial_log_temperature() -> float:
    """softplus^-1(1.0)                 reference: src/losses/layer_selector.py:58-63"""
    return math.log(math.e - 1.0)


def make_selector_state(num_points: int, student_dim: int, teacher_dim: int):
    """Same RNG consumption order as    reference: src/losses/layer_selector.py:51-63"""
    proj_s = torch.empty(student_dim, student_dim)
    proj_t = torch.empty(student_dim, teacher_dim)
    torch.nn.init.orthogonal_(proj_s)
    torch.nn.init.orthogonal_(proj_t)
    log_t = torch.full((num_points,), initial_log_temperature())
    return proj_s, proj_t, log_t


def extraction_layers(student_depth: int, num_points: int) -> list[int]:
    """reference: src/losses/combined.py:34-40 (python round = banker's rounding)"""
    if num_points == 1:
        return [student_depth - 1]
    return [round(i * (student_depth - 1) / (num_points - 1)) for i in range(num_points)]


# --------------------------------------------------------------------------------------
# a4-a6  selector                      reference: src/losses/layer_selector.py:69-152
# --------------------------------------------------------------------------------------
def selector_forward(student_tokens, teacher_tokens, teacher_attns, layers,
                     proj_s, proj_t, log_temps, ranks_override=None):
    """Returns mixed tokens/attn per student layer plus the diagnostics the parity tests
    compare (ranks, distances, mixing weights, principal-angle cosines).

    `ranks_override` (test hook, not in the reference): evaluate everything downstream of the
    Marchenko-Pastur decision at the given ranks.  The rank is a discontinuous integer; when an
    eigenvalue sits within 1e-4 of the threshold the kernels may land on the other side of it, and
    the parity tests then compare against the reference algorithm run AT THE KERNEL'S RANK."""
    t_keys = sorted(teacher_tokens.keys())                # :123
    d_s = proj_s.shape[0]
    d_t = teacher_tokens[t_keys[0]].shape[2]

    ranks = {}
    with torch.no_grad():                                 # :69-74
        for key in t_keys:
            z = teacher_tokens[key].reshape(-1, d_t) @ proj_t.T
            ranks[key] = min(mp_rank(z), d_s - 1)
    if ranks_override is not None:
        ranks = {key: int(ranks_override[j]) for j, key in enumerate(t_keys)}

    tok_stack = torch.stack([teacher_tokens[k] for k in t_keys])       # :128
    att_stack = torch.stack([teacher_attns[k] for k in t_keys])        # :129

    bases, spectral = {}, {}
    with torch.no_grad():                                 # :131-138
        for key in t_keys:
            z = teacher_tokens[key].reshape(-1, d_t) @ proj_t.T
            bases[key], spectral[key] = top_subspace(z, ranks[key])

    temps = F.softplus(log_temps)                         # :65-67
    out_tok, out_att = {}, {}
    diag = SimpleNamespace(ranks=ranks, dist={}, weights={}, cosines={})
    for i, layer in enumerate(layers):                    # :143
        s = student_tokens[layer]
        zs = s.reshape(-1, s.shape[2]) @ proj_s.T         # :86-88
        zs = zs.float()
        zs = zs - zs.mean(dim=0, keepdim=True)            # :90-91
        _, _, vt_s = torch.linalg.svd(zs, full_matrices=False)   # :92
        dist = torch.zeros(len(t_keys))
        cos_all = []
        for j, key in enumerate(t_keys):                  # :95-105
            k = ranks[key]
            overlap = vt_s[:k] @ bases[key]               # == U_s^T U_t
            cosines = torch.linalg.svdvals(overlap)       # :99
            angles = torch.acos(cosines.clamp(max=1.0 - torch.finfo(cosines.dtype).eps))  # :100
            sw = spectral[key]
            dist[j] = (sw * angles.pow(2)).sum() / sw.sum()       # :105
            cos_all.append(cosines.detach())
        mix = F.softmax(-dist / temps[i], dim=0)          # :107-108
        diag.dist[layer] = dist.detach()
        diag.weights[layer] = mix.detach()
        diag.cosines[layer] = cos_all
        mix_cast = mix.to(tok_stack.dtype)                # :110
        out_tok[layer] = (mix_cast.view(-1, 1, 1, 1) * tok_stack).sum(dim=0)          # :111
        out_att[layer] = (mix_cast.view(-1, 1, 1, 1, 1) * att_stack).sum(dim=0)       # :112
    return out_tok, out_att, diag


# --------------------------------------------------------------------------------------
# a7  _align_token_count               reference: src/losses/combined.py:9-14
# --------------------------------------------------------------------------------------
def align_tokens(tokens: torch.Tensor, target_n: int) -> torch.Tensor:
    if tokens.shape[1] == target_n:
        return tokens
    return F.interpolate(tokens.transpose(1, 2), size=target_n, mode="linear",
                         align_corners=False).transpose(1, 2)


# --------------------------------------------------------------------------------------
# a8  geometric_relational_loss        reference: src/losses/relational.py:5-50
# --------------------------------------------------------------------------------------
def token_importance(attn: torch.Tensor, n_student: int, has_cls: bool) -> torch.Tensor:
    if has_cls:
        w = attn[:, :, 0, 1:].mean(dim=1)                 # :22-24
    else:
        w = attn.mean(dim=(1, 2))                         # :25-27
    if w.shape[1] != n_student:                           # :29-32
        w = F.interpolate(w.unsqueeze(1), size=n_student, mode="linear",
                          align_corners=False).squeeze(1)
    return w / w.sum(dim=-1, keepdim=True)                # :34


def procrustes_loss(student, teacher, attn, has_cls: bool) -> torch.Tensor:
    s = student.float()                                   # :18-19
    t = teacher.float()
    w = token_importance(attn, s.shape[1], has_cls)
    s = s - (w.unsqueeze(-1) * s).sum(dim=1, keepdim=True)        # :36-39
    t = t - (w.unsqueeze(-1) * t).sum(dim=1, keepdim=True)
    root = w.unsqueeze(-1).sqrt()                         # :41-43
    s = root * s
    t = root * t
    energy = (s * s).sum(dim=(1, 2)) + (t * t).sum(dim=(1, 2))    # :45-46
    cross = torch.bmm(s.transpose(1, 2), t)               # :47
    nuclear = torch.linalg.matrix_norm(cross, ord="nuc")  # :48
    return (energy - 2.0 * nuclear).mean()                # :50


# --------------------------------------------------------------------------------------
# a10  BASDLoss.forward                reference: src/losses/combined.py:48-85
# --------------------------------------------------------------------------------------
def uwso(values: list[torch.Tensor]) -> torch.Tensor:
    eps = torch.finfo(values[0].dtype).eps                # :81
    inv = torch.stack([1.0 / v.detach().clamp(min=eps) for v in values])   # :82
    share = inv / inv.sum()                               # :83
    return sum(share[i] * values[i] for i in range(len(values)))          # :85


def basd_forward(logits, targets, student_tokens, teacher_tokens, teacher_attns, *,
                 layers, proj_s, proj_t, log_temps, n_student_tokens, has_cls,
                 criterion, ranks_override=None):
    """Full loss. Returns (loss, details)."""
    ce = criterion(logits, targets)                       # :56
    mixed_tok, mixed_att, diag = selector_forward(        # :58-61
        student_tokens, teacher_tokens, teacher_attns, layers, proj_s, proj_t, log_temps,
        ranks_override=ranks_override)
    geo_terms = []
    for layer in layers:                                  # :63-75
        aligned = align_tokens(mixed_tok[layer], n_student_tokens)
        geo_terms.append(procrustes_loss(student_tokens[layer], aligned,
                                         mixed_att[layer], has_cls))
    geo = torch.stack(geo_terms).mean()                   # :76
    diag.ce = ce.detach()
    diag.geo = geo.detach()
    diag.geo_terms = [g.detach() for g in geo_terms]
    return uwso([ce, geo]), diag                          # :78-85


class OracleBASD(torch.nn.Module):
    """Module wrapper with the reference's constructor/forward signature
    (reference: src/losses/combined.py:17-55) so benches can time it like the original."""

    def __init__(self, base_criterion, student_dim, teacher_dim, student_depth,
                 num_student_tokens, *, config, teacher_has_cls_token):
        super().__init__()
        self.base_criterion = base_criterion
        self.has_cls = teacher_has_cls_token
        self.num_student_tokens = num_student_tokens
        self.token_layers = extraction_layers(student_depth, config.num_extraction_points)
        proj_s, proj_t, log_t = make_selector_state(len(self.token_layers), student_dim,
                                                    teacher_dim)
        self.register_buffer("proj_s", proj_s)
        self.register_buffer("proj_t", proj_t)
        self.log_temperatures = torch.nn.Parameter(log_t)
        self.last = None

    def forward(self, logits, targets, student_tokens, teacher_tokens, teacher_attns):
        loss, self.last = basd_forward(
            logits, targets, student_tokens, teacher_tokens, teacher_attns,
            layers=self.token_layers, proj_s=self.proj_s, proj_t=self.proj_t,
            log_temps=self.log_temperatures, n_student_tokens=self.num_student_tokens,
            has_cls=self.has_cls, criterion=self.base_criterion)
        return loss
