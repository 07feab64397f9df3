"""Callers either side of the loss: teacher / student
intermediate capture and the calibration-time intrinsic-dimension estimate, with the same
signatures as the reference's helpers.

f1  ``extract_intermediates``    reference: src/models/teacher.py:27-39, 180-216
    The reference recomputes and stores the full softmax(QK^T) map of every teacher layer
    (B, H, N+1, N+1) although the loss reads only its CLS row averaged over the heads
    (relational.py:22-24).  Here the attention hook emits that (B, N) importance row directly --
    one (1 x hd) . (hd x N+1) product per head instead of the whole map -- which ``BASDLoss``
    accepts in place of the map (the 5.7 GB attention stack of C2 is never materialised).
    ``full_maps=True`` restores the reference's output for callers that want the maps.
f2  ``extract_student``          reference: src/training/trainer.py:16-37
f3  ``estimate_intrinsic_dim``   reference: src/models/teacher.py:161-177, on the device MP-rank kernel

These are host-side plumbing (forward hooks); the arithmetic that matters stays in the kernels.
"""
from __future__ import annotations

import torch


def _to_token_format(t: torch.Tensor, feature_format: str, has_cls_token: bool) -> torch.Tensor:
    """reference: src/models/teacher.py:151-158"""
    if feature_format == "nhwc":
        t = t.permute(0, 3, 1, 2).flatten(2).transpose(1, 2)
    elif feature_format == "nchw":
        t = t.flatten(2).transpose(1, 2)
    if has_cls_token:
        t = t[:, 1:, :]
    return t


def make_attn_capture_hook(capture_dict: dict, layer_idx: int, *, apply_softmax: bool = True):
    """Full (B, H, N, N) attention capture -- the reference's hook (teacher.py:27-39)."""
    def hook(mod, inp, out):
        x_in = inp[0]
        b, n, c = x_in.shape
        nh = mod.num_heads
        hd = c // nh
        qkv = mod.qkv(x_in).reshape(b, n, 3, nh, hd).permute(2, 0, 3, 1, 4)
        attn = (qkv[0] @ qkv[1].transpose(-2, -1)) * (hd ** -0.5)
        capture_dict[layer_idx] = attn.softmax(dim=-1) if apply_softmax else attn
    return hook


def make_importance_capture_hook(capture_dict: dict, layer_idx: int, *, has_cls_token: bool = True):
    """Importance row only: mean over heads of the CLS query's softmax row without the CLS column
    (has_cls_token) or the mean over heads and queries (no CLS) -- exactly what
    relational.py:22-27 extracts from the full map."""
    def hook(mod, inp, out):
        x_in = inp[0]
        b, n, c = x_in.shape
        nh = mod.num_heads
        hd = c // nh
        qkv = mod.qkv(x_in).reshape(b, n, 3, nh, hd).permute(2, 0, 3, 1, 4)
        q, k = qkv[0], qkv[1]                                  # (B, H, N, hd)
        if has_cls_token:
            row = (q[:, :, :1, :] @ k.transpose(-2, -1)) * (hd ** -0.5)      # (B, H, 1, N)
            capture_dict[layer_idx] = row.softmax(dim=-1)[:, :, 0, 1:].mean(dim=1)
        else:
            attn = ((q @ k.transpose(-2, -1)) * (hd ** -0.5)).softmax(dim=-1)
            capture_dict[layer_idx] = attn.mean(dim=(1, 2))
    return hook


@torch.no_grad()
def extract_intermediates(teacher, x: torch.Tensor, *, full_maps: bool = False):
    """(tokens per layer, attention per layer) for ``BASDLoss.forward``.  ``teacher`` is the
    reference's ``TeacherModel`` tuple (model, layer_paths, attn_subpath, has_cls_token,
    feature_format).  CNN teachers give one layer of tokens and uniform importance."""
    if teacher.feature_format != "token":
        features = teacher.model.forward_features(x)
        features = _to_token_format(features, teacher.feature_format, teacher.has_cls_token)
        b, n, _ = features.shape
        if full_maps:
            attn = torch.ones(b, 1, n, n, device=features.device, dtype=features.dtype) / n
        else:
            attn = torch.full((b, n), 1.0 / n, device=features.device, dtype=torch.float32)
        return {0: features}, {0: attn}

    hooks, tokens, attns = [], {}, {}
    for idx, path in enumerate(teacher.layer_paths):
        module = teacher.model.get_submodule(path)

        def make_token_hook(i):
            def hook(mod, inp, out):
                tokens[i] = _to_token_format(out, teacher.feature_format, teacher.has_cls_token)
            return hook
        hooks.append(module.register_forward_hook(make_token_hook(idx)))
        if teacher.attn_subpath is not None:
            attn_mod = teacher.model.get_submodule(f"{path}.{teacher.attn_subpath}")
            hook = (make_attn_capture_hook(attns, idx, apply_softmax=True) if full_maps else
                    make_importance_capture_hook(attns, idx, has_cls_token=teacher.has_cls_token))
            hooks.append(attn_mod.register_forward_hook(hook))
    teacher.model(x)
    for h in hooks:
        h.remove()
    return tokens, attns


def extract_student(model: torch.nn.Module, x: torch.Tensor, layer_indices: list[int], *,
                    layer_paths: list[str], has_cls_token: bool):
    """(logits, {layer index: tokens without CLS}) -- reference: trainer.py:16-37.  ``model`` may be
    wrapped (DistributedDataParallel / accelerate): the hooks go on the inner module's blocks, the
    forward runs through the wrapper so that its gradient hooks fire."""
    hooks, captured = [], {}
    inner = model
    while hasattr(inner, "module") and isinstance(getattr(inner, "module"), torch.nn.Module):
        inner = inner.module
    for idx in layer_indices:
        block = inner.get_submodule(layer_paths[idx])

        def make_token_hook(i, _has_cls=has_cls_token):
            def hook(mod, inp, out):
                captured[i] = out[:, 1:, :] if _has_cls else out
            return hook
        hooks.append(block.register_forward_hook(make_token_hook(idx)))
    logits = model(x)
    for h in hooks:
        h.remove()
    return logits, captured


@torch.no_grad()
def estimate_intrinsic_dim(teacher, images: torch.Tensor) -> int:
    """Marchenko-Pastur rank of the teacher's last-layer tokens on calibration images, on the
    device kernel (reference: teacher.py:161-177; needs at least D token rows, which the
    reference's 10 D / tokens-per-image calibration set guarantees)."""
    from .losses import marchenko_pastur_rank
    captured = {}
    mod = teacher.model.get_submodule(teacher.layer_paths[-1])
    h = mod.register_forward_hook(lambda m, i, o: captured.update(out=o))
    teacher.model(images)
    h.remove()
    tokens = _to_token_format(captured["out"], teacher.feature_format, teacher.has_cls_token)
    flat = tokens.reshape(-1, tokens.shape[-1]).float()
    return marchenko_pastur_rank(flat)
