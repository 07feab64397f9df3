"""the capture helpers either side of the loss, against a plain
restatement of the reference's hooks on a toy ViT (no timm in this image)."""
import types

import pytest
import torch
import torch.nn as nn

from basd_b200 import capture


class Attention(nn.Module):                      # timm-shaped: .qkv, .num_heads, .proj
    def __init__(self, dim, heads):
        super().__init__()
        self.num_heads = heads
        self.qkv = nn.Linear(dim, 3 * dim)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        b, n, c = x.shape
        hd = c // self.num_heads
        qkv = self.qkv(x).reshape(b, n, 3, self.num_heads, hd).permute(2, 0, 3, 1, 4)
        attn = ((qkv[0] @ qkv[1].transpose(-2, -1)) * hd ** -0.5).softmax(-1)
        return self.proj((attn @ qkv[2]).transpose(1, 2).reshape(b, n, c))


class Block(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.attn = Attention(dim, heads)
        self.mlp = nn.Linear(dim, dim)

    def forward(self, x):
        x = x + self.attn(x)
        return x + self.mlp(x)


class ToyViT(nn.Module):
    def __init__(self, dim=32, heads=4, depth=3, classes=5):
        super().__init__()
        self.cls = nn.Parameter(torch.randn(1, 1, dim))
        self.blocks = nn.ModuleList([Block(dim, heads) for _ in range(depth)])
        self.head = nn.Linear(dim, classes)

    def forward(self, x):                        # x: (B, N, dim) "patches"
        x = torch.cat([self.cls.expand(x.shape[0], -1, -1), x], dim=1)
        for blk in self.blocks:
            x = blk(x)
        return self.head(x[:, 0])


def _teacher(model, has_cls=True):
    return types.SimpleNamespace(model=model, layer_paths=[f"blocks.{i}" for i in range(len(model.blocks))],
                                 attn_subpath="attn", has_cls_token=has_cls, feature_format="token")


def test_importance_rows_equal_the_cls_rows_of_the_full_maps():
    torch.manual_seed(0)
    model = ToyViT().eval()
    x = torch.randn(2, 9, 32)
    tok_full, maps = capture.extract_intermediates(_teacher(model), x, full_maps=True)
    tok_rows, rows = capture.extract_intermediates(_teacher(model), x)
    assert sorted(rows) == [0, 1, 2]
    for k in maps:
        assert maps[k].shape == (2, 4, 10, 10) and rows[k].shape == (2, 9)
        ref = maps[k][:, :, 0, 1:].mean(dim=1)                    # relational.py:22-24
        assert torch.allclose(rows[k], ref, atol=1e-6)
        assert torch.equal(tok_full[k], tok_rows[k]) and tok_rows[k].shape == (2, 9, 32)


def test_student_capture_strips_cls_and_returns_logits():
    torch.manual_seed(1)
    model = ToyViT()
    x = torch.randn(3, 9, 32)
    logits, toks = capture.extract_student(model, x, [0, 2], layer_paths=[f"blocks.{i}" for i in range(3)],
                                           has_cls_token=True)
    assert logits.shape == (3, 5) and sorted(toks) == [0, 2]
    assert toks[2].shape == (3, 9, 32) and toks[2].requires_grad


def test_cnn_teacher_gives_one_layer_and_uniform_importance():
    net = types.SimpleNamespace(forward_features=lambda x: x)
    teacher = types.SimpleNamespace(model=net, feature_format="nchw", has_cls_token=False)
    x = torch.randn(2, 16, 7, 7)
    toks, attn = capture.extract_intermediates(teacher, x)
    assert toks[0].shape == (2, 49, 16) and attn[0].shape == (2, 49)
    assert torch.allclose(attn[0], torch.full((2, 49), 1 / 49))
    _, full = capture.extract_intermediates(teacher, x, full_maps=True)
    assert full[0].shape == (2, 1, 49, 49)


@pytest.mark.gpu
def test_intrinsic_dim_matches_the_oracle_rule_on_device():
    from oracle import ref_port as rp
    torch.manual_seed(2)
    model = ToyViT(dim=32, depth=2).cuda().eval()
    images = torch.randn(40, 9, 32, device="cuda") * torch.logspace(0, -1.5, 32, device="cuda")
    got = capture.estimate_intrinsic_dim(_teacher(model), images)
    out = {}
    h = model.blocks[-1].register_forward_hook(lambda m, i, o: out.update(o=o))
    with torch.no_grad():
        model(images)
    h.remove()
    flat = out["o"][:, 1:, :].reshape(-1, 32).float().cpu()
    assert got == rp.mp_rank(flat)
