"""Synthetic inputs captured from random-init backbones (benchmark / test input generation, next to
``synthetic.py``; models themselves are out of scope).  BASELINE.json asks for
"synthetic ImageNet-shaped features from random-init DeiT/ViT/ResNet backbones": this file builds
timm-shaped ViTs (``blocks.{i}.attn.qkv``, ``num_heads``; timm itself is not in the image) with
timm's initialisation (trunc-normal 0.02 weights, zero biases) and wraps torchvision's ResNet-50
behind ``forward_features``; random images are run through them and tokens / attention are captured
with ``capture.py`` exactly as the reference's trainer does (src/training/trainer.py:16-37,
src/models/teacher.py:180-216).
"""
import math
import types

import torch
import torch.nn as nn

from . import capture

VIT = {                       # name: (dim, depth, heads)
    "deit_tiny": (192, 12, 3),
    "deit_small": (384, 12, 6),
    "deit_base": (768, 12, 12),
    "vit_large": (1024, 24, 16),
}


class Attention(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.num_heads = heads
        self.qkv = nn.Linear(dim, 3 * dim)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        b, n, c = x.shape
        hd = c // self.num_heads
        qkv = self.qkv(x).reshape(b, n, 3, self.num_heads, hd).permute(2, 0, 3, 1, 4)
        out = nn.functional.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2])
        return self.proj(out.transpose(1, 2).reshape(b, n, c))


class Mlp(nn.Module):
    def __init__(self, dim, ratio=4):
        super().__init__()
        self.fc1 = nn.Linear(dim, ratio * dim)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(ratio * dim, dim)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class Block(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = Attention(dim, heads)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = Mlp(dim)

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        return x + self.mlp(self.norm2(x))


class VisionTransformer(nn.Module):
    """Pre-norm ViT/DeiT, patch 16, CLS token, learned positions."""

    def __init__(self, dim, depth, heads, img=224, patch=16, classes=1000):
        super().__init__()
        self.embed_dim = dim
        self.patch_embed = nn.Conv2d(3, dim, patch, patch)
        n = (img // patch) ** 2
        self.cls_token = nn.Parameter(torch.zeros(1, 1, dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, n + 1, dim))
        self.blocks = nn.ModuleList([Block(dim, heads) for _ in range(depth)])
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        self.head = nn.Linear(dim, classes)
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        nn.init.normal_(self.cls_token, std=1e-6)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                nn.init.zeros_(m.bias)

    def forward(self, x):
        x = self.patch_embed(x).flatten(2).transpose(1, 2)
        x = torch.cat([self.cls_token.expand(x.shape[0], -1, -1), x], dim=1) + self.pos_embed
        for blk in self.blocks:
            x = blk(x)
        return self.head(self.norm(x)[:, 0])


class ResNetFeatures(nn.Module):
    """torchvision ResNet-50 (random init) with timm's ``forward_features``: (B, 2048, 7, 7)."""

    def __init__(self):
        super().__init__()
        import torchvision
        net = torchvision.models.resnet50(weights=None)
        self.body = nn.Sequential(net.conv1, net.bn1, net.relu, net.maxpool, net.layer1, net.layer2,
                                  net.layer3, net.layer4)

    def forward_features(self, x):
        return self.body(x)


def fan_in_init(model: nn.Module) -> None:
    """The reference re-initialises every STUDENT it builds (train.py:19-32, applied at :41-56): linear
    weights truncated-normal with std sqrt(2 / fan_in) and zero bias, LayerNorm to (1, 0), convolutions
    normal with std sqrt(2 / fan_out).  Teachers keep their library initialisation."""
    for m in model.modules():
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=math.sqrt(2.0 / m.in_features))
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        elif isinstance(m, nn.LayerNorm):
            nn.init.ones_(m.weight)
            nn.init.zeros_(m.bias)
        elif isinstance(m, nn.Conv2d):
            fan_out = m.kernel_size[0] * m.kernel_size[1] * m.out_channels // m.groups
            nn.init.normal_(m.weight, std=math.sqrt(2.0 / fan_out))
            if m.bias is not None:
                nn.init.zeros_(m.bias)


def vit(name, seed, classes=1000, img=224, patch=16, student=False):
    torch.manual_seed(seed)
    model = VisionTransformer(*VIT[name], img=img, patch=patch, classes=classes)
    if student:
        fan_in_init(model)
    return model.eval()


def vit_teacher(model):
    return types.SimpleNamespace(model=model, layer_paths=[f"blocks.{i}" for i in range(len(model.blocks))],
                                 attn_subpath="attn", has_cls_token=True, feature_format="token")


def cnn_teacher(seed):
    torch.manual_seed(seed)
    return types.SimpleNamespace(model=ResNetFeatures().eval(), layer_paths=[], attn_subpath=None,
                                 has_cls_token=False, feature_format="nchw")


@torch.no_grad()
def backbone_inputs(student_name, teacher_name, layers, batch, *, seed=0, device="cpu", classes=1000,
                    importance_rows=False, img=224, patch=16, chunk=32, model_seed=0):
    """(logits, targets, student tokens, teacher tokens, teacher attention) captured from
    random-init backbones on ``batch`` random img x img images (processed ``chunk`` at a time).
    ``teacher_name`` = "resnet50" gives the CNN teacher (one layer of 49 tokens at 224 x 224, uniform
    attention).  ``seed`` draws the images and targets, ``model_seed`` the weights (data-parallel
    ranks share the models and differ in their images)."""
    gen = torch.Generator().manual_seed(seed)
    student = vit(student_name, 100 + model_seed, classes, img, patch, student=True).to(device)
    paths = [f"blocks.{i}" for i in range(len(student.blocks))]
    if teacher_name == "resnet50":
        teacher = cnn_teacher(200 + model_seed)
        teacher.model.to(device)
    else:
        teacher = vit_teacher(vit(teacher_name, 200 + model_seed, classes, img, patch).to(device))
    parts = []
    for lo in range(0, batch, chunk):
        images = torch.randn(min(chunk, batch - lo), 3, img, img, generator=gen).to(device)
        logits, st = capture.extract_student(student, images, list(layers), layer_paths=paths,
                                             has_cls_token=True)
        te, at = capture.extract_intermediates(teacher, images, full_maps=not importance_rows)
        parts.append((logits, st, te, at))
    cat = lambda dicts: {k: torch.cat([d[k] for d in dicts]).contiguous() for k in dicts[0]}
    targets = torch.randint(0, classes, (batch,), generator=gen).to(device)
    return (torch.cat([p[0] for p in parts]), targets, cat([p[1] for p in parts]),
            cat([p[2] for p in parts]), cat([p[3] for p in parts]))


# workload key -> (student, teacher, image size, patch): the backbones behind SURVEY section 8's C1-C4
WORKLOAD_BACKBONES = {
    "c1": ("deit_tiny", "deit_small", 32, 4),
    "c2": ("deit_small", "deit_base", 224, 16),
    "c3": ("deit_small", "resnet50", 224, 16),
    "c4": ("deit_base", "vit_large", 224, 16),
}


def workload_inputs(key, work, *, seed=0, device="cuda"):
    """Inputs of workload ``key`` (``synthetic.WORKLOADS``) at ``work.batch`` from random-init
    backbones, tokens cast to the workload's token dtype."""
    from .synthetic import _extraction_layers
    student, teacher, img, patch = WORKLOAD_BACKBONES[key]
    layers = _extraction_layers(work.student_depth, work.num_points)
    logits, targets, st, te, at = backbone_inputs(student, teacher, layers, work.batch, seed=seed,
                                                  device=device, classes=work.num_classes, img=img,
                                                  patch=patch)
    assert st[layers[0]].shape[1:] == (work.n_student, work.d_student), st[layers[0]].shape
    assert te[0].shape[1:] == (work.n_teacher, work.d_teacher), te[0].shape
    st = {k: v.to(work.token_dtype) for k, v in st.items()}
    te = {k: v.to(work.token_dtype) for k, v in te.items()}
    at = {k: v.to(work.attn_dtype) for k, v in at.items()}
    return logits.float(), targets, st, te, at
