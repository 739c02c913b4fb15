"""Host-side mirror of the model interface the reference builds its student / teacher from.

The reference calls ``timm.create_model("vit_small_patch16_224" | "vit_base_patch16_224", pretrained=False,
num_classes=...)`` (ref/src/models/model_registry.py:167-172, 228-233) and wraps the student in ``QATWrapper``
(ref/src/models/model_registry.py:99-124).  timm is not installed on the B200 image, so this module provides the same
module tree (names, types, shapes -- SURVEY.md App. B) as plain ``nn.Module`` parameter containers, so that
``get_default_qat_qconfig`` + ``prepare_qat`` + ``state_dict()`` + ``convert()`` see exactly what they see upstream.
``install_timm_shim()`` registers it as ``timm`` so the reference's own ``model_registry`` imports unmodified.

The forward methods below are ordinary PyTorch (they make the modules usable anywhere); the B200 hot path does NOT
call them -- ``qatvit_b200.engine`` reads the parameters / observer buffers and runs its own CUDA kernels.
"""
from __future__ import annotations

import importlib.machinery
import sys
import types
from functools import partial

import torch
import torch.nn as nn
import torch.nn.functional as F


class TimmLayerNorm(nn.LayerNorm):
    """Stand-in for ``timm.layers.LayerNorm`` (a subclass of nn.LayerNorm): torch.ao's qconfig
    propagation matches on the exact type, so this variant is NOT observed (101 fake-quant modules);
    plain ``nn.LayerNorm`` (older timm) is observed (126).  SURVEY.md §0.6."""

    def forward(self, x):
        return F.layer_norm(x, self.normalized_shape, self.weight, self.bias, self.eps)


class PatchEmbed(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768):
        super().__init__()
        self.img_size = (img_size, img_size)
        self.patch_size = (patch_size, patch_size)
        self.grid_size = (img_size // patch_size, img_size // patch_size)
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size, bias=True)
        self.norm = nn.Identity()

    def forward(self, x):
        x = self.proj(x)
        x = x.flatten(2).transpose(1, 2)
        return self.norm(x)


class Attention(nn.Module):
    def __init__(self, dim, num_heads):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.q_norm = nn.Identity()
        self.k_norm = nn.Identity()
        self.attn_drop = nn.Dropout(0.0)
        self.norm = nn.Identity()
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(0.0)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, self.head_dim).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        q, k = self.q_norm(q), self.k_norm(k)
        x = F.scaled_dot_product_attention(q, k, v, dropout_p=0.0)
        x = x.transpose(1, 2).reshape(B, N, C)
        x = self.norm(x)
        x = self.proj(x)
        return self.proj_drop(x)


class Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.drop1 = nn.Dropout(0.0)
        self.norm = nn.Identity()
        self.fc2 = nn.Linear(hidden, dim)
        self.drop2 = nn.Dropout(0.0)

    def forward(self, x):
        return self.drop2(self.fc2(self.norm(self.drop1(self.act(self.fc1(x))))))


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio, norm_layer):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads)
        self.ls1 = nn.Identity()
        self.drop_path1 = nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))
        self.ls2 = nn.Identity()
        self.drop_path2 = nn.Identity()

    def forward(self, x):
        x = x + self.drop_path1(self.ls1(self.attn(self.norm1(x))))
        x = x + self.drop_path2(self.ls2(self.mlp(self.norm2(x))))
        return x


class VisionTransformer(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=10, embed_dim=768, depth=12,
                 num_heads=12, mlp_ratio=4.0, ln_variant="subclass"):
        super().__init__()
        norm_layer = partial(TimmLayerNorm if ln_variant == "subclass" else nn.LayerNorm, eps=1e-6)
        self.num_classes = num_classes
        self.embed_dim = self.num_features = embed_dim
        self.patch_embed = PatchEmbed(img_size, patch_size, in_chans, embed_dim)
        n = self.patch_embed.num_patches
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.randn(1, n + 1, embed_dim) * 0.02)
        self.pos_drop = nn.Dropout(0.0)
        self.patch_drop = nn.Identity()
        self.norm_pre = nn.Identity()
        self.blocks = nn.Sequential(*[Block(embed_dim, num_heads, mlp_ratio, norm_layer) for _ in range(depth)])
        self.norm = norm_layer(embed_dim)
        self.fc_norm = nn.Identity()
        self.head_drop = nn.Dropout(0.0)
        self.head = nn.Linear(embed_dim, num_classes)
        self._init_weights()

    def _init_weights(self):
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        nn.init.normal_(self.cls_token, std=1e-6)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    def forward_features(self, x):
        x = self.patch_embed(x)
        x = torch.cat([self.cls_token.expand(x.shape[0], -1, -1), x], dim=1)
        x = x + self.pos_embed
        x = self.pos_drop(x)
        x = self.norm_pre(self.patch_drop(x))
        x = self.blocks(x)
        return self.norm(x)

    def forward_head(self, x):
        x = x[:, 0]
        x = self.head_drop(self.fc_norm(x))
        return self.head(x)

    def forward(self, x):
        return self.forward_head(self.forward_features(x))


_CFG = {
    "vit_small_patch16_224": dict(embed_dim=384, depth=12, num_heads=6),
    "vit_base_patch16_224": dict(embed_dim=768, depth=12, num_heads=12),
    # small shapes for fast tests (not timm names; same structure)
    "vit_test_tiny": dict(embed_dim=128, depth=2, num_heads=2),
    "vit_test_teacher": dict(embed_dim=256, depth=2, num_heads=4),
}


def create_model(name: str, pretrained: bool = False, num_classes: int = 10, ln_variant: str = "subclass", **kw):
    """The ``timm.create_model`` surface the reference uses (model_registry.py:167-172,228-233)."""
    if pretrained:
        raise RuntimeError("no network: pretrained weights unavailable")
    cfg = dict(_CFG[name])
    cfg.update(kw)
    return VisionTransformer(num_classes=num_classes, ln_variant=ln_variant, **cfg)


class QATWrapper(nn.Module):
    """Same contract as ref/src/models/model_registry.py:99-124 (classification branch): quant -> model -> dequant."""

    def __init__(self, model: nn.Module, task: str = "classification"):
        super().__init__()
        from torch.ao.quantization import DeQuantStub, QuantStub
        self.quant = QuantStub()
        self.model = model
        self.dequant = DeQuantStub()
        self.task = task

    def forward(self, x, **kwargs):
        return self.dequant(self.model(self.quant(x)))

    def fuse_model(self) -> None:
        return


def install_timm_shim() -> None:
    """Make ``import timm; timm.create_model(...)`` resolve to this module (SURVEY.md App. C recipe)."""
    if "timm" in sys.modules:
        return
    shim = types.ModuleType("timm")
    shim.__version__ = "0.0-qatvit-b200-shim"
    shim.__spec__ = importlib.machinery.ModuleSpec("timm", None)
    shim.create_model = create_model
    sys.modules["timm"] = shim
