"""oracle/vit_ref.py -- TEST INFRASTRUCTURE ONLY (checker + reported CPU baseline; never the product path).

CPU restatement of the reference's QAT-distillation step, built from STOCK torch ops:

* the timm ``VisionTransformer`` the reference instantiates through ``timm.create_model``
  (ref/src/models/model_registry.py:167-172, 228-233; structure: SURVEY.md App. B).  timm is not
  installed anywhere this repo runs, so only its *structure* (module names, types, shapes) is
  restated here -- every piece of arithmetic is a live ATen op (``F.layer_norm``, ``F.linear``,
  ``F.scaled_dot_product_attention``, ``F.gelu``), i.e. the same kernels the reference executes;
* ``QATWrapper`` (ref/src/models/model_registry.py:99-124): imported from /root/reference through a
  ``timm`` shim when that tree exists (this container), otherwise the 10-line restatement below;
* the QAT enable block (ref/src/training/qat_trainer.py:300-316) and the train-step body
  (ref/src/training/qat_trainer.py:333-364) which live inline in ``main()`` upstream.

The fake-quant arithmetic itself is torch.ao's (``FusedMovingAvgObsFakeQuantize`` ->
``torch.fused_moving_avg_obs_fake_quant``), exactly as the reference gets it from ``prepare_qat``.
"""
from __future__ import annotations

import importlib.machinery
import os
import sys
import types
from functools import partial
from typing import Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

REFERENCE_ROOT = "/root/reference"

# ref/src/training/qat_trainer.py:36-46
DEFAULT_HPARAMS = {
    "lr": 1.5e-4,
    "weight_decay": 1e-3,
    "label_smoothing": 0.1,
    "kd_temp": 4.0,
    "kd_alpha": 0.5,
    "qat_start_epoch": 2,
    "epochs": 10,
    "batch_size": 256,
    "qat_backend": "qnnpack",
}


# ----------------------------------------------------------------------------------------------
# timm VisionTransformer, structure only
# ----------------------------------------------------------------------------------------------
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
    # small shapes for fast CPU tests (not timm names; same structure)
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


# ----------------------------------------------------------------------------------------------
# QATWrapper: the reference's own class when /root/reference is present, else a restatement
# ----------------------------------------------------------------------------------------------
def install_timm_shim() -> None:
    if "timm" in sys.modules:
        return
    shim = types.ModuleType("timm")
    shim.__version__ = "0.0-shim"
    shim.__spec__ = importlib.machinery.ModuleSpec("timm", None)
    shim.create_model = create_model
    sys.modules["timm"] = shim


def load_reference_registry():
    """Import ref/src/models/model_registry.py unmodified (SURVEY.md App. C recipe) or return None."""
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "models")):
        return None
    install_timm_shim()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from src.models import model_registry  # type: ignore
    return model_registry


class _QATWrapperRestated(nn.Module):
    """ref/src/models/model_registry.py:99-124, classification branch."""

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


def qat_wrapper_cls(prefer_reference: bool = True):
    if prefer_reference:
        reg = load_reference_registry()
        if reg is not None:
            return reg.QATWrapper
    return _QATWrapperRestated


def make_student(name="vit_small_patch16_224", num_classes=10, seed=0, ln_variant="subclass",
                 prefer_reference=True) -> nn.Module:
    """create_student('vit', qat_wrapper=True) -- model_registry.py:384-396."""
    torch.manual_seed(seed)
    return qat_wrapper_cls(prefer_reference)(create_model(name, num_classes=num_classes, ln_variant=ln_variant))


def make_teacher(name="vit_base_patch16_224", num_classes=10, seed=1) -> nn.Module:
    """create_teacher('vit') with random-init weights (no network) -- qat_trainer.py:257-260."""
    torch.manual_seed(seed)
    t = create_model(name, num_classes=num_classes)
    # a random-init ViT gives near-zero logits; widen the head so the KL term is exercised
    with torch.no_grad():
        t.head.weight.mul_(25.0)
    t.eval()
    for p in t.parameters():
        p.requires_grad = False
    return t


def enable_qat(model: nn.Module, backend: str = "fbgemm") -> nn.Module:
    """ref/src/training/qat_trainer.py:300-308."""
    from torch.ao.quantization import get_default_qat_qconfig, prepare_qat
    import warnings
    model.train()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model.qconfig = get_default_qat_qconfig(backend)
        prepared = prepare_qat(model, inplace=False)
    prepared.train()
    return prepared


def make_optimizer(params, hparams: Dict, lr_scale: float = 1.0):
    """qat_trainer.py:271-276."""
    return torch.optim.AdamW(params, lr=float(hparams["lr"]) * lr_scale,
                             weight_decay=float(hparams["weight_decay"]))


def distill_loss(student_out, teacher_out, labels, hparams: Dict):
    """qat_trainer.py:265-268,343-349."""
    T = float(hparams["kd_temp"])
    alpha = float(hparams["kd_alpha"])
    loss_ce = F.cross_entropy(student_out, labels, label_smoothing=float(hparams["label_smoothing"]))
    loss_kd = F.kl_div(torch.log_softmax(student_out / T, dim=1), torch.softmax(teacher_out / T, dim=1),
                       reduction="batchmean") * (T ** 2)
    return alpha * loss_kd + (1.0 - alpha) * loss_ce, loss_kd, loss_ce


def distill_step(student: nn.Module, teacher: nn.Module, images, labels, optimizer: Optional[torch.optim.Optimizer],
                 hparams: Dict, clip: bool = True):
    """One iteration of the hot loop, qat_trainer.py:337-361.  Returns (loss, student_out, teacher_out)."""
    with torch.no_grad():
        teacher_out = teacher(images)
    student_out = student(images)
    loss, _, _ = distill_loss(student_out, teacher_out, labels, hparams)
    if optimizer is not None:
        optimizer.zero_grad(set_to_none=True)
    loss.backward()
    if clip:
        torch.nn.utils.clip_grad_norm_(student.parameters(), 1.0)
    if optimizer is not None:
        optimizer.step()
    return loss.detach(), student_out.detach(), teacher_out


def synthetic_batch(batch: int, seed: int = 0, img: int = 224):
    """images randn(B,3,224,224) seed-fixed, labels randint(0,10)."""
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(batch, 3, img, img, generator=g)
    labels = torch.randint(0, 10, (batch,), generator=g)
    return images, labels
