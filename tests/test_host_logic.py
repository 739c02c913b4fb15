"""CPU: host-side logic of the engines that needs no device -- the guard that refuses timm VisionTransformer variants the
fused path does not compute (create_student / create_teacher forward **kwargs to timm: ref model_registry.py:167-172,228-233),
and the buffer views that let an engine built for batch B run the ragged last batch of an epoch (the reference's DataLoaders
have no drop_last: ref qat_trainer.py:227-254)."""
import math

import pytest
import torch
from torch import nn


def _tiny():
    import qatvit_b200  # noqa: F401
    from qatvit_b200 import vit
    return vit.create_model("vit_test_tiny", num_classes=10, img_size=64)


def test_default_vit_is_accepted_and_dims_follow_the_module_tree():
    from qatvit_b200.engine import _ViTDims
    v = _tiny()
    d = _ViTDims(v, 5)
    assert (d.B, d.T, d.M) == (5, v.patch_embed.num_patches + 1, 5 * (v.patch_embed.num_patches + 1))
    assert d.D == v.embed_dim and d.hd == 64 and d.H * 64 == d.D and d.L == len(v.blocks)
    assert d.F == v.blocks[0].mlp.fc1.out_features and d.C == 10 and d.Kc == 3 * d.ps * d.ps
    d3 = d.with_batch(3)
    assert (d3.B, d3.M, d3.T, d3.D) == (3, 3 * d.T, d.T, d.D) and d.B == 5        # a copy: the construction dims are untouched


@pytest.mark.parametrize("mutate,what", [
    (lambda v: setattr(v, "global_pool", "avg"), "global_pool"),
    (lambda v: setattr(v, "fc_norm", nn.LayerNorm(v.embed_dim)), "fc_norm"),
    (lambda v: setattr(v, "pos_drop", nn.Dropout(0.1)), "pos_drop"),
    (lambda v: setattr(v, "no_embed_class", True), "no_embed_class"),
    (lambda v: setattr(v, "num_prefix_tokens", 5), "prefix"),
    (lambda v: setattr(v.blocks[1], "ls1", nn.Linear(4, 4)), "blocks.1.ls1"),
    (lambda v: setattr(v.blocks[0], "drop_path2", nn.Dropout(0.2)), "blocks.0.drop_path2"),
    (lambda v: setattr(v.blocks[0].attn, "q_norm", nn.LayerNorm(64)), "blocks.0.attn.q_norm"),
    (lambda v: setattr(v.blocks[0].attn, "attn_drop", nn.Dropout(0.5)), "blocks.0.attn.attn_drop"),
    (lambda v: setattr(v.blocks[0].mlp, "act", nn.GELU(approximate="tanh")), "mlp.act"),
    (lambda v: setattr(v.blocks[0].mlp, "act", nn.ReLU()), "mlp.act"),
    (lambda v: setattr(v.blocks[0].mlp.fc1, "bias", None), "Linear without bias"),
])
def test_unsupported_vit_variants_are_refused(mutate, what):
    """A variant would run through the engine as if it were the default ViT (a different function than the module tree, no
    error): every one of them must be named in a NotImplementedError when the engine's dims are taken."""
    from qatvit_b200.engine import _ViTDims
    v = _tiny()
    mutate(v)
    with pytest.raises(NotImplementedError, match="not on the fused path") as e:
        _ViTDims(v, 2)
    assert what in str(e.value)


def test_identity_like_modules_are_not_variants():
    from qatvit_b200.engine import _ViTDims
    v = _tiny()
    v.pos_drop = nn.Dropout(0.0)                 # timm's default: p = 0
    v.blocks[0].ls1 = nn.Identity()
    v.blocks[0].attn.proj_drop = nn.Dropout(0.0)
    _ViTDims(v, 2)


def test_head_dim_and_token_limits():
    from qatvit_b200 import vit
    from qatvit_b200.engine import _ViTDims
    v = _tiny()
    v.blocks[0].attn.num_heads = v.embed_dim // 32               # head_dim 32
    with pytest.raises(NotImplementedError, match="head_dim 64"):
        _ViTDims(v, 2)
    big = vit.create_model("vit_test_tiny", num_classes=10, img_size=240)    # 15 x 15 patches + cls = 226 tokens > 224
    with pytest.raises(NotImplementedError, match="224 tokens"):
        _ViTDims(big, 2)


def test_batch_buffer_views_share_storage_and_match_an_engine_built_for_the_tail():
    from qatvit_b200.engine import _BatchBuffers, _ViTDims

    class Bufs(_BatchBuffers):
        def __init__(self, vit, batch):
            self.binds = []
            self.d = _ViTDims(vit, batch)
            self._bb_init(torch.device("cpu"), self.d)
            self._bb_add("x", lambda d: (d.M, d.D), count=2)
            self._bb_add("planes", lambda d: (2, d.M, 3 * d.D), torch.bfloat16)
            self._bb_add("lse", lambda d: (d.B * d.H * d.T,))
            self._bb_add("pad", lambda d: (d.B, d.ldS), zero=True)

        def _bb_on_bind(self, b):
            self.binds.append(b)

    v = _tiny()
    eng = Bufs(v, 8)
    full_ptr = eng.x[0].data_ptr()
    assert eng.x[0].shape == (8 * eng.d.T, eng.d.D) and len(eng.x) == 2
    eng.x[0].fill_(7.0)
    eng.pad.fill_(3.0)
    eng._bb_bind(5)                                              # ragged tail: 5 of 8 images
    T, D, H = eng.d.T, eng.d.D, eng.d.H
    assert (eng.d.B, eng.d.M) == (5, 5 * T)
    assert eng.x[0].shape == (5 * T, D) and eng.x[0].is_contiguous() and eng.x[0].data_ptr() == full_ptr
    assert eng.planes.shape == (2, 5 * T, 3 * D) and eng.planes.is_contiguous()       # the layout an engine built for 5 would own
    assert eng.planes.stride(0) == 5 * T * 3 * D
    assert eng.lse.shape == (5 * H * T,)
    assert float(eng.x[0][0, 0]) == 7.0                                               # same storage, nothing copied
    assert float(eng.pad.abs().sum()) == 0.0                                          # zero-padded buffers are cleared on re-layout
    tail_views = (eng.x[0], eng.planes)
    eng._bb_bind(5)                                              # same size again: no re-bind
    eng._bb_bind(8)
    assert eng.x[0].shape == (8 * T, D) and (eng.d.B, eng.d.M) == (8, 8 * T)
    eng._bb_bind(5)                                              # views are cached per batch size: the very same tensor objects
    assert eng.x[0] is tail_views[0] and eng.planes is tail_views[1]
    assert eng.binds == [5, 8, 5]
    small = Bufs(v, 5)                                           # an engine built for the tail owns tensors of the same geometry
    for name in ("planes", "lse", "pad"):
        a, b = getattr(eng, name), getattr(small, name)
        assert a.shape == b.shape and a.stride() == b.stride() and a.dtype == b.dtype
    assert math.prod(eng.x[1].shape) == math.prod(small.x[1].shape)


def test_batch_check_accepts_any_batch_up_to_the_construction_batch():
    from qatvit_b200.engine import _BatchBuffers, _ViTDims
    v = _tiny()
    eng = _BatchBuffers()
    eng._bb_init(torch.device("cpu"), _ViTDims(v, 8))
    assert eng._bb_check(torch.zeros(8, 3, 64, 64), "student") == 8
    assert eng._bb_check(torch.zeros(1, 3, 64, 64), "student") == 1
    for bad in (torch.zeros(9, 3, 64, 64), torch.zeros(0, 3, 64, 64), torch.zeros(4, 3, 32, 32), torch.zeros(4, 1, 64, 64),
                torch.zeros(3, 64, 64)):
        with pytest.raises(RuntimeError, match="any batch 1..8 is accepted"):
            eng._bb_check(bad, "student")
