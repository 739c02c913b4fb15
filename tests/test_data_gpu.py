"""GPU parity of the input transform (qv_resize_normalize_u8 through qatvit_b200.data.GpuImageTransform) against the oracle
(oracle/resize_ref.py, pinned to Pillow + torchvision) and the committed golden outputs of the live reference pipeline
(ref/src/training/qat_trainer.py:210-216): BIT-EXACT float32 (integer resample, IEEE divide / subtract)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_resize_normalize_golden_bit_exact(cuda_dev):
    from qatvit_b200.data import GpuImageTransform
    g = np.load(os.path.join(GOLD, "resize.npz"))
    table = g["level_table"]
    n_cases = 0
    for k in g.files:
        if not k.startswith("in_"):
            continue
        n = k[3:]
        u8 = g["u8_" + n]
        want = np.stack([table[c][u8[:, :, c]] for c in range(3)]).astype(np.float32)
        y = GpuImageTransform(int(g["size_" + n][0]))(torch.from_numpy(g[k])[None].contiguous().to(cuda_dev))
        assert np.array_equal(y[0].cpu().numpy(), want), n
        n_cases += 1
    assert n_cases >= 8


@pytest.mark.parametrize("B,H,W", [(256, 32, 32), (7, 32, 32), (3, 40, 32), (2, 32, 48), (1, 64, 64)])
def test_resize_normalize_vs_oracle(cuda_dev, B, H, W):
    from oracle import resize_ref as rr
    from qatvit_b200.data import GpuImageTransform
    rng = np.random.default_rng(B * 1000 + H + W)
    imgs = rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8)
    tf = GpuImageTransform(224)
    y = tf(torch.from_numpy(imgs).to(cuda_dev))
    torch.cuda.synchronize()
    pick = range(B) if B <= 8 else [0, 1, B // 2, B - 1]
    for i in pick:
        assert np.array_equal(y[i].cpu().numpy(), rr.transform(imgs[i])), i
    if B > 8:      # every image went through the same code: spot-check the rest through a checksum of checksums
        ref = np.stack([rr.transform(imgs[i]).astype(np.float64).sum() for i in range(0, B, 16)])
        assert np.array_equal(y[::16].double().sum(dim=(1, 2, 3)).cpu().numpy(), ref)


def test_resize_normalize_edges(cuda_dev):
    from qatvit_b200.data import GpuImageTransform
    tf = GpuImageTransform(224)
    empty = tf(torch.empty(0, 32, 32, 3, dtype=torch.uint8, device=cuda_dev))
    assert tuple(empty.shape) == (0, 3, 224, 224)
    with pytest.raises(RuntimeError, match="uint8 CUDA"):
        tf(torch.zeros(2, 32, 32, 3, dtype=torch.uint8))                       # CPU tensor: no fallback
    with pytest.raises(RuntimeError, match="uint8 CUDA"):
        tf(torch.zeros(2, 32, 32, 3, device=cuda_dev))                          # wrong dtype
    with pytest.raises(RuntimeError, match="channels"):
        tf(torch.zeros(2, 32, 32, 1, dtype=torch.uint8, device=cuda_dev))
    out = torch.empty(2, 3, 224, 224, device=cuda_dev)
    x = torch.full((2, 32, 32, 3), 255, dtype=torch.uint8, device=cuda_dev)
    assert tf(x, out=out) is out
    want = (torch.tensor(1.0) - torch.tensor([0.485, 0.456, 0.406])) / torch.tensor([0.229, 0.224, 0.225])
    assert torch.equal(out[:, :, 0, 0].cpu(), want.expand(2, 3))               # saturated white stays exactly white


def test_transform_feeds_the_training_step(cuda_dev):
    """uint8 batch -> GpuImageTransform -> QATDistillStep: the images the engine sees equal the CPU pipeline's, so the step equals
    the step on CPU-transformed images bit for bit."""
    import copy
    from parity_utils import build_models
    from oracle import resize_ref as rr
    from qatvit_b200.data import GpuImageTransform
    from qatvit_b200.engine import QATDistillStep
    vr, prepared, teacher = build_models("fbgemm", "vit_test_tiny", "vit_test_teacher", 64)
    B = 4
    rng = np.random.default_rng(3)
    raw = rng.integers(0, 256, (B, 16, 16, 3), dtype=np.uint8)
    labels = torch.tensor([1, 3, 5, 7], device=cuda_dev)
    losses = []
    for mode in ("gpu", "cpu"):
        step = QATDistillStep(copy.deepcopy(prepared).to(cuda_dev), copy.deepcopy(teacher).to(cuda_dev), B, dict(vr.DEFAULT_HPARAMS))
        if mode == "gpu":
            images = GpuImageTransform(64)(torch.from_numpy(raw).to(cuda_dev))
        else:
            images = torch.from_numpy(np.stack([rr.transform(r, size=64) for r in raw])).to(cuda_dev)
        losses.append(step(images, labels).clone())
    assert torch.equal(losses[0], losses[1])
