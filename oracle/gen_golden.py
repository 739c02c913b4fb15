"""oracle/gen_golden.py -- regenerate tests/golden/*.npz / *.json from the REFERENCE's own code paths.

Run in the build container (needs torch and, for the model census, /root/reference):
    python oracle/gen_golden.py
Sources of truth:
  * torch.fused_moving_avg_obs_fake_quant on CPU -- the op behind every FusedMovingAvgObsFakeQuantize the reference's
    prepare_qat call creates (ref/src/training/qat_trainer.py:306-307);
  * torch.ops.quantized.linear (x86 engine) -- what the converted model executes (ref qat_trainer.py:379-380);
  * the reference's unmodified QATWrapper / model_registry imported through the timm shim -- module census and
    state_dict layout of best_qat.pth (ref qat_trainer.py:385).
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")


def fq_cases():
    g = torch.Generator().manual_seed(20261018)
    cases = {}
    idx = 0
    for (qmin, qmax, sym, C) in [(0, 127, False, 0), (0, 255, False, 0), (-128, 127, True, 0), (-128, 127, True, 6)]:
        for kind in ["normal", "tiny", "positive", "negative", "zero_then_ties", "ties"]:
            shape = (6, 20)
            x = torch.randn(shape, generator=g)
            if kind == "normal":
                xs = [x * 2 + 0.3, x * 3 - 1, x * 0.5]
            elif kind == "tiny":
                xs = [x * 1e-4, x * 2e-4, x * 1e-5]
            elif kind == "positive":
                xs = [x.abs() + 0.1, x.abs() * 2, x.abs()]
            elif kind == "negative":
                xs = [-x.abs() - 0.1, -x.abs() * 2, -x.abs()]
            elif kind == "zero_then_ties":
                xs = [torch.zeros(shape), (torch.randint(-300, 300, shape, generator=g).float() + 0.5) * 0.05, x]
            else:
                t = (torch.randint(-100, 100, shape, generator=g).float() + 0.5) * 0.125
                xs = [t, t * 2, t]
            if C:
                xs[0][0] = xs[0][0].abs() + 1e-3
                xs[0][1] = -xs[0][1].abs() - 1e-3
                xs[0][2] = 0.0
            mn = torch.empty(0) if C else torch.tensor(float("inf"))
            mx = torch.empty(0) if C else torch.tensor(float("-inf"))
            s, z = torch.ones(1), torch.zeros(1, dtype=torch.int32)
            for step, xi in enumerate(xs):
                xt = xi.clone().requires_grad_(True)
                y = torch.fused_moving_avg_obs_fake_quant(xt, torch.tensor([1]), torch.tensor([1]), mn, mx, s, z, 0.01, qmin,
                                                          qmax, 0, bool(C), sym)
                y.backward(torch.ones_like(y))
                key = f"c{idx}_s{step}"
                cases[key + "_x"] = xi.numpy().copy()
                cases[key + "_y"] = y.detach().numpy().copy()
                cases[key + "_mask"] = xt.grad.numpy().astype(np.uint8)
                cases[key + "_min"] = mn.numpy().reshape(-1).copy()
                cases[key + "_max"] = mx.numpy().reshape(-1).copy()
                cases[key + "_scale"] = s.numpy().copy()
                cases[key + "_zp"] = z.numpy().copy()
            cases[f"c{idx}_cfg"] = np.array([qmin, qmax, int(sym), C], dtype=np.int64)
            idx += 1
    cases["n_cases"] = np.array([idx])
    np.savez_compressed(os.path.join(OUT, "fq_cases.npz"), **cases)
    print("fq_cases:", idx, "cases")


def cqp_ties():
    """(min, max) pairs engineered onto exact .5 zero-point ties: they pin fbgemm's float-scale / '+'-error variant."""
    rng = np.random.default_rng(7)
    rows = []
    while len(rows) < 1500:
        qmin, qmax = [(0, 255), (0, 127), (-128, 127)][len(rows) % 3]
        k = rng.integers(1, qmax - qmin)
        s0 = np.float32(rng.uniform(0.001, 0.1))
        mn = np.float32(-(k + 0.5) * s0)
        mx = np.float32((qmax - qmin - (k + 0.5)) * s0)
        x = torch.tensor([float(mn), float(mx)], dtype=torch.float32)
        tmn, tmx = torch.tensor(float("inf")), torch.tensor(float("-inf"))
        s, z = torch.ones(1), torch.zeros(1, dtype=torch.int32)
        torch.fused_moving_avg_obs_fake_quant(x, torch.tensor([1]), torch.tensor([1]), tmn, tmx, s, z, 0.01, qmin, qmax, 0,
                                              False, False)
        rows.append((float(mn), float(mx), qmin, qmax, float(s), int(z)))
    a = np.array(rows, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "cqp_ties.npz"), rows=a)
    print("cqp_ties:", len(rows))


def qlinear_cases():
    torch.manual_seed(3)
    out = {}
    for i, per_channel in enumerate([True, False]):
        M, K, N = 37, 96, 24
        x = torch.randn(M, K) * 2
        w = torch.randn(N, K) * 0.05
        b = torch.randn(N) * 0.1
        sx, zx, sy, zy = 0.031, 63, 0.047, 58
        qx = torch.quantize_per_tensor(x, sx, zx, torch.quint8)
        if per_channel:
            sw = (w.abs().amax(1) / 127.0).clamp_min(1e-8)
            qw = torch.quantize_per_channel(w, sw, torch.zeros(N, dtype=torch.int64), 0, torch.qint8)
        else:
            sw = (w.abs().max() / 127.0).reshape(1)
            qw = torch.quantize_per_tensor(w, float(sw), 0, torch.qint8)
        packed = torch.ops.quantized.linear_prepack(qw, b)
        qy = torch.ops.quantized.linear(qx, packed, sy, zy)
        out[f"q{i}_qx"] = qx.int_repr().numpy()
        out[f"q{i}_qw"] = qw.int_repr().numpy()
        out[f"q{i}_sw"] = sw.numpy().astype(np.float32)
        out[f"q{i}_b"] = b.numpy()
        out[f"q{i}_qy"] = qy.int_repr().numpy()
        out[f"q{i}_p"] = np.array([sx, zx, sy, zy], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "qlinear_cases.npz"), **out)
    print("qlinear_cases ok")


def model_census():
    from oracle import vit_ref as vr
    reg = vr.load_reference_registry()
    if reg is None:
        print("model census skipped: /root/reference not present")
        return
    torch.manual_seed(0)
    student = reg.create_model("vit_small_patch16_224_student", pretrained=False, num_classes=10, qat_wrapper=True)
    census = {}
    for backend in ("fbgemm", "qnnpack"):
        prepared = vr.enable_qat(student, backend)
        x, _ = vr.synthetic_batch(1)
        prepared(x)     # first call sizes the per-channel observer buffers
        census[backend] = {k: [str(v.dtype), list(v.shape)] for k, v in prepared.state_dict().items()}
    census["reference_class"] = f"{type(student).__module__}.{type(student).__name__}"
    with open(os.path.join(OUT, "student_state_dict_census.json"), "w") as f:
        json.dump(census, f)
    print("census:", {k: len(v) for k, v in census.items() if isinstance(v, dict)})
    # a tiny end-to-end golden step through the REFERENCE's own QATWrapper class (structure restated, arithmetic = torch)
    torch.manual_seed(0)
    tiny = reg.QATWrapper(vr.create_model("vit_test_tiny", num_classes=10, img_size=64))
    teacher = vr.create_model("vit_test_teacher", num_classes=10, img_size=64).eval()
    wsum = float(sum(v.double().abs().sum() for v in tiny.state_dict().values()))
    prepared = vr.enable_qat(tiny, "fbgemm")
    images, labels = vr.synthetic_batch(2, seed=9, img=64)
    loss, s_out, t_out = vr.distill_step(prepared, teacher, images, labels, None, dict(vr.DEFAULT_HPARAMS), clip=False)
    # weights / images are re-created from the seeds by the test (wsum guards against RNG drift between torch builds)
    np.savez_compressed(os.path.join(OUT, "tiny_step.npz"), loss=np.array([float(loss)]), student_out=s_out.numpy(),
                        teacher_out=t_out.numpy(), head_w_grad=prepared.model.head.weight.grad.numpy(),
                        weight_abs_sum=np.array([wsum]))
    print("tiny_step loss", float(loss))


def resize_cases():
    """The reference's input transform (ref qat_trainer.py:210-216) run LIVE through Pillow + torchvision on small uint8 images:
    a random CIFAR-sized image and extreme patterns at 224, random and non-square inputs at smaller target sizes (same code path,
    smaller fixture).  Stored: inputs, the uint8 images Pillow's bicubic resize returns, and ToTensor + Normalize as the 3 x 256
    table of the float32 value each uint8 level maps to per channel (the map is pointwise)."""
    from PIL import Image
    from torchvision import transforms
    import torch
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    rng = np.random.default_rng(2024)
    cases = {"rand224": (rng.integers(0, 256, (32, 32, 3), dtype=np.uint8), 224),
             "checker224": (np.zeros((32, 32, 3), np.uint8), 224), "white224": (np.full((32, 32, 3), 255, np.uint8), 224),
             "binary224": ((rng.integers(0, 2, (32, 32, 3)) * 255).astype(np.uint8), 224),
             "rand96": (rng.integers(0, 256, (32, 32, 3), dtype=np.uint8), 96),
             "wide64": (rng.integers(0, 256, (32, 48, 3), dtype=np.uint8), 64),
             "tall80": (rng.integers(0, 256, (40, 32, 3), dtype=np.uint8), 80),
             "down24": (rng.integers(0, 256, (50, 70, 3), dtype=np.uint8), 24)}
    cases["checker224"][0][::2, ::2] = 255
    out = {}
    for k, (v, size) in cases.items():
        rz = transforms.Resize(size, interpolation=transforms.InterpolationMode.BICUBIC)
        pil = rz(Image.fromarray(v))
        full = transforms.Compose([rz, transforms.ToTensor(), transforms.Normalize(mean=mean, std=std)])(Image.fromarray(v)).numpy()
        u8 = np.array(pil)
        out["in_" + k] = v
        out["u8_" + k] = u8
        out["size_" + k] = np.array([size])
        # the float output is a pointwise function of the resized uint8 image: check, then store only the table
        lut_probe = transforms.Compose([transforms.ToTensor(), transforms.Normalize(mean=mean, std=std)])
        assert np.array_equal(full, lut_probe(Image.fromarray(u8)).numpy())
    levels = np.repeat(np.arange(256, dtype=np.uint8)[None, :, None], 3, axis=2)            # [1, 256, 3] image holding every level
    out["level_table"] = transforms.Compose([transforms.ToTensor(), transforms.Normalize(mean=mean, std=std)])(
        Image.fromarray(levels)).numpy()[:, 0, :]                                           # [3, 256] float32
    np.savez_compressed(os.path.join(OUT, "resize.npz"), **out)
    print("resize cases", {k: v.shape for k, v in out.items() if k.startswith("u8_")})


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    resize_cases()
    fq_cases()
    cqp_ties()
    qlinear_cases()
    model_census()
