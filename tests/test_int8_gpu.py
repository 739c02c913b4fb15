"""GPU parity of the converted-int8 path (BASELINE.json configs[4]; SURVEY.md §8 a12 / §8c config-5 oracle).

Tier 1 (bit-exact): qv_int8_linear codes == torch.ops.quantized.linear codes (the reference's CPU engine) on identical quint8
inputs -- the committed golden vectors, live random cases at ViT shapes, and EVERY quantized module of a converted student fed
the CPU run's own inputs; qv_quantize_u8 == torch.quantize_per_tensor; dynamic qparams == the Python observer formula.
Tier 2: end-to-end logits of the float-glue executor vs the same glue on CPU with stock ops (oracle/int8_ref.py): the glue's
fp32 LayerNorm / softmax differ in the last bits between CPU and GPU, which can flip a dynamic-quantisation code, so the bound
is a few output quantisation steps (tolerance written below)."""
import copy
import os
import warnings

import numpy as np
import pytest
import torch

from parity_utils import build_models

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _run_ours(dev, qx, sx, zx, qw_int, sw, bias, sy, zy):
    from qatvit_b200 import ops
    qw = qw_int.to(torch.int8).contiguous().to(dev)
    wsum = qw_int.to(torch.int32).sum(1).to(torch.int32).contiguous().to(dev)
    M, N = qx.shape[0], qw.shape[0]
    qy = torch.empty(M, N, dtype=torch.uint8, device=dev)
    y = torch.empty(M, N, device=dev)
    sxd = torch.tensor([sx], dtype=torch.float32, device=dev)
    zxd = torch.tensor([zx], dtype=torch.int32, device=dev)
    args = (qx.contiguous().to(dev), sxd, zxd, qw, sw.to(torch.float32).contiguous().to(dev), wsum,
            None if bias is None else bias.to(dev), float(sy), int(zy))
    ops.int8_linear(*args, qy=qy)
    ops.int8_linear(*args, y=y)
    torch.cuda.synchronize()
    return qy.cpu(), y.cpu()


def test_int8_linear_golden_vectors(cuda_dev):
    z = np.load(os.path.join(GOLDEN, "qlinear_cases.npz"))
    for i in range(2):
        sx, zx, sy, zy = z[f"q{i}_p"]
        qy, y = _run_ours(cuda_dev, torch.from_numpy(z[f"q{i}_qx"]), float(sx), int(zx), torch.from_numpy(z[f"q{i}_qw"]),
                          torch.from_numpy(z[f"q{i}_sw"]), torch.from_numpy(z[f"q{i}_b"]), float(sy), int(zy))
        ref = torch.from_numpy(z[f"q{i}_qy"])
        assert torch.equal(qy, ref)
        assert torch.equal(y, (ref.float() - float(zy)) * np.float32(sy))


@pytest.mark.parametrize("M,N,K", [(197 * 4, 1152, 384), (197 * 4, 384, 1536), (197 * 2, 1536, 384), (196 * 2, 384, 768),
                                   (8, 10, 384), (130, 24, 96), (1, 16, 16), (300, 200, 144),
                                   (200, 2304, 64)])          # N > 2048: per-chunk column terms instead of the per-CTA table
@pytest.mark.parametrize("per_channel", [True, False])
def test_int8_linear_vs_quantized_linear(cuda_dev, M, N, K, per_channel):
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g) * 2
    w = torch.randn(N, K, generator=g) * 0.05
    b = torch.randn(N, generator=g) * 0.1
    sx, zx, sy, zy = 0.0313, 131, 0.0471, 117
    qx = torch.quantize_per_tensor(x, sx, zx, torch.quint8)
    if per_channel:
        sw = (w.abs().amax(1) / 127.0).clamp_min(1e-8)
        qw = torch.quantize_per_channel(w, sw, torch.zeros(N, dtype=torch.int64), 0, torch.qint8)
    else:
        sw = (w.abs().max() / 127.0).reshape(1)
        qw = torch.quantize_per_tensor(w, float(sw), 0, torch.qint8)
    ref = torch.ops.quantized.linear(qx, torch.ops.quantized.linear_prepack(qw, b), sy, zy)
    qy, y = _run_ours(cuda_dev, qx.int_repr(), sx, zx, qw.int_repr(), sw, b, sy, zy)
    assert torch.equal(qy, ref.int_repr())
    assert torch.equal(y, ref.dequantize())
    if N % 8 == 0:
        # third output form: the centred codes q_y - z_y as one bf16 plane (the attention operand of the compact executor)
        from qatvit_b200 import ops
        dev = cuda_dev
        qwd = qw.int_repr().to(torch.int8).contiguous().to(dev)
        codes = ops.int8_linear_codes(qx.int_repr().contiguous().to(dev), torch.tensor([sx], dtype=torch.float32, device=dev),
                                      torch.tensor([zx], dtype=torch.int32, device=dev), qwd, sw.to(torch.float32).contiguous().to(dev),
                                      qw.int_repr().to(torch.int32).sum(1).to(torch.int32).contiguous().to(dev), b.to(dev), sy, zy,
                                      torch.full((M, N), 7.0, dtype=torch.bfloat16, device=dev))
        torch.cuda.synchronize()
        assert torch.equal(codes.float().cpu(), ref.int_repr().float() - float(zy))


@pytest.mark.parametrize("shape", [(4, 197, 384), (3, 5, 7), (1, 3, 64, 64)])
def test_quantize_and_dynamic_qparams(cuda_dev, shape):
    from qatvit_b200 import ops
    from oracle import int8_ref
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(shape, generator=g) * 3 + 0.7
    acc = ops.new_minmax(cuda_dev)
    xd = x.to(cuda_dev)
    ops.minmax_accumulate(xd, acc)
    s = torch.empty(1, device=cuda_dev)
    zp = torch.empty(1, dtype=torch.int32, device=cuda_dev)
    ops.qparams_from_minmax(acc[0], 0, 255, s, zp)
    s_ref, z_ref = int8_ref.dynamic_qparams(x)
    assert float(s) == np.float32(s_ref) and int(zp) == z_ref
    q = ops.quantize_u8(xd, s, zp)
    ref = torch.quantize_per_tensor(x, s_ref, z_ref, torch.quint8).int_repr()
    assert torch.equal(q.cpu(), ref)
    # the one-launch form (qparams from the accumulator + codes): same qparams, same codes
    s2 = torch.empty(1, device=cuda_dev)
    zp2 = torch.empty(1, dtype=torch.int32, device=cuda_dev)
    q2 = ops.quantize_u8_dyn(xd, acc[0], s2, zp2, torch.empty(shape, dtype=torch.uint8, device=cuda_dev))
    assert torch.equal(s2, s) and torch.equal(zp2, zp) and torch.equal(q2, q)


@pytest.mark.parametrize("M,N,sy,zy", [(197 * 4, 1536, 0.0471, 117), (197 * 2, 1152, 0.31, 0), (64, 64, 1e-3, 255), (48, 16, 0.02, 64)])
def test_compact_glue_on_codes_is_bit_identical_to_the_fp32_glue(cuda_dev, M, N, sy, zy):
    """A converted Linear's output is (q - zy) * sy.  The compact glue works on q: centred codes for attention, table lookups
    for GELU + dynamic re-quantisation.  Against the elementwise fp32 chain on the dequantised tensor (the values qv_int8_linear
    writes as y, pinned to ref.dequantize() above): identical min / max, qparams and codes, bit for bit."""
    from qatvit_b200 import ops
    g = torch.Generator().manual_seed(M + N)
    q = torch.randint(0, 256, (M, N), generator=g, dtype=torch.int32).to(torch.uint8)
    q[0, :4] = torch.tensor([0, 255, zy, max(zy - 1, 0)], dtype=torch.uint8)         # the extreme codes are present
    qd = q.to(cuda_dev)
    y = ((qd.float() - float(zy)) * torch.tensor(sy, dtype=torch.float32, device=cuda_dev)).contiguous()   # = qv_int8_linear's y
    # attention operand: one exact bf16 plane of centred codes, codes * sy == y
    codes = ops.codes_from_u8(qd, zy, torch.empty(M, N, dtype=torch.bfloat16, device=cuda_dev))
    assert torch.equal(codes.float(), qd.float() - float(zy))
    assert torch.equal(codes.float() * torch.tensor(sy, dtype=torch.float32, device=cuda_dev), y)
    # GELU + min / max + dynamic re-quantisation: fp32 chain ...
    acc_a = ops.new_minmax(cuda_dev)
    gy = torch.empty_like(y)
    ops.gelu_minmax(y, gy, acc_a[0])
    s_a = torch.empty(1, device=cuda_dev)
    z_a = torch.empty(1, dtype=torch.int32, device=cuda_dev)
    ops.qparams_from_minmax(acc_a[0], 0, 255, s_a, z_a)
    q_a = ops.quantize_u8(gy, s_a, z_a)
    # ... vs the tables on the codes
    acc_b = ops.new_minmax(cuda_dev)
    ops.gelu_u8_minmax(qd, sy, zy, acc_b[0])
    s_b = torch.empty(1, device=cuda_dev)
    z_b = torch.empty(1, dtype=torch.int32, device=cuda_dev)
    q_b = ops.gelu_u8_requant(qd, sy, zy, acc_b[0], s_b, z_b, torch.empty(M, N, dtype=torch.uint8, device=cuda_dev))
    torch.cuda.synchronize()
    assert torch.equal(acc_a, acc_b)
    assert torch.equal(s_a, s_b) and torch.equal(z_a, z_b)
    assert torch.equal(q_a, q_b)
    # and against torch on the CPU: GELU(erf) of the dequantised tensor, Python-observer qparams, quantize_per_tensor
    # (CPU and GPU erf differ in the last bit on a few values, which can move a code by one)
    from oracle import int8_ref
    g_cpu = torch.nn.functional.gelu(y.cpu())
    s_ref, z_ref = int8_ref.dynamic_qparams(g_cpu)
    assert abs(float(s_b) - s_ref) <= 1e-6 * s_ref and int(z_b) == z_ref
    d = (q_b.cpu().int() - torch.quantize_per_tensor(g_cpu, s_ref, z_ref, torch.quint8).int_repr().int()).abs()
    assert int(d.max()) <= 1 and float((d > 0).float().mean()) < 1e-2


def test_compact_glue_rejects_bad_arguments(cuda_dev):
    from qatvit_b200 import ops
    q = torch.zeros(3, 5, dtype=torch.uint8, device=cuda_dev)                  # 15 codes: not a multiple of 16
    acc = ops.new_minmax(cuda_dev)
    with pytest.raises(RuntimeError, match="multiple of 16"):
        ops.codes_from_u8(q, 0, torch.empty(3, 5, dtype=torch.bfloat16, device=cuda_dev))
    with pytest.raises(RuntimeError, match="multiple of 16"):
        ops.gelu_u8_minmax(q, 0.1, 0, acc[0])
    q16 = torch.zeros(4, 4, dtype=torch.uint8, device=cuda_dev)
    with pytest.raises(RuntimeError, match="zero point"):
        ops.codes_from_u8(q16, 256, torch.empty(4, 4, dtype=torch.bfloat16, device=cuda_dev))
    with pytest.raises(RuntimeError):
        ops.gelu_u8_minmax(q16, 0.0, 0, acc[0])                                # scale must be positive
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.codes_from_u8(q16.cpu(), 0, torch.empty(4, 4, dtype=torch.bfloat16, device=cuda_dev))


def _converted(backend, sname, tname, img, B):
    from torch.ao.quantization import convert
    vr, prepared, teacher = build_models(backend, sname, tname, img)
    images, labels = vr.synthetic_batch(B, seed=3, img=img)
    hp = dict(vr.DEFAULT_HPARAMS)
    for _ in range(2):                                   # give the observers a couple of EMA steps
        vr.distill_step(prepared, teacher, images, labels, None, hp, clip=False)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        conv = convert(copy.deepcopy(prepared).eval(), inplace=False)      # ref qat_trainer.py:377-379
    return conv, images


@pytest.mark.parametrize("backend,sname,img,B,compact", [("fbgemm", "vit_test_tiny", 64, 4, True), ("qnnpack", "vit_test_tiny", 64, 3, True),
                                                         ("fbgemm", "vit_small_patch16_224", 224, 2, True),
                                                         ("fbgemm", "vit_test_tiny", 64, 4, False), ("qnnpack", "vit_test_tiny", 64, 3, "attn")])
def test_converted_student_per_layer_and_end_to_end(cuda_dev, backend, sname, img, B, compact):
    from qatvit_b200 import ops
    from qatvit_b200.int8 import ConvertedStudent
    from oracle import int8_ref
    conv, images = _converted(backend, sname, "vit_test_teacher", img, B)
    trace = {}
    ref_logits = int8_ref.converted_forward(conv, images, trace)
    ex = ConvertedStudent(conv, B, cuda_dev, compact=compact)
    assert (ex.c_attn, ex.c_gelu, ex.c_ln) == (compact in (True, "attn"), compact in (True, "gelu"), compact is True)
    # tier 1: every quantized module, on the CPU run's own quint8 input, gives bit-identical codes
    qlins = {"head": ex.head}
    for i, blk in enumerate(ex.blocks):
        qlins.update({f"blocks.{i}.attn.qkv": blk["qkv"], f"blocks.{i}.attn.proj": blk["proj"],
                      f"blocks.{i}.mlp.fc1": blk["fc1"], f"blocks.{i}.mlp.fc2": blk["fc2"]})
    assert set(qlins) | {"patch_embed.proj"} == set(trace)
    for name, ql in qlins.items():
        qx, qy = trace[name]
        x2 = qx.int_repr().reshape(-1, ql.K).contiguous().to(cuda_dev)
        sx = torch.tensor([qx.q_scale()], dtype=torch.float32, device=cuda_dev)
        zx = torch.tensor([qx.q_zero_point()], dtype=torch.int32, device=cuda_dev)
        got = ops.int8_linear(x2, sx, zx, ql.qw, ql.sw, ql.wsum, ql.bias, ql.sy, ql.zy)
        assert torch.equal(got.cpu(), qy.int_repr().reshape(-1, ql.N)), name
    # the conv: quantise + im2col on device from the float image, then the same int8 GEMM
    qx, qy = trace["patch_embed.proj"]
    ops.im2col_u8(images.to(cuda_dev), ex.in_scale, ex.in_zp, B, 3, ex.HW, ex.ps, ex.q_img)
    got = ops.int8_linear(ex.q_img, ex.in_scale, ex.in_zp, ex.conv.qw, ex.conv.sw, ex.conv.wsum, ex.conv.bias, ex.conv.sy,
                          ex.conv.zy)
    assert torch.equal(got.cpu().view(B, ex.P, ex.D), qy.int_repr().flatten(2).transpose(1, 2))
    # tier 2: end to end through the float glue
    ours = {}
    logits = ex(images.to(cuda_dev), ours)
    torch.cuda.synchronize()
    # (a) up to the first dynamic quantisation everything is decided on identical inputs: same qparams, and the codes differ
    #     only where the fp32 LayerNorm of the two sides rounds differently across a .5 boundary (<= 1 code on < 0.1 %)
    qh, s, z = ours["blocks.0.attn.qkv"]
    ref_qx = trace["blocks.0.attn.qkv"][0]
    assert abs(float(s) - ref_qx.q_scale()) <= 1e-6 * ref_qx.q_scale() and int(z) == ref_qx.q_zero_point()
    d = (qh.cpu().int() - ref_qx.int_repr().reshape(qh.shape).int()).abs()
    assert int(d.max()) <= 1 and float((d > 0).float().mean()) < 1e-3
    # (b) logits: a flipped code perturbs everything downstream (12 blocks x 4 dynamic re-quantisations), so the bound is a
    #     few output quantisation steps of the head (128-level grid): 6 steps max, relative L2 < 8e-2
    step = conv.model.head.scale
    diff = logits.cpu() - ref_logits
    assert float(diff.abs().max()) <= 6.0 * step + 1e-6
    assert float(diff.norm() / ref_logits.norm()) < 8e-2


def test_compact_glue_end_to_end_against_the_fp32_glue(cuda_dev):
    """The same converted student through the realisations of the float glue.  The table half (GELU + re-quantisation on fc1's
    codes, qparams folded into the quantising pass) and the LayerNorm half (codes from the saved row statistics instead of an fp32
    copy of the LayerNorm output) must reproduce the fp32 glue BIT FOR BIT, logits and every dynamic (scale, zero point)
    included; the attention half computes softmax(QK^T)V from exact integer codes instead of fp32 hi/lo planes of the
    dequantised values (products exact instead of ~2^-16 relative), so block 0's fc2 input may differ by a code on a few elements
    and the logits by a few head quantisation steps -- the bound the CPU mirror is held to."""
    from qatvit_b200.int8 import ConvertedStudent
    conv, images = _converted("fbgemm", "vit_test_tiny", "vit_test_teacher", 64, 4)
    x = images.to(cuda_dev)
    runs = {}
    for mode in (False, "gelu", "ln", ("gelu", "ln"), True):
        ex = ConvertedStudent(conv, 4, cuda_dev, compact=mode)
        tr = {}
        runs[mode] = (ex(x, tr).clone(), tr, ex.dyn_scale.clone(), ex.dyn_zp.clone())
    torch.cuda.synchronize()
    base, full = runs[False], runs[True]
    for mode in ("gelu", "ln", ("gelu", "ln")):
        other = runs[mode]
        assert torch.equal(base[0], other[0]), mode
        assert torch.equal(base[2], other[2]) and torch.equal(base[3], other[3]), mode
        for k in base[1]:
            for a, b in zip(base[1][k], other[1][k]):
                assert torch.equal(a, b), (mode, k)
    qa, sa, za = base[1]["blocks.0.mlp.fc2"]
    qb, sb, zb = full[1]["blocks.0.mlp.fc2"]
    assert abs(float(sa) - float(sb)) <= 1e-5 * float(sa) and abs(int(za) - int(zb)) <= 1
    d = (qa.int() - qb.int()).abs()
    assert int(d.max()) <= 1 and float((d > 0).float().mean()) < 1e-2
    step = conv.model.head.scale
    diff = (full[0] - base[0]).cpu()
    assert torch.isfinite(full[0]).all() and float(diff.abs().max()) <= 6.0 * step + 1e-6
    with pytest.raises(ValueError):
        ConvertedStudent(conv, 4, cuda_dev, compact="planes")


@pytest.mark.parametrize("compact", [True, False])
def test_converted_executor_takes_a_ragged_tail_batch(cuda_dev, compact):
    """The reference's eval DataLoader has no drop_last (ref qat_trainer.py:237-254: 10 000 % 256 = 16): an executor built for B
    images fed b < B must give, bit for bit, what an executor built for b gives -- then run the full batch again unchanged."""
    from qatvit_b200.int8 import ConvertedStudent
    conv, images = _converted("fbgemm", "vit_test_tiny", "vit_test_teacher", 64, 5)
    x = images.to(cuda_dev)
    big = ConvertedStudent(conv, 5, cuda_dev, compact=compact)
    full_a = big(x).clone()
    for b in (3, 1):
        tail = big(x[:b].contiguous()).clone()
        ref = ConvertedStudent(conv, b, cuda_dev, compact=compact)(x[:b].contiguous()).clone()
        torch.cuda.synchronize()
        assert tail.shape == (b, 10) and torch.equal(tail, ref), b
    full_b = big(x).clone()
    torch.cuda.synchronize()
    assert torch.equal(full_a, full_b)
    for bad in (torch.zeros(6, 3, 64, 64, device=cuda_dev), torch.zeros(0, 3, 64, 64, device=cuda_dev),
                torch.zeros(2, 3, 32, 32, device=cuda_dev), x[:2].cpu()):
        with pytest.raises(RuntimeError, match="1..5"):
            big(bad)


def test_best_converted_pth_reader_runs_identically(cuda_dev, tmp_path):
    """ConvertedStudent.from_state_dict(best_converted.pth) == ConvertedStudent(converted module): same operands, same
    kernels, bit-identical logits (the stock file format of ref qat_trainer.py:386-388 is all the executor needs)."""
    from qatvit_b200.int8 import ConvertedStudent
    conv, images = _converted("fbgemm", "vit_test_tiny", "vit_test_teacher", 64, 4)
    path = tmp_path / "best_converted.pth"
    torch.save(conv.state_dict(), path)
    a = ConvertedStudent(conv, 4, cuda_dev)
    b = ConvertedStudent.from_state_dict(str(path), 4, cuda_dev)
    la = a(images.to(cuda_dev)).clone()
    lb = b(images.to(cuda_dev)).clone()
    torch.cuda.synchronize()
    assert torch.isfinite(la).all() and torch.equal(la, lb)
