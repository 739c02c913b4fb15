"""Error behaviour of the C ABI on a GPU box: bad arguments come back as negative codes with a message (raised as RuntimeError
by the Python layer), nothing is launched, and the library keeps working afterwards (errors by code +
qv_last_error(); no allocation, no fallback)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_bad_arguments_raise_and_library_survives(cuda_dev):
    from qatvit_b200 import ops
    from qatvit_b200.ops import Op, PAIRS_FP32
    dev = cuda_dev
    n0 = ops.launch_count()
    a = ops.split_planes(torch.randn(64, 40, device=dev))
    b = ops.split_planes(torch.randn(32, 40, device=dev))
    with pytest.raises(RuntimeError, match="must be 1 or 2"):
        ops.gemm(Op.full(a), Op.full(b), 64, 32, 40, (3, 1))
    with pytest.raises(RuntimeError, match="empty gemm"):
        ops.gemm(Op.full(a), Op.full(b), 64, 32, 0, PAIRS_FP32, out=torch.empty(64, 32, device=dev))
    with pytest.raises(RuntimeError, match="row pitch"):               # K = 44: row pitch not a multiple of 8 bf16
        bad = ops.split_planes(torch.randn(64, 44, device=dev))
        ops.gemm(Op.full(bad), Op.full(ops.split_planes(torch.randn(32, 44, device=dev))), 64, 32, 44, PAIRS_FP32)
    with pytest.raises(RuntimeError, match="T <= 224"):
        ops.attn_fwd(torch.zeros(2, 225, 192, dtype=torch.bfloat16, device=dev), 1, 225, 1, 0.125,
                     torch.empty(2, 225, 64, dtype=torch.bfloat16, device=dev))
    with pytest.raises(RuntimeError, match="multiple of 16"):
        z = torch.zeros(1, dtype=torch.int32, device=dev)
        ops.int8_linear(torch.zeros(8, 24, dtype=torch.uint8, device=dev), torch.ones(1, device=dev), z,
                        torch.zeros(16, 24, dtype=torch.int8, device=dev), torch.ones(1, device=dev),
                        torch.zeros(16, dtype=torch.int32, device=dev), None, 1.0, 0)
    with pytest.raises(RuntimeError, match="GELU"):                    # GELU epilogue exists only for plane output
        from qatvit_b200 import _lib
        import ctypes
        args = _lib.GemmArgs()
        Op.full(a).fill(args.a)
        Op.full(b).fill(args.b)
        args.a_planes = args.b_planes = 2
        args.M, args.N, args.K = 64, 32, 40
        out = torch.empty(64, 32, device=dev)
        ops.Out.full(out).fill(args.out)
        args.act = 1
        _lib.check(_lib.lib().qv_gemm_bf16(ctypes.byref(args), None), "gemm")
    with pytest.raises(RuntimeError, match="contiguous"):
        ops.minmax_accumulate(torch.randn(8, 8, device=dev).t(), ops.new_minmax(dev))
    with pytest.raises(RuntimeError, match="must be torch.float32"):
        ops.minmax_accumulate(torch.zeros(8, dtype=torch.float64, device=dev), ops.new_minmax(dev))
    launched = ops.launch_count() - n0
    # only the helper kernels above (split_planes, minmax reset) ran; no failing call launched anything
    assert launched <= 12
    # the library still works
    out = ops.gemm(Op.full(a), Op.full(b), 64, 32, 40, PAIRS_FP32)
    torch.cuda.synchronize()
    ref = (a[0].double() + a[1].double()) @ (b[0].double() + b[1].double()).t()
    assert float((out.double() - ref).abs().max()) < 1e-3


def test_engine_rejects_wrong_shapes_and_cpu_models(cuda_dev):
    import copy
    from parity_utils import build_models
    from qatvit_b200.engine import QATDistillStep
    vr, prepared, teacher = build_models("fbgemm", "vit_test_tiny", "vit_test_teacher", 64)
    with pytest.raises(RuntimeError, match="no CPU fallback|move the prepared model"):
        QATDistillStep(prepared, teacher, 2, dict(vr.DEFAULT_HPARAMS))              # CPU model
    step = QATDistillStep(copy.deepcopy(prepared).to(cuda_dev), copy.deepcopy(teacher).to(cuda_dev), 2, dict(vr.DEFAULT_HPARAMS))
    images, labels = vr.synthetic_batch(3, seed=1, img=64)
    with pytest.raises(RuntimeError, match="built for batch 2"):
        step(images.to(cuda_dev), labels.to(cuda_dev))
