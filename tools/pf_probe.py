"""Isolated A/B of the experimental L2-prefetching producer of the CTA-pair GEMM (make -C qat-vit_b200/csrc prefetch):
QV_GEMM_PAIR=51 (off) vs 115 (bit 6: on) inside one process.  Usage: python tools/pf_probe.py [N K]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("QV_LIB", os.path.join(ROOT, "qat-vit_b200", "lib", "libqatvit_b200_pf.so"))
import torch  # noqa: E402
import qatvit_b200  # noqa: E402,F401
from qatvit_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
M = 50432
shapes = [(int(sys.argv[1]), int(sys.argv[2]))] if len(sys.argv) >= 3 else [(2304, 768), (768, 3072)]
for N, K in shapes:
    g = torch.Generator().manual_seed(0)
    am = ops.split_planes_mix((torch.randn(M, K, generator=g) * 1.3).to(dev))
    wm = ops.split_planes_mix((torch.randn(N, K, generator=g) * 0.02).to(dev), weight=True)
    out = torch.empty(M, N, device=dev)
    res = {}
    ref = None
    for mode in ("51", "115", "51", "115"):
        os.environ["QV_GEMM_PAIR"] = mode
        for _ in range(2):
            ops.gemm(ops.Op.full(am), ops.Op.full(wm), M, N, K, (2, 2), out=out, mix=True)
        st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        st.record()
        for _ in range(10):
            ops.gemm(ops.Op.full(am), ops.Op.full(wm), M, N, K, (2, 2), out=out, mix=True)
        en.record()
        torch.cuda.synchronize()
        res.setdefault(mode, []).append(st.elapsed_time(en) * 100)
        if ref is None:
            ref = out.clone()
        assert torch.equal(out, ref)
    print(f"N={N} K={K}: no prefetch {res['51']} us | L2 prefetch {res['115']} us (bit-identical outputs)", flush=True)
