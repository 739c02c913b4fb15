"""Debug helper: repeat the gradient-planes epilogue case and report where fused and unfused planes differ."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import qatvit_b200  # noqa: E402,F401
from qatvit_b200 import ops  # noqa: E402
from qatvit_b200.ops import Op, PAIRS_EXACT_B  # noqa: E402

dev = torch.device("cuda", 0)
M, N, K, gelu = 1576, 384, 384, False
g = torch.Generator().manual_seed(1)
a = torch.randn(M, K, generator=g).to(dev)
b = torch.randint(-128, 128, (N, K), generator=g).float().to(dev)
ap, bp = ops.split_planes(a), b.bfloat16()[None].contiguous()
y = (torch.randn(M, N, generator=g) * 2.0).to(dev)
fq = (torch.tensor([4.0 / 127], device=dev), torch.tensor([63], dtype=torch.int32, device=dev), 0, 127)
wsc = (torch.rand(N, generator=g) * 0.02 + 0.001).to(dev)
gmat = ops.gemm(Op.full(ap), Op.full(bp), M, N, K, PAIRS_EXACT_B)
part = torch.empty(-(-M // 64), N, device=dev)
ref = torch.empty(2, M, N, dtype=torch.bfloat16, device=dev)
ops.gp_planes(gmat, y, fq, wsc, True, gelu, M, N, ref, part, 64)
torch.cuda.synchronize()
bad = 0
for it in range(200):
    out = torch.full((2, M, N), float("nan"), dtype=torch.bfloat16, device=dev)
    slab = torch.empty(-(-M // 32), N, device=dev)
    ops.gemm(Op.full(ap), Op.full(bp), M, N, K, PAIRS_EXACT_B, out_planes=out, col_scale=wsc, grad_of=(y, fq, gelu, slab))
    gm2 = ops.gemm(Op.full(ap), Op.full(bp), M, N, K, PAIRS_EXACT_B)
    torch.cuda.synchronize()
    if not torch.equal(gm2, gmat):
        print(it, "PLAIN GEMM differs:", int((gm2 != gmat).sum()))
    d = out.view(torch.int16) != ref.view(torch.int16)
    if d.any():
        bad += 1
        idx = d.nonzero()
        rows, cols = idx[:, 1], idx[:, 2]
        print(it, "mismatches", int(d.sum()), "rows", int(rows.min()), int(rows.max()), "cols", int(cols.min()), int(cols.max()),
              "planes", idx[:, 0].unique().tolist(), "row%32 uniq", (rows % 32).unique().tolist()[:8], "col//32", (cols // 32).unique().tolist()[:12])
        r, c = int(rows[0]), int(cols[0])
        print("   first:", r, c, "fused", out[:, r, c].tolist(), "ref", ref[:, r, c].tolist(), "g", float(gmat[r, c]), "y", float(y[r, c]))
print("bad iterations:", bad, "of 200")
