"""GPU probe: only the fc2 dgrad + gradient-planes epilogue GEMM (EPI 2) at the bench shape, a few launches -- the target of an
`ncu --set full --import-source on -k regex:qv_gemm_kernel -s 3 -c 1` capture."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import qatvit_b200  # noqa: E402,F401
from qatvit_b200 import ops  # noqa: E402
from qatvit_b200.ops import Op  # noqa: E402

dev = torch.device("cuda", 0)
M, N, K = 50432, 384, 1536          # fc2: out 384, in 1536; dgrad output width = 1536
g = torch.Generator().manual_seed(0)
gpl = ops.split_planes(torch.randn(M, N, generator=g).to(dev) * 1e-3)
codes_t = torch.randint(-127, 128, (1, K, N), generator=g).to(dev).bfloat16()
y_raw = torch.randn(M, K, generator=g).to(dev)
sc, zp = torch.tensor([0.05], device=dev), torch.tensor([64], dtype=torch.int32, device=dev)
gpo = torch.empty(2, M, K, dtype=torch.bfloat16, device=dev)
part = torch.empty(-(-M // 32), K, device=dev)
cs1 = (torch.rand(K, generator=g) * 0.01 + 0.001).to(dev)
for _ in range(int(os.environ.get("QV_N", "6"))):
    ops.gemm(Op.full(gpl), Op.full(codes_t), M, K, N, (2, 1), out_planes=gpo, col_scale=cs1, grad_of=(y_raw, (sc, zp, 0, 127), True, part))
torch.cuda.synchronize()
print("done")
