"""GPU probe: only the teacher's fc1 GEMM (mixed operands, GELU + mixed planes out; CTA pairs, 256-wide tiles) at the bench shape --
the target of an `ncu --set full --import-source on -k regex:qv_gemm_kernel -s 3 -c 1` capture.  QV_FC1_PLAIN=1: fp32 output instead."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import qatvit_b200  # noqa: E402,F401
from qatvit_b200 import ops  # noqa: E402
from qatvit_b200.ops import Op  # noqa: E402

dev = torch.device("cuda", 0)
M, N, K = 50432, int(os.environ.get("QV_N_OUT", 3072)), int(os.environ.get("QV_K_IN", 768))
g = torch.Generator().manual_seed(0)
a = (torch.randn(M, K, generator=g) * 1.3).to(dev)
w = (torch.randn(N, K, generator=g) * 0.02).to(dev)
bias = torch.randn(N, generator=g).to(dev)
am, wm = ops.split_planes_mix(a), ops.split_planes_mix(w, weight=True)
out = torch.empty(M, N, device=dev)
planes = torch.empty(2, M, N, dtype=torch.bfloat16, device=dev)
plain = os.environ.get("QV_FC1_PLAIN", "0") != "0"
gelu = os.environ.get("QV_FC1_GELU", "1") != "0"
omix = os.environ.get("QV_FC1_OMIX", "1") != "0"
st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(int(os.environ.get("QV_N", "6"))):
    if i == 3:
        st.record()
    if plain:
        ops.gemm(Op.full(am), Op.full(wm), M, N, K, (2, 2), bias=bias, out=out, mix=True)
    else:
        ops.gemm(Op.full(am), Op.full(wm), M, N, K, (2, 2), bias=bias, out_planes=planes, gelu=gelu, mix=True, out_mix=omix)
en.record()
torch.cuda.synchronize()
print("done", st.elapsed_time(en) / 3 * 1e3, "us per launch")
