import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import qatvit_b200
from test_int8_gpu import _run_ours
from oracle import fq_oracle as fo
z = np.load(os.path.join(ROOT, "tests/golden/qlinear_cases.npz"))
dev = torch.device("cuda")
for i in range(2):
    sx, zx, sy, zy = z[f"q{i}_p"]
    qx, qw, sw, b = [torch.from_numpy(z[f"q{i}_{k}"]) for k in ("qx", "qw", "sw", "b")]
    qy, y = _run_ours(dev, qx, float(sx), int(zx), qw, sw, b, float(sy), int(zy))
    ref = torch.from_numpy(z[f"q{i}_qy"])
    bad = (qy != ref).nonzero()
    print("case", i, "mismatches", len(bad), "of", ref.numel())
    acc = (qx.int() - int(zx)) @ qw.int().t()
    for (m, n) in bad[:10].tolist():
        swn = float(sw[n] if sw.numel() > 1 else sw[0])
        bs = np.float32(sx) * np.float32(swn)
        print(m, n, "ours", int(qy[m, n]), "ref", int(ref[m, n]), "acc", int(acc[m, n]), "bq", float(b[n]) / bs,
              "real", (int(acc[m, n]) + round(float(b[n]) / bs)) * (bs / np.float32(sy)) + zy)
    print("cols with mismatch:", sorted(set(bad[:, 1].tolist())), "rows:", sorted(set(bad[:, 0].tolist()))[:20])
