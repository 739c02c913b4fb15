"""GPU probe: clock64 timeline of CTA 0 of the fused attention FORWARD (needs `make -C qat-vit_b200/csrc debug`; loads
qat-vit_b200/lib/libqatvit_b200_dbg.so through QV_LIB).  QV_NPL=1: integer-code student operands; 2: hi/lo planes (teacher).
Prints one item's events of the MMA warp and of one softmax warp of each query tile, plus the launch's wall time."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("QV_LIB", os.path.join(ROOT, "qat-vit_b200", "lib", "libqatvit_b200_dbg.so"))
import torch  # noqa: E402
import qatvit_b200  # noqa: E402,F401
from qatvit_b200 import _lib, ops  # noqa: E402

TAGS = {21: "mma: wait qk_full", 22: "mma: qk_full ok", 23: "mma: S issued", 24: "mma: p_ready[0] ok", 25: "mma: issue PV0",
        26: "mma: PV0 issued", 27: "mma: p_ready[1] ok", 28: "mma: o_full[0] ok, issue PV1", 29: "mma: PV1 issued",
        30: "wait s_full", 31: "s_full ok", 32: "pass 1 (max) done", 33: "pass 2 (exp) done, p_ready", 34: "o_full ok",
        35: "output done", 36: "O in registers, tmem_free", 37: "O normalised + split", 38: "hi plane stored", 39: "lo plane stored"}


def main():
    dev = torch.device("cuda", 0)
    npl = int(os.environ.get("QV_NPL", "1"))
    H = 6 if npl == 1 else 12
    B, T = 256, 197
    D = H * 64
    torch.manual_seed(0)
    if npl == 1:
        s = torch.tensor([0.0437], device=dev)
        qkv = torch.randint(-60, 68, (1, B * T, 3 * D), device=dev).to(torch.bfloat16)
        kw = dict(qk_scale=s, v_scale=s)
    else:
        qkv = ops.split_planes(torch.randn(B * T, 3 * D, device=dev))
        kw = {}
    out = torch.empty(2, B * T, D, dtype=torch.bfloat16, device=dev)
    lse = torch.empty(B * H * T, device=dev)
    for _ in range(2):
        ops.attn_fwd(qkv, B, T, H, 0.125, out, lse=lse, **kw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.attn_fwd(qkv, B, T, H, 0.125, out, lse=lse, **kw)
    e1.record()
    torch.cuda.synchronize()
    print(f"NPL {npl}: {e0.elapsed_time(e1) / 5 * 1e3:.1f} us per launch, {B * H} items, {B * H / 148:.1f} per CTA")
    L = _lib.lib()
    buf = (ctypes.c_ulonglong * (3 * 8192))()
    L.qv_debug_read.restype = ctypes.c_int
    assert L.qv_debug_read(buf) == 0
    ev = []
    for who in range(3):
        for i in range(8192):
            v = buf[who * 8192 + i]
            if v == 0:
                break
            ev.append((v & 0xffffffffffff, v >> 48, who))
    ev.sort()
    starts = [i for i, e in enumerate(ev) if e[1] == 21]
    k = min(len(starts) - 2, 30)
    t0 = ev[starts[k]][0]
    prev = t0
    for t, tag, who in ev[starts[k]:starts[k + 1] + 4]:
        name = ("        tile0: " if who == 1 else "                    tile1: " if who == 2 else "") + TAGS.get(tag, str(tag))
        print(f"{(t - t0) / 1.9:9.0f} ns  (+{(t - prev) / 1.9:6.0f})  {name}")
        prev = t


if __name__ == "__main__":
    main()
