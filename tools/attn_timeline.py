"""GPU probe: per-event clock64 timeline of CTA 0 of the attention backward kernel (needs `make -C qat-vit_b200/csrc debug`;
loads qat-vit_b200/lib/libqatvit_b200_dbg.so through QV_LIB).  Prints the MMA thread's and one compute warp's events for one item."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("QV_LIB", os.path.join(ROOT, "qat-vit_b200", "lib", "libqatvit_b200_dbg.so"))
import torch  # noqa: E402
import qatvit_b200  # noqa: E402,F401
from qatvit_b200 import _lib, ops  # noqa: E402

TAGS = {1: "mma: wait ld_full", 2: "mma: ld_full ok", 3: "mma: wait cmp_done", 4: "mma: cmp_done ok", 5: "mma: epi ok, issue MMA2",
        6: "mma: MMA2 issued", 10: "cmp: item start", 11: "cmp: delta done", 12: "cmp: wait mma1_done", 13: "cmp: mma1_done ok",
        14: "cmp: chunk done", 15: "out: acc_done ok", 16: "out: sub-pass out done"}


def main():
    dev = torch.device("cuda", 0)
    B, H, T = 148 * 3 // 6 + 1, 6, 197
    D = H * 64
    torch.manual_seed(0)
    sval = 0.0437
    y_raw = (torch.randint(-60, 68, (B * T, 3 * D), device=dev).float() + 0.3 * torch.randn(B * T, 3 * D, device=dev)) * sval
    s = torch.tensor([sval], device=dev)
    fq = (s, torch.tensor([60], dtype=torch.int32, device=dev), 0, 127)
    cp = torch.empty(1, B * T, 3 * D, dtype=torch.bfloat16, device=dev)
    ops.act_planes(y_raw, fq, False, cp, codes_only=True)
    out = torch.empty(2, B * T, D, dtype=torch.bfloat16, device=dev)
    lse = torch.empty(B * H * T, device=dev)
    ops.attn_fwd(cp, B, T, H, 0.125, out, qk_scale=s, v_scale=s, lse=lse)
    dOp = ops.split_planes(torch.randn(B * T, D, device=dev))
    g_qkv = torch.empty(B * T, 3 * D, device=dev)
    wsc = torch.rand(3 * D, device=dev) * 0.02 + 0.001
    planes = torch.empty(2, B * T, 3 * D, dtype=torch.bfloat16, device=dev)
    slab = torch.empty(B * 2 * 4, 3 * D, device=dev)
    fused = os.environ.get("QV_FUSED", "0") == "1"
    if fused:
        ops.attn_bwd_gp(cp, s, out, dOp, lse, B, T, H, 0.125, y_raw, fq, wsc, planes, slab)
    else:
        ops.attn_bwd(cp, s, out, dOp, lse, B, T, H, 0.125, g_qkv)
    torch.cuda.synchronize()
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
    starts = [i for i, e in enumerate(ev) if e[1] == 10]      # compute warps: item starts of CTA 0
    t0 = ev[starts[1]][0]
    end = starts[2] if len(starts) > 2 else len(ev)
    prev = t0
    for t, tag, who in ev[starts[1]:min(end + 3, len(ev))]:
        print(f"{(t - t0) / 1.9:9.0f} ns  (+{(t - prev) / 1.9:6.0f})  {TAGS.get(tag, tag)}")
        prev = t
    print("items timed:", len(starts))


if __name__ == "__main__":
    main()
