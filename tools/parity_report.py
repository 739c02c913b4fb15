"""Tier-2 parity report for BASELINE.json configs[1] (ViT-B/16 teacher -> ViT-S/16 QAT student, batch 256, fbgemm qconfig),
forward of the prepared student ("report code-mismatch rate and max |dscale| / scale").

For every activation fake-quant module of the student (patch embed, 48 block Linears, head = 50 stages) it compares the integer
codes  q = clamp(rint(x / scale) + zero_point)  and the observer's scale between

  forced      : the reference path on the host CPU (oracle/vit_ref.py: stock prepare_qat + live ATen ops) fed OUR raw tensor at
                every fake-quant input (tests/parity_utils.py) vs this library -- identical inputs, so every code and every scale
                must agree exactly (mismatch rate 0, dscale 0);
  free-running: the same CPU reference left alone vs this library -- differences here are the chaos of re-quantisation
                (a 1-ulp GEMM difference flips a code, the flip moves the next layer's input by a full step);
  torch-cuda  : the same CPU reference vs STOCK torch CUDA eager (the reference's own GPU path) -- the floor any
                implementation with a different summation order sits on.

Usage (GPU box):  python tools/parity_report.py [--batch 256] [--out profiles/parity_config1.json]
Test infrastructure: imports oracle/ (the checker); nothing here is on the product path."""
import argparse
import copy
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def act_fq_modules(model):
    """name -> module for every ACTIVATION fake-quant of the prepared student that the engine keeps raw (same keys as
    parity_utils.engine_raw_tensors)."""
    out = {}
    for name, m in model.named_modules():
        if name.endswith("activation_post_process") and hasattr(m, "fake_quant_enabled") and "weight_fake_quant" not in name:
            if name.startswith("quant.") or name == "quant.activation_post_process":
                continue
            out[name] = m
    return out


def codes_of(x, scale, zp, qmin, qmax):
    import torch
    inv = 1.0 / scale.float()
    return torch.clamp(torch.round(x * inv) + zp.float(), qmin, qmax).to(torch.uint8)


def run_with_code_capture(model, images, names, device_store="cpu", forced=None):
    """Forward `model` once; returns {name: (codes uint8, scale float, zp int)} computed from the INPUT of each activation
    fake-quant and the qparams the module holds after its forward (the ones the fused op applied)."""
    import torch
    cap, handles = {}, []
    mods = act_fq_modules(model)
    for name in names:
        m = mods[name]

        def pre(mod, inp, name=name):
            x = inp[0]
            if forced is not None:
                ours = forced[name]
                ours = ours() if callable(ours) else ours
                x = ours.to(x.device)
                cap[name] = [x.detach()]
                return (x + (inp[0] - inp[0].detach()),)
            cap[name] = [x.detach()]
            return None

        def post(mod, inp, out, name=name):
            x = cap[name][0]
            sc, zp = mod.scale.detach().clone(), mod.zero_point.detach().clone()
            q = codes_of(x, sc, zp, mod.activation_post_process.quant_min, mod.activation_post_process.quant_max)
            cap[name] = (q.to(device_store), float(sc), int(zp))
        handles += [m.register_forward_pre_hook(pre), m.register_forward_hook(post)]
    with torch.no_grad():
        model(images)
    for h in handles:
        h.remove()
    return cap


def compare(a, b, dev):
    """per-stage mismatch statistics between two captures."""
    import torch
    worst_rate, worst_stage, tot_mis, tot_n, worst_ds, max_step = 0.0, None, 0, 0, 0.0, 0
    per = {}
    for name in a:
        qa, sa, za = a[name]
        qb, sb, zb = b[name]
        qa, qb = qa.to(dev).reshape(-1), qb.to(dev).reshape(-1)
        diff = (qa.to(torch.int16) - qb.to(torch.int16)).abs()
        mis = int((diff != 0).sum())
        n = qa.numel()
        ds = abs(sa - sb) / abs(sb)
        per[name] = {"mismatch_rate": mis / n, "max_code_diff": int(diff.max()), "dscale_rel": ds, "zero_point_equal": za == zb}
        tot_mis += mis
        tot_n += n
        max_step = max(max_step, int(diff.max()))
        if mis / n >= worst_rate:
            worst_rate, worst_stage = mis / n, name
        worst_ds = max(worst_ds, ds)
    return {"stages": len(per), "codes_compared": tot_n, "code_mismatch_rate": tot_mis / max(tot_n, 1), "worst_stage": worst_stage,
            "worst_stage_mismatch_rate": worst_rate, "max_code_diff": max_step, "max_dscale_over_scale": worst_ds,
            "zero_points_equal": all(v["zero_point_equal"] for v in per.values()),
            "first_stage": per[next(iter(per))], "last_stage": per[list(per)[-1]]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "parity_config1.json"))
    args = ap.parse_args()
    import torch
    import qatvit_b200  # noqa: F401
    from qatvit_b200.engine import QATDistillStep
    from oracle import vit_ref as vr
    from parity_utils import engine_raw_tensors
    dev = torch.device("cuda", 0)
    torch.set_num_threads(os.cpu_count() or 1)
    B = args.batch
    torch.manual_seed(0)
    student = vr.enable_qat(vr.make_student(prefer_reference=False), "fbgemm")
    teacher = vr.create_model("vit_base_patch16_224", num_classes=10).eval()
    images, labels = vr.synthetic_batch(B, seed=1)
    hp = dict(vr.DEFAULT_HPARAMS)
    t0 = time.time()
    # ---- ours (free-running == forced: the library never sees the reference) ----
    ours_model = copy.deepcopy(student).to(dev)
    step = QATDistillStep(ours_model, copy.deepcopy(teacher).to(dev), B, hp)
    step.predict(images.to(dev))
    torch.cuda.synchronize()
    raw = engine_raw_tensors(step.student_engine, lazy=True)
    names = list(raw.keys())
    mods = act_fq_modules(ours_model)
    assert set(names) == set(mods.keys()), (set(names) ^ set(mods.keys()))
    ours = {}
    for name in names:
        m = mods[name]
        x = raw[name]().to(dev)
        ours[name] = (codes_of(x, m.scale.detach(), m.zero_point.detach(), m.activation_post_process.quant_min,
                               m.activation_post_process.quant_max).cpu(), float(m.scale), int(m.zero_point))
    # ---- CPU reference, free-running and forced ----
    cpu_free = run_with_code_capture(copy.deepcopy(student), images, names)
    cpu_forced = run_with_code_capture(copy.deepcopy(student), images, names, forced=raw)
    # ---- stock torch CUDA eager (the reference's own GPU path) ----
    cuda_free = run_with_code_capture(copy.deepcopy(student).to(dev), images.to(dev), names)
    rep = {"workload": f"BASELINE configs[1]: ViT-S/16 QAT student forward, batch {B}, fbgemm qconfig, synthetic 224x224, first "
                       "observer step (min/max initialised from this batch)",
           "stages": "50 activation fake-quant modules (patch embed, 12 x (qkv, proj, fc1, fc2), head)",
           "forced__this_library_vs_cpu_reference": compare(ours, cpu_forced, dev),
           "free_running__this_library_vs_cpu_reference": compare(ours, cpu_free, dev),
           "free_running__stock_torch_cuda_vs_cpu_reference": compare(cuda_free, cpu_free, dev),
           "torch": torch.__version__, "device": torch.cuda.get_device_name(0), "host_threads": torch.get_num_threads(),
           "seconds": round(time.time() - t0, 1)}
    f = rep["forced__this_library_vs_cpu_reference"]
    rep["forced_is_bit_exact"] = bool(f["code_mismatch_rate"] == 0.0 and f["max_dscale_over_scale"] == 0.0 and f["zero_points_equal"])
    with open(args.out, "w") as fh:
        json.dump(rep, fh, indent=1)
    print(json.dumps({k: rep[k] for k in rep if k.endswith("reference") or k == "forced_is_bit_exact"}, indent=1)[:3000])
    if not rep["forced_is_bit_exact"]:
        raise SystemExit("forced parity is not bit-exact")


if __name__ == "__main__":
    main()
