"""Debug helper (GPU): per-parameter gradient errors of the fused step vs the CPU reference + phase timings."""
import copy, sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
t0 = time.time()
import torch
print("import torch", time.time() - t0, flush=True)
import qatvit_b200
from qatvit_b200.engine import QATDistillStep
from oracle import vit_ref as vr

def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))

backend = sys.argv[1] if len(sys.argv) > 1 else "fbgemm"
sname, tname, img, B = ("vit_test_tiny", "vit_test_teacher", 64, 4)
if len(sys.argv) > 2 and sys.argv[2] == "full":
    sname, tname, img, B = ("vit_small_patch16_224", "vit_base_patch16_224", 224, 8)
torch.manual_seed(0)
kw = dict(img_size=img) if img != 224 else {}
student = vr.qat_wrapper_cls(prefer_reference=False)(vr.create_model(sname, num_classes=10, **kw))
torch.manual_seed(1)
teacher = vr.create_model(tname, num_classes=10, **kw)
with torch.no_grad():
    teacher.head.weight.mul_(8.0)
    for p in student.parameters():
        if p.dim() == 1:
            p.add_(0.02 * torch.randn_like(p))
teacher.eval()
for p in teacher.parameters():
    p.requires_grad = False
prepared = vr.enable_qat(student, backend)
images, labels = vr.synthetic_batch(B, seed=3, img=img)
hp = dict(vr.DEFAULT_HPARAMS)
t0 = time.time()
dev = torch.device("cuda")
gs, gt = copy.deepcopy(prepared).to(dev), copy.deepcopy(teacher).to(dev)
torch.cuda.synchronize(); print("to cuda", time.time() - t0, flush=True)
t0 = time.time()
step = QATDistillStep(gs, gt, B, hp)
torch.cuda.synchronize(); print("engine build", time.time() - t0, flush=True)
for it in range(2):
    t0 = time.time()
    loss_ref, s_ref, t_ref = vr.distill_step(prepared, teacher, images, labels, None, hp, clip=False)
    print("cpu ref step", time.time() - t0, flush=True)
    ref_grads = {n: p.grad.clone() for n, p in prepared.named_parameters()}
    prepared.zero_grad(set_to_none=True)
    t0 = time.time()
    out3 = step(images.to(dev), labels.to(dev))
    torch.cuda.synchronize(); print("gpu step", time.time() - t0, flush=True)
    print("loss", float(out3[0]), float(loss_ref), "teacher rel", rel(step.teacher_engine.logits, t_ref))
    def l2(a, b):
        a, b = a.detach().double().cpu(), b.detach().double().cpu()
        return float((a - b).norm() / b.norm().clamp_min(1e-30))
    errs = sorted(((rel(p.grad, ref_grads[n]), l2(p.grad, ref_grads[n]), n) for n, p in gs.named_parameters()), reverse=True)
    for e, e2, n in errs[:10]:
        print(f"  max-rel {e:.3e}  l2-rel {e2:.3e} {n}")
    allg = torch.cat([p.grad.detach().flatten().cpu() for _, p in gs.named_parameters()])
    allr = torch.cat([ref_grads[n].flatten() for n, _ in gs.named_parameters()])
    print("  global grad l2-rel", float((allg - allr).norm() / allr.norm()), "grad norm", float(allr.norm()))
    hd = step.student_engine.head
    print("  logits raw gpu", step.student_logits_raw[0].cpu().tolist()[:5], "ref", s_ref[0].tolist()[:5])
    sd_r, sd_g = prepared.state_dict(), gs.state_dict()
    bad = [(k, rel(sd_g[k].float(), sd_r[k].float())) for k in sd_r if k.endswith(("scale", "min_val", "max_val")) and "weight_fake" not in k]
    bad.sort(key=lambda x: -x[1])
    print("  act observer worst:", bad[:3])
    wbad = [k for k in sd_r if "weight_fake_quant" in k and not torch.equal(sd_g[k].cpu(), sd_r[k])]
    print("  weight observer mismatches:", wbad[:5])

# ---- layer-local forward comparison (raw inputs of every activation fake-quant, in call order) ----
print("---- forward trace ----")
raws = {}
hooks = []
for name, m in prepared.named_modules():
    if name.endswith("activation_post_process") and not name.endswith("activation_post_process.activation_post_process") \
            and "weight_fake_quant" not in name:
        hooks.append(m.register_forward_pre_hook(lambda mod, inp, name=name: raws.__setitem__(name, inp[0].detach().clone())))
blk_in = {}
for i, b in enumerate(prepared.model.blocks):
    hooks.append(b.register_forward_pre_hook(lambda mod, inp, i=i: blk_in.__setitem__(i, inp[0].detach().clone())))
with torch.no_grad():
    prepared(images)
    step.student_engine.forward(images.to(dev), labels.to(dev), step.teacher_engine.logits)
torch.cuda.synchronize()
se = step.student_engine
def l2(a, b):
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
D = se.d.D
print("p_raw", l2(se.p_raw.view(B, se.d.P, D), raws["model.patch_embed.proj.activation_post_process"].flatten(2).transpose(1, 2)))
for i in range(se.d.L):
    pre = f"model.blocks.{i}."
    print(i, "x_in %.2e" % l2(se.x_in[i].view(B, se.d.T, D), blk_in[i]),
          "qkv %.2e" % l2(se.qkv_raw[i], raws[pre + "attn.qkv.activation_post_process"]),
          "proj %.2e" % l2(se.a_raw[i], raws[pre + "attn.proj.activation_post_process"]),
          "fc1 %.2e" % l2(se.f_raw[i], raws[pre + "mlp.fc1.activation_post_process"]),
          "fc2 %.2e" % l2(se.m_raw[i], raws[pre + "mlp.fc2.activation_post_process"]))
print("logits_raw", l2(se.logits_raw, raws["model.head.activation_post_process"]))

# ---- noise floor: the reference's own CUDA path (stock torch eager) vs its CPU path ----
print("---- stock torch CUDA vs CPU (reference vs reference) ----")
torch.manual_seed(0)
student2 = vr.qat_wrapper_cls(prefer_reference=False)(vr.create_model(sname, num_classes=10, **kw))
with torch.no_grad():
    for p in student2.parameters():
        if p.dim() == 1:
            p.add_(0.02 * torch.randn_like(p))
# same weights as `prepared` started from: copy parameters
p2 = vr.enable_qat(student2, backend)
p2.load_state_dict({k: v for k, v in prepared.state_dict().items() if "activation_post_process" not in k and "weight_fake_quant" not in k}, strict=False)
p2_cpu = copy.deepcopy(p2)
p2_gpu = copy.deepcopy(p2).to(dev)
lc, sc, _ = vr.distill_step(p2_cpu, teacher, images, labels, None, hp, clip=False)
lg, sg, _ = vr.distill_step(p2_gpu, gt, images.to(dev), labels.to(dev), None, hp, clip=False)
print("loss cpu/gpu", float(lc), float(lg), "logits l2", l2(sg, sc))
gc = torch.cat([p.grad.flatten() for p in p2_cpu.parameters()])
gg = torch.cat([p.grad.flatten().cpu() for p in p2_gpu.parameters()])
print("global grad l2-rel (torch cuda vs torch cpu)", float((gg - gc).norm() / gc.norm()))

# ---- forced parity: the CPU reference re-run with OUR raw tensors substituted at every activation fake-quant ----
print("---- forced (teacher-forced) parity ----")
for h in hooks:
    h.remove()
torch.manual_seed(0)
ref3 = copy.deepcopy(p2)      # fresh observers, same weights
gs3 = copy.deepcopy(p2).to(dev)
step3 = QATDistillStep(gs3, gt, B, hp)
out3 = step3(images.to(dev), labels.to(dev))
torch.cuda.synchronize()
se3 = step3.student_engine
G = int(se3.d.P ** 0.5)
forced = {"model.patch_embed.proj.activation_post_process": se3.p_raw.view(B, se3.d.P, D).transpose(1, 2).reshape(B, D, G, G).cpu(),
          "model.head.activation_post_process": se3.logits_raw.cpu()}
for i in range(se3.d.L):
    pre = f"model.blocks.{i}."
    forced[pre + "attn.qkv.activation_post_process"] = se3.qkv_raw[i].view(B, se3.d.T, -1).cpu()
    forced[pre + "attn.proj.activation_post_process"] = se3.a_raw[i].view(B, se3.d.T, -1).cpu()
    forced[pre + "mlp.fc1.activation_post_process"] = se3.f_raw[i].view(B, se3.d.T, -1).cpu()
    forced[pre + "mlp.fc2.activation_post_process"] = se3.m_raw[i].view(B, se3.d.T, -1).cpu()
stage_err = {}
def mk(name):
    def pre_hook(mod, inp):
        x = inp[0]
        ours = forced[name]
        stage_err[name] = l2(x, ours)
        return (x + (ours - x).detach(),)
    return pre_hook
for name, m in ref3.named_modules():
    if name in forced:
        m.register_forward_pre_hook(mk(name))
l3, s3, _ = vr.distill_step(ref3, teacher, images, labels, None, hp, clip=False)
print("loss gpu/forced-ref", float(out3[0]), float(l3), "worst stage err", max(stage_err.values()), max(stage_err, key=stage_err.get))
rg = {n: p.grad for n, p in ref3.named_parameters()}
errs = sorted(((rel(p.grad, rg[n]), l2(p.grad, rg[n]), n) for n, p in gs3.named_parameters()), reverse=True)
for e, e2, n in errs[:6]:
    print(f"  max-rel {e:.3e}  l2-rel {e2:.3e} {n}")
sd_r, sd_g = ref3.state_dict(), gs3.state_dict()
bad = [(k, rel(sd_g[k].float(), sd_r[k].float())) for k in sd_r if k.endswith(("scale", "min_val", "max_val", "zero_point"))]
bad.sort(key=lambda x: -x[1])
print("  observer worst:", bad[:3])
