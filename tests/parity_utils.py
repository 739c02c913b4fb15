"""Shared helpers for the model-level parity tests.

Why "forced" parity: fake-quant rounding is discontinuous, so two correct implementations whose GEMM summation
order differs by 1 ulp diverge end-to-end (a perturbation eps becomes ~sqrt(eps) after one re-quantisation).  The
reference's own CUDA path differs from its own CPU path by ~6e-2 on ViT-S logits (measured).
To compare arithmetic rather than chaos, the CPU reference is re-run with OUR raw (pre-fake-quant) tensor
substituted, straight-through, at the input of every activation fake-quant module: all integer codes / STE masks
are then decided on identical inputs, and every stage, the loss and every gradient must agree to float accuracy."""
import copy

import torch


def rel_max(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def rel_l2(a, b):
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def build_models(backend, student_name, teacher_name, img, seed=0, ln_variant="subclass"):
    from oracle import vit_ref as vr
    torch.manual_seed(seed)
    kw = dict(img_size=img) if img != 224 else {}
    student = vr.qat_wrapper_cls(prefer_reference=False)(vr.create_model(student_name, num_classes=10, ln_variant=ln_variant, **kw))
    torch.manual_seed(seed + 1)
    teacher = vr.create_model(teacher_name, num_classes=10, **kw)
    with torch.no_grad():
        teacher.head.weight.mul_(8.0)
        for p in student.parameters():      # perturb LN gains / biases so every gradient path is exercised
            if p.dim() == 1:
                p.add_(0.02 * torch.randn_like(p))
    teacher.eval()
    for p in teacher.parameters():
        p.requires_grad = False
    prepared = vr.enable_qat(student, backend)
    return vr, prepared, teacher


def engine_raw_tensors(se, lazy=False):
    """name of each activation fake-quant module of the prepared student -> the engine's raw input for it (CPU).
    lazy=True: values are zero-argument callables that copy the tensor off the device when the hook needs it (batch 256:
    8 GB of raw tensors would otherwise sit in host memory at once)."""
    d = se.d
    B, T, D, P = d.B, d.T, d.D, d.P
    G = int(round(P ** 0.5))
    get = {"model.patch_embed.proj.activation_post_process":
           lambda: se.p_raw.view(B, P, D).transpose(1, 2).reshape(B, D, G, G).cpu(),
           "model.head.activation_post_process": lambda: se.logits_raw.cpu()}

    def tok(t):
        return lambda: t.view(B, T, -1).cpu()
    for i in range(d.L):
        pre = f"model.blocks.{i}."
        get[pre + "attn.qkv.activation_post_process"] = tok(se.qkv_raw[i])
        get[pre + "attn.proj.activation_post_process"] = tok(se.a_raw[i])
        get[pre + "mlp.fc1.activation_post_process"] = tok(se.f_raw[i])
        get[pre + "mlp.fc2.activation_post_process"] = tok(se.m_raw[i])
        if se.ln_obs:
            get[pre + "norm1.activation_post_process"] = tok(se.h1_raw[i])
            get[pre + "norm2.activation_post_process"] = tok(se.h2_raw[i])
    if se.ln_obs:
        get["model.norm.activation_post_process"] = tok(se.hN_raw)
    return get if lazy else {k: f() for k, f in get.items()}


def install_forcing_hooks(ref_model, forced, stage_err, keep_ref_codes=None):
    """Pre-hooks on the activation fake-quant modules of the CPU reference: the module sees OUR raw tensor, value-exact
    (`ours + (x - x.detach())`: the bracket is exactly zero, so the forward value is `ours` bit for bit while the gradient still
    flows into the reference's own x) -- every observer and every integer code of the reference is then decided on the very
    tensor our kernels saw, which is what lets the observer state be compared with torch.equal.
    keep_ref_codes: optional dict filled with name -> (input tensor as forced, module) for code-level comparisons."""
    handles = []

    def mk(name):
        def pre_hook(mod, inp):
            x = inp[0]
            ours = forced[name]
            if callable(ours):
                ours = ours()
            stage_err[name] = rel_l2(x, ours)
            if keep_ref_codes is not None:
                keep_ref_codes[name] = ours
            return (ours + (x - x.detach()),)
        return pre_hook
    for name, m in ref_model.named_modules():
        if name in forced:
            handles.append(m.register_forward_pre_hook(mk(name)))
    return handles
