"""oracle/int8_ref.py -- TEST INFRASTRUCTURE ONLY.  CPU side of BASELINE configs[4]: the converted student run with STOCK
torch ops (torch.quantize_per_tensor, the converted nnq.Linear / nnq.Conv2d modules = torch.ops.quantized.linear / conv2d on
the reference's CPU engine, F.layer_norm, F.scaled_dot_product_attention, F.gelu) and the float glue defined in SURVEY.md
§8c -- the reference's own converted forward cannot run, so the glue is ours and identical on both sides.
Follows ref/src/training/qat_trainer.py:377-380 (convert + evaluate_quantized_cpu) for what is being executed."""
import torch
import torch.nn.functional as F


def dynamic_qparams(x: torch.Tensor):
    """per-tensor affine quint8 0..255 from the batch min/max (torch/ao/quantization/observer.py:349-427 formula)."""
    mn = torch.clamp(x.min(), max=0.0)
    mx = torch.clamp(x.max(), min=0.0)
    scale = torch.clamp((mx - mn) / 255.0, min=torch.finfo(torch.float32).eps)
    zp = torch.clamp(0 - torch.round(mn / scale), 0, 255)
    return float(scale), int(zp)


def _q(x):
    s, z = dynamic_qparams(x)
    return torch.quantize_per_tensor(x, s, z, torch.quint8)


@torch.no_grad()
def converted_forward(converted, images: torch.Tensor, trace=None) -> torch.Tensor:
    """trace (optional dict): collects the quint8 input / output of every quantized module for per-layer parity checks."""
    vit = converted.model
    B = images.shape[0]

    def run(name, mod, qx):
        qy = mod(qx)
        if trace is not None:
            trace[name] = (qx, qy)
        return qy.dequantize()

    qx = converted.quant(images)                                   # nnq.Quantize, static qparams
    qp = vit.patch_embed.proj(qx)                                  # nnq.Conv2d
    if trace is not None:
        trace["patch_embed.proj"] = (qx, qp)
    x = qp.dequantize().flatten(2).transpose(1, 2)
    x = torch.cat([vit.cls_token.expand(B, -1, -1), x], dim=1) + vit.pos_embed
    for i, blk in enumerate(vit.blocks):
        h = F.layer_norm(x, (x.shape[-1],), blk.norm1.weight, blk.norm1.bias, blk.norm1.eps)
        qkv = run(f"blocks.{i}.attn.qkv", blk.attn.qkv, _q(h))
        Bq, N, _ = qkv.shape
        H = blk.attn.num_heads
        q, k, v = qkv.reshape(Bq, N, 3, H, -1).permute(2, 0, 3, 1, 4).unbind(0)
        o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(Bq, N, -1)
        x = x + run(f"blocks.{i}.attn.proj", blk.attn.proj, _q(o))
        h = F.layer_norm(x, (x.shape[-1],), blk.norm2.weight, blk.norm2.bias, blk.norm2.eps)
        f = run(f"blocks.{i}.mlp.fc1", blk.mlp.fc1, _q(h))
        x = x + run(f"blocks.{i}.mlp.fc2", blk.mlp.fc2, _q(F.gelu(f)))
    xn = F.layer_norm(x, (x.shape[-1],), vit.norm.weight, vit.norm.bias, vit.norm.eps)[:, 0]
    return run("head", vit.head, _q(xn))
