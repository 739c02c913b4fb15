"""GPU parity of the fused clip + AdamW (qv_clip_adamw) against torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW --
the two calls of the reference loop (ref/src/training/qat_trainer.py:360-361) -- on identical parameters and gradients.
Tolerance: 2e-6 relative on parameters after 4 steps (fp32 op-order differences of the foreach path)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("max_norm,grad_mag", [(1.0, 5.0), (1.0, 1e-3), (0.0, 1.0)])
def test_clip_adamw_matches_torch(cuda_dev, max_norm, grad_mag):
    from qatvit_b200.optim import FusedClipAdamW
    g = torch.Generator().manual_seed(0)
    shapes = [(1152, 384), (384,), (10, 384), (1, 1, 384), (7,), (3, 5, 16, 16)]
    ref_params = [torch.nn.Parameter((torch.randn(s, generator=g) * 0.05).to(cuda_dev)) for s in shapes]
    our_params = [torch.nn.Parameter(p.detach().clone()) for p in ref_params]
    total = sum(p.numel() for p in our_params)
    arena = torch.zeros(total, device=cuda_dev)
    ours = FusedClipAdamW(our_params, arena, lr=7.5e-5, weight_decay=1e-3, max_norm=max_norm)
    ref = torch.optim.AdamW(ref_params, lr=7.5e-5, weight_decay=1e-3)
    ids = [id(p) for p in our_params]
    for it in range(4):
        grads = [(torch.randn(s, generator=g) * grad_mag).to(cuda_dev) for s in shapes]
        off = 0
        for p, gr in zip(ref_params, grads):
            p.grad = gr.clone()
            arena[off:off + gr.numel()].copy_(gr.reshape(-1) * 2.0)       # "summed over 2 ranks": grad_scale = 0.5 undoes it
            off += gr.numel()
        if max_norm > 0:
            ref_norm = torch.nn.utils.clip_grad_norm_(ref_params, max_norm)
        else:
            ref_norm = torch.linalg.vector_norm(torch.cat([p.grad.reshape(-1) for p in ref_params]))
        ref.step()
        norm = ours.step(grad_scale=0.5, write_back_grad=True)
        torch.cuda.synchronize()
        assert abs(float(norm) - float(ref_norm)) <= 1e-5 * float(ref_norm)
        off = 0
        for p, q in zip(our_params, ref_params):
            assert float((p.detach() - q.detach()).abs().max()) <= 2e-6 * float(q.detach().abs().max()) + 1e-9, it
            n = q.numel()
            assert torch.allclose(arena[off:off + n].view_as(q), q.grad, rtol=1e-5, atol=1e-9)    # clipped grads written back
            off += n
    assert [id(p) for p in our_params] == ids                     # Parameter objects unchanged (state_dict / modules keep working)
    assert all(p.data_ptr() >= ours.flat.data_ptr() and p.data_ptr() < ours.flat.data_ptr() + 4 * total for p in our_params)
