"""Multi-step training trajectory of the fused path against the reference loop on CPU.

The single-step tests (tests/test_engine_gpu.py) pin the arithmetic on identical inputs; this one lets both sides run FREE for 40
optimizer steps of the reference's hot loop (ref/src/training/qat_trainer.py:333-361: teacher forward, student fake-quant forward,
KL + CE, backward, clip_grad_norm_(1.0), AdamW) on the same two alternating batches and the same initial weights:

  * ours: QATDistillStep + FusedClipAdamW (the configuration bench.py times), observers + EMA + weights all updated on the device;
  * reference: oracle/vit_ref.distill_step with torch.optim.AdamW on the CPU (stock torch.ao modules).

Integer codes flip chaotically between two correct implementations, so the comparison is on what training is
for -- the loss curve: every step's loss within 2 % of the reference's (measured 0.25-0.31 %), both curves falling
by the same amount, and the trained observer ranges / weights of the two students close in aggregate.  Tolerances are written
next to each assertion."""
import copy

import pytest
import torch

from parity_utils import build_models, rel_l2

pytestmark = pytest.mark.gpu

STEPS = 40


@pytest.mark.parametrize("backend", ["fbgemm", "qnnpack"])
def test_forty_step_trajectory_tracks_the_cpu_reference(cuda_dev, backend):
    from qatvit_b200.engine import QATDistillStep
    from qatvit_b200.optim import FusedClipAdamW
    vr, prepared, teacher = build_models(backend, "vit_test_tiny", "vit_test_teacher", 64)
    hp = dict(vr.DEFAULT_HPARAMS)
    B = 8
    batches = [vr.synthetic_batch(B, seed=20 + i, img=64) for i in range(2)]

    gpu_student = copy.deepcopy(prepared).to(cuda_dev)
    step = QATDistillStep(gpu_student, copy.deepcopy(teacher).to(cuda_dev), B, hp)
    opt = FusedClipAdamW(gpu_student.parameters(), step.grad_arena, lr=float(hp["lr"]) * 0.5,
                         weight_decay=float(hp["weight_decay"]), max_norm=1.0)
    dev_batches = [(im.to(cuda_dev), lb.to(cuda_dev)) for im, lb in batches]
    ours = []
    for it in range(STEPS):
        out3 = step(*dev_batches[it % 2])
        opt.step()
        ours.append(out3[0].clone())               # no host sync inside the loop: read the losses at the end
    torch.cuda.synchronize()
    ours = [float(v) for v in ours]

    ref_opt = vr.make_optimizer(prepared.parameters(), hp, 0.5)
    ref = []
    for it in range(STEPS):
        im, lb = batches[it % 2]
        loss, _, _ = vr.distill_step(prepared, teacher, im, lb, ref_opt, hp, clip=True)
        ref.append(float(loss))

    assert all(v == v and abs(v) < 1e4 for v in ours), ours
    # (1) step 0 sees identical weights and inputs: only code flips separate the two losses (same bound as the single-step test)
    assert abs(ours[0] - ref[0]) <= 1e-2 * abs(ref[0]), (ours[0], ref[0])
    # (2) the curves stay together: 2 % of the reference loss at every step (two CPU runs whose initial weights differ by 3e-4
    #     relative drift apart by 1.2 % over these 40 steps; measured GPU vs CPU 2.5e-3 / 3.1e-3: profiles/r02bh_trajectory.log)
    worst = max(range(STEPS), key=lambda i: abs(ours[i] - ref[i]) / abs(ref[i]))
    assert abs(ours[worst] - ref[worst]) <= 2e-2 * abs(ref[worst]), (worst, ours[worst], ref[worst], ours, ref)
    # (3) training trains: each batch's loss has fallen, and by the same amount on both sides (within a tenth of the drop)
    for b in (0, 1):
        drop_ref = ref[b] - ref[STEPS - 2 + b]
        drop_ours = ours[b] - ours[STEPS - 2 + b]
        assert drop_ref > 0.1 * ref[b], ("the reference itself did not learn", ref)
        assert abs(drop_ours - drop_ref) <= 0.1 * drop_ref, (b, drop_ours, drop_ref)
    # (4) the trained students agree in aggregate: weights (all parameters, l2) and the activation observers' ranges
    w_ours = torch.cat([p.detach().flatten().cpu() for p in gpu_student.parameters()])
    w_ref = torch.cat([p.detach().flatten() for p in prepared.parameters()])
    assert rel_l2(w_ours, w_ref) < 1e-2, rel_l2(w_ours, w_ref)
    sd_ours, sd_ref = gpu_student.state_dict(), prepared.state_dict()
    assert list(sd_ours.keys()) == list(sd_ref.keys())
    rng_ours = torch.cat([sd_ours[k].flatten().cpu().float() for k in sd_ref if k.endswith("activation_post_process.scale")])
    rng_ref = torch.cat([sd_ref[k].flatten().float() for k in sd_ref if k.endswith("activation_post_process.scale")])
    assert rel_l2(rng_ours, rng_ref) < 1e-2, rel_l2(rng_ours, rng_ref)
    print(f"trajectory[{backend}]: max |dloss|/loss {abs(ours[worst] - ref[worst]) / abs(ref[worst]):.2e} (step {worst}), "
          f"step 0 {abs(ours[0] - ref[0]) / abs(ref[0]):.2e}, loss {ref[0]:.4f} -> {ref[-2]:.4f} (cpu) / {ours[0]:.4f} -> {ours[-2]:.4f} (gpu), "
          f"weights rel l2 {rel_l2(w_ours, w_ref):.2e}, activation scales rel l2 {rel_l2(rng_ours, rng_ref):.2e}")
