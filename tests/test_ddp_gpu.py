"""2-GPU NCCL check of the data-parallel step (BASELINE.json configs[2]): after the one exchange
(qatvit_b200.ddp.GradSync over NCCL) every rank holds (a) the SUM over ranks of the per-rank gradients that single-GPU
engines produce on the same shards from the same incoming state, and (b) rank 0's activation-observer running min/max
(DDP's broadcast_buffers rule, torch/nn/parallel/distributed.py:1590-1591).  Skipped with fewer than 2 GPUs."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import copy
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import qatvit_b200  # noqa: F401
    from qatvit_b200.ddp import GradSync
    from qatvit_b200.engine import QATDistillStep
    from parity_utils import build_models
    vr, prepared, teacher = build_models("fbgemm", "vit_test_tiny", "vit_test_teacher", 64)
    B = 4
    images, labels = vr.synthetic_batch(B * world, seed=21, img=64)
    hp = dict(vr.DEFAULT_HPARAMS)
    student = copy.deepcopy(prepared).to(dev)
    n_grad = QATDistillStep.count_trainable(student)
    n_obs = QATDistillStep.count_activation_observers(student)
    sync = GradSync(n_grad, n_obs, dev)
    step = QATDistillStep(student, copy.deepcopy(teacher).to(dev), B, hp, grad_buffer=sync.grad_arena)
    sync.bind_observers(step.activation_observers())
    sl = slice(rank * B, (rank + 1) * B)
    step(images[sl].to(dev), labels[sl].to(dev))
    local = sync.grad_arena.clone()
    local_obs = [(float(a), float(b)) for a, b in step.activation_observers()]
    sync.all_reduce()
    torch.cuda.synchronize()
    reduced_blocking = sync.grad_arena.clone()
    # the overlapped mode (all-reduce issued layer by layer during backward) must give the same sums: rerun the same step on a
    # fresh copy of the model / observer state
    student2 = copy.deepcopy(prepared).to(dev)
    sync2 = GradSync(n_grad, n_obs, dev)
    step2 = QATDistillStep(student2, copy.deepcopy(teacher).to(dev), B, hp, grad_buffer=sync2.grad_arena)
    sync2.bind_observers(step2.activation_observers())
    step2(images[sl].to(dev), labels[sl].to(dev), grad_sync=sync2)
    torch.cuda.synchronize()
    assert torch.equal(sync2.grad_arena, reduced_blocking), "overlapped all-reduce differs from the blocking one"
    assert [(float(a), float(b)) for a, b in step2.activation_observers()] == [(float(a), float(b)) for a, b in step.activation_observers()]
    q.put((rank, local.cpu(), sync.grad_arena.cpu().clone(), local_obs, [(float(a), float(b)) for a, b in step.activation_observers()]))
    dist.barrier()
    dist.destroy_process_group()


def test_nccl_gradient_sum_and_rank0_observers():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    world, port = 2, 29500 + os.getpid() % 400
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    total = res[0][1] + res[1][1]
    assert float(total.abs().max()) > 0
    for rank, _, reduced, _, obs_after in res:
        assert torch.equal(reduced, total) or torch.allclose(reduced, total, rtol=1e-6, atol=1e-9)
        assert obs_after == res[0][3]                      # rank 0's state everywhere
    assert res[1][3] != res[0][3]                          # the shards really had different local observer state
