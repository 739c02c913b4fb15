"""CPU, world_size 2, gloo: the one exchange step of the data-parallel path (qatvit_b200/ddp.py) -- gradient SUM over
ranks plus rank-0-authoritative observer state piggy-backed on the same buffer."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    """A TCP port nobody listens on right now (a pid-derived guess collided with other listeners on busy hosts)."""
    import socket
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, async_op, q, rehome=False):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import qatvit_b200  # noqa: F401
    from qatvit_b200.ddp import GradSync
    torch.manual_seed(100 + rank)
    obs = [(torch.tensor(float(-1 - rank - i)), torch.tensor(float(1 + rank + i))) for i in range(5)]
    gs = GradSync(1000, len(obs), device="cpu", bucket_bytes=1024)      # several buckets
    gs.bind_observers(obs, rehome=rehome)
    if rehome:       # the observer buffers now live in the tail of the exchange buffer: same objects, same values
        assert all(a.data_ptr() >= gs.tail.data_ptr() for a, _ in obs)
        assert [(float(a), float(b)) for a, b in obs] == [(float(-1 - rank - i), float(1 + rank + i)) for i in range(5)]
    g = torch.randn(1000)
    gs.grad_arena.copy_(g)
    if async_op == "overlapped":        # layer-by-layer suffixes, as the engine's backward reports them
        gs.begin_step(min_bucket_bytes=512)
        for lo in (900, 870, 600, 590, 300, 0):
            gs.grads_final_from(lo)
        gs.end_step()
    else:
        works = gs.all_reduce(async_op=async_op)
        if async_op:
            gs.finish(works)
    q.put((rank, g, gs.grad_arena.clone(), [(float(a), float(b)) for a, b in obs]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("async_op", [False, True, "overlapped"])
def test_gradient_sum_and_rank0_observer_state(async_op):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, async_op, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    total = res[0][1] + res[1][1]
    for rank, _, reduced, obs in res:
        assert torch.allclose(reduced, total, atol=1e-6)                  # SUM; the 1/world is applied in the clip pass
        assert obs == [(float(-1 - i), float(1 + i)) for i in range(5)]   # everyone ends with rank 0's running min/max


def test_rehomed_observers_need_no_pack_or_unpack():
    """bind_observers(rehome=True): the running min / max buffers ARE the tail of the exchange buffer -- rank != 0 clears them
    before the exchange, nothing is copied afterwards, and every rank still ends with rank 0's state."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, "overlapped", q, True)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    total = res[0][1] + res[1][1]
    for rank, _, reduced, obs in res:
        assert torch.allclose(reduced, total, atol=1e-6)
        assert obs == [(float(-1 - i), float(1 + i)) for i in range(5)]
