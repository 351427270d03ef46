"""The N>1 host logic on CPU with gloo (world_size 2): SyncBN statistic exchange and the bucketed
gradient all-reduce behind DataParallelWithCallback."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from michigan_b200.networks import sync_batchnorm as sbn
    torch.manual_seed(0)
    x_full = torch.randn(4, 6, 5, 5, dtype=torch.float64)
    x = x_full[rank * 2:(rank + 1) * 2]
    sums = torch.cat([x.sum(dim=(0, 2, 3)), (x * x).sum(dim=(0, 2, 3))])
    sbn.allreduce_sums(sums)
    ref = torch.cat([x_full.sum(dim=(0, 2, 3)), (x_full * x_full).sum(dim=(0, 2, 3))])
    ok_stats = torch.allclose(sums, ref)
    # bucketed gradient averaging: reference semantics = mean over replicas of per-replica mean losses
    lin = torch.nn.Linear(7, 5)
    with torch.no_grad():
        for p in lin.parameters():
            p.fill_(0.5)
    wrap = sbn.DataParallelWithCallback(lin, device_ids=[0], bucket_mb=1e-5)   # tiny buckets -> several flushes
    opt = torch.optim.SGD(lin.parameters(), lr=1.0)
    wrap.attach_optimizer(opt)
    inp = torch.full((3, 7), float(rank + 1))
    wrap(inp).sum().backward()
    opt.step()
    expect_grad_w = torch.full((5, 7), 3.0 * (1 + 2) / 2)
    ok_grad = torch.allclose(lin.weight.grad, expect_grad_w) and torch.allclose(lin.bias.grad, torch.full((5,), 3.0))
    gathered = [torch.zeros_like(lin.weight) for _ in range(world)]
    dist.all_gather(gathered, lin.weight.detach())
    ok_same = all(torch.equal(g, gathered[0]) for g in gathered)
    q.put((rank, ok_stats, ok_grad, ok_same))
    dist.destroy_process_group()


def test_syncbn_sums_and_gradient_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(30)
    assert all(ok_s and ok_g and ok_w for _, ok_s, ok_g, ok_w in res), res
