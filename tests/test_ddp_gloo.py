"""The N>1 host logic on CPU with gloo (world_size 2): SyncBN statistic exchange (sums + sample count), the rank-0
parameter/buffer broadcast of DataParallelWithCallback, and the staged gradient all-reduce that runs inside the
hand-written backward (GradReducer), first on a toy network with a known answer, then on the real generator /
discriminator glue with the no-op library of tests/dryrun.py (plumbing, no deadlock, `.grad` = views of the flat buffer)."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _ToyFn(torch.autograd.Function):
    """y = (x @ W1^T) @ W2^T with a hand-written backward that follows the protocol of networks/autograd.py."""

    @staticmethod
    def forward(ctx, net, x, *params):
        h = x @ net.l1.weight.t()
        ctx.net, ctx.x, ctx.h, ctx.params = net, x, h, params
        return h @ net.l2.weight.t()

    @staticmethod
    def backward(ctx, dy):
        net = ctx.net
        grads = {}
        red = getattr(net, "_grad_reducer", None)
        if red is not None:
            red.begin(ctx.params)
        grads[id(net.l2.weight)] = dy.t() @ ctx.h
        if red is not None:
            red.reduce_stage(0, lambda p: grads.get(id(p)))
        dh = dy @ net.l2.weight
        grads[id(net.l1.weight)] = dh.t() @ ctx.x
        if red is not None:
            red.reduce_stage(1, lambda p: grads.get(id(p)))
            return (None, None) + red.finish(ctx.params, lambda p: grads.get(id(p)))
        return (None, None) + tuple(grads.get(id(p)) for p in ctx.params)


class _Toy(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.l1 = torch.nn.Linear(7, 5, bias=False)
        self.l2 = torch.nn.Linear(5, 3, bias=False)
        self.unused = torch.nn.Linear(2, 2, bias=False)      # like backgroud_enc.layer4: in the state dict, never executed
        self.register_buffer("stat", torch.zeros(3))

    def grad_stages(self):
        return [[self.l2.weight], [self.l1.weight, self.unused.weight]]

    def forward(self, x):
        return _ToyFn.apply(self, x, *self.parameters())


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    for p in (os.path.dirname(HERE), HERE, os.path.join(os.path.dirname(HERE), "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from michigan_b200.networks import sync_batchnorm as sbn
    res = {}
    # ---- statistics: [sum | sum of squares | count], uneven shards (3 + 1 samples)
    torch.manual_seed(0)
    x_full = torch.randn(4, 6, 5, 5, dtype=torch.float64)
    x = x_full[:3] if rank == 0 else x_full[3:]
    sums = torch.cat([x.sum(dim=(0, 2, 3)), (x * x).sum(dim=(0, 2, 3)), torch.zeros(1, dtype=torch.float64)])
    cnt = sbn.allreduce_sums(sums, x.shape[0] * 25)
    ref = torch.cat([x_full.sum(dim=(0, 2, 3)), (x_full * x_full).sum(dim=(0, 2, 3))])
    res["stats"] = bool(torch.allclose(sums[:-1], ref)) and float(sums[-1]) == 100.0 and cnt == 0.0
    # ---- construction: ranks start from DIFFERENT weights (the reference's train.py sets no seed); rank 0's win
    torch.manual_seed(100 + rank)
    net = _Toy()
    net.stat.fill_(float(rank + 1))
    w0 = net.l1.weight.detach().clone()
    wrap = sbn.DataParallelWithCallback(net, device_ids=[0])
    gathered = [torch.zeros_like(w0) for _ in range(world)]
    dist.all_gather(gathered, net.l1.weight.detach().clone())
    res["bcast"] = all(torch.equal(g, gathered[0]) for g in gathered) and float(net.stat[0]) == 1.0 and \
        (rank != 0 or torch.equal(net.l1.weight.detach(), w0))
    # ---- staged gradient averaging == gradient of the mean over ranks of the per-rank losses
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    torch.manual_seed(7)
    xs = torch.randn(world, 4, 7)
    for it in range(2):
        opt.zero_grad()            # set_to_none=True (torch default): the views are re-installed by every backward
        wrap(xs[rank]).pow(2).sum().backward()
        if it == 0:
            W1, W2 = net.l1.weight.detach().clone().requires_grad_(), net.l2.weight.detach().clone().requires_grad_()
            sum(((xs[r] @ W1.t()) @ W2.t()).pow(2).sum() for r in range(world)).div(world).backward()
            res["grad"] = bool(torch.allclose(net.l1.weight.grad, W1.grad, atol=1e-5) and torch.allclose(net.l2.weight.grad, W2.grad, atol=1e-5))
            red = net._grad_reducer
            res["views"] = net.l1.weight.grad.data_ptr() == red.views[id(net.l1.weight)].data_ptr() and net.unused.weight.grad is None
        opt.step()
    # gradient accumulation (no zero_grad): the second backward must not corrupt the installed views
    opt.zero_grad()
    wrap(xs[rank]).pow(2).sum().backward()
    g1 = net.l1.weight.grad.clone()
    wrap(xs[rank]).pow(2).sum().backward()
    res["accum"] = bool(torch.allclose(net.l1.weight.grad, 2 * g1, atol=1e-5))
    dist.all_gather(gathered, net.l1.weight.detach().clone())
    res["same"] = all(torch.equal(g, gathered[0]) for g in gathered)
    # ---- the real networks' glue under world 2 (no-op library: shapes and control flow only)
    from dryrun import dry_run
    from michigan_b200 import networks
    from michigan_b200.options import make_opt
    from helpers import preprocessed
    torch.manual_seed(200 + rank)
    opt_ = make_opt(is_train=True, ngf=64, ndf=64, crop_size=64, gpu_ids=[])
    G = networks.SPADEBGenerator(opt_).train()
    D = networks.MultiscaleDiscriminator(opt_).train()
    G.init_weights("xavier", 0.02)
    both = torch.nn.ModuleList([G, D])
    sbn.DataParallelWithCallback(both, device_ids=[0])
    u = G.head_0.conv_0.weight_u.detach().clone()
    gu = [torch.zeros_like(u) for _ in range(world)]
    dist.all_gather(gu, u)
    res["bcast_G"] = all(torch.equal(g, gu[0]) for g in gu)
    _, pre = preprocessed(dict(batch=1, size=64, data_seed=rank))
    with dry_run():
        fake = G(pre["input_ref"], orient_mask=pre["orient_mask"], image_ref=pre["image_ref"], input_tag=pre["input_tag"],
                 noise=pre["noise"], image_tag=pre["image_tag"])
        xin = torch.cat([torch.cat([torch.zeros(1, 4, 64, 64), fake], 1), torch.cat([torch.zeros(1, 4, 64, 64), pre["image_tag"]], 1)], 0)
        sum(t.mean() for o in D(xin) for t in o).backward()
    rg, rd = G._grad_reducer, D._grad_reducer
    res["real_views"] = all(p.grad is not None and p.grad.data_ptr() == rg.views[id(p)].data_ptr()
                            for n, p in G.named_parameters() if not n.startswith("backgroud_enc.layer4")) and \
        all(p.grad.data_ptr() == rd.views[id(p)].data_ptr() for p in D.parameters()) and \
        G.backgroud_enc.layer4.conv.weight.grad is None
    q.put((rank, res))
    dist.destroy_process_group()


def test_dataparallel_runtime_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(30)
    for rank, r in res:
        assert all(r.values()), (rank, r)
