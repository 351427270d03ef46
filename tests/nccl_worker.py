"""Worker of tests/test_gpu_multi.py (launched with torch.distributed.run, one process per GPU, NCCL).

Checks, for a full batch sharded over the ranks (SURVEY.md §8e caveat "verify with a cross-rank checksum"):
  * ranks that start from DIFFERENT weights are made equal by DataParallelWithCallback (rank 0's state);
  * the rank-sharded train-mode generator output equals the full-batch oracle's rows for that shard (SyncBN statistics
    really are global), running statistics and spectral-norm u/v equal the full-batch oracle's and are BIT-IDENTICAL
    across ranks (the peer exchange sums in rank order on every rank);
  * after backward() of the reference-style generator / discriminator losses, every rank holds the SAME averaged
    gradient (bitwise), equal to the full-batch oracle gradient (mean over ranks of per-rank mean losses == full-batch
    mean for equal shards);
  * post optimizer-step weights are bit-identical across ranks.
Prints one JSON line on rank 0.
"""
import json
import os
import random
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.dirname(HERE), HERE, os.path.join(os.path.dirname(HERE), "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _bits_equal_across_ranks(t):
    """max over ranks of |t - t_rank0| == 0 and identical byte checksum."""
    ref = t.detach().clone()
    dist.broadcast(ref, src=0)
    same = torch.equal(ref, t.detach())
    flag = torch.tensor([1 if same else 0], device=t.device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return bool(flag.item())


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    import michigan_oracle as orc
    from helpers import preprocessed, reference_layout_state
    from michigan_b200.networks import sync_batchnorm as sbn
    from michigan_b200.options import make_opt
    from michigan_b200.pix2pix_model import Pix2PixModel
    from michigan_b200.synth import synthetic_batch

    size, per = 128, 2
    B = per * world
    cfg = dict(ngf=64, ndf=64, size=size, batch=B, data_seed=17)
    sdG = reference_layout_state("G", cfg, 41)
    sdD = reference_layout_state("D", cfg, 42)
    opt = make_opt(is_train=True, ngf=64, ndf=64, crop_size=size, batchSize=per)
    model = Pix2PixModel(opt)
    model.netG.load_state_dict(sdG)
    model.netD.load_state_dict(sdD)
    if rank != 0:                                   # the reference's train.py sets no seed: ranks would differ
        with torch.no_grad():
            for p_ in model.parameters():
                p_.add_(0.01 * (rank + 1))
            for b_ in model.buffers():
                if b_.dtype.is_floating_point:
                    b_.add_(0.5)
    wrap = sbn.DataParallelWithCallback(model, device_ids=[local])
    res = {"world": world}
    res["bcast"] = all(_bits_equal_across_ranks(t) for t in list(model.parameters()) + list(model.buffers()))
    model.train()
    optG, optD = model.create_optimizers(opt)

    data = synthetic_batch(B, size, 17)
    g = torch.Generator().manual_seed(3)
    data["image_ref"] = torch.rand(B, 3, size, size, generator=g) * 2 - 1        # samples differ => shards differ
    data["image_tag"] = data["image_ref"].clone()
    shard = {k: (v[rank * per:(rank + 1) * per] if torch.is_tensor(v) else v[rank * per:(rank + 1) * per]) for k, v in data.items()}

    th = int(size * 0.05); th = th if th % 2 == 1 else th + 1
    random.seed(5)
    k = random.choice([max(th - 4, 1), max(th - 2, 1), th, th + 2, th + 4])

    # ---------------- oracle: FULL batch on the CPU (every rank computes it; seconds at 128x128)
    oopt = orc.default_opt(ngf=64, ndf=64, crop_size=size, isTrain=True)
    pre = dict(input_ref=orc.one_hot(data["label_ref"]), input_tag=orc.one_hot(data["label_tag"]), image_ref=data["image_ref"],
               image_tag=data["image_tag"], orient_mask=data["orient"], noise=data["noise"])
    osdG = {k_: v.clone() for k_, v in sdG.items()}
    osdD = {k_: v.clone() for k_, v in sdD.items()}
    namesG = [n for n, _ in model.netG.named_parameters()]
    namesD = [n for n, _ in model.netD.named_parameters()]
    for n in namesG:
        osdG[n].requires_grad_(True)
    losses_o, fake_o = orc.compute_generator_loss(osdG, osdD, oopt, pre, rng_k=k)
    orc.trainer_loss(losses_o).backward()

    # ---------------- generator step on the shard
    random.seed(5)
    optG.zero_grad()
    g_losses, fake = wrap(dict(shard), mode="generator")
    sum(g_losses.values()).mean().backward()
    torch.cuda.synchronize()
    err = (fake.detach().cpu() - fake_o.detach()[rank * per:(rank + 1) * per]).abs().max().item()
    e = torch.tensor([err], device="cuda"); dist.all_reduce(e, op=dist.ReduceOp.MAX)
    res["g_out_max_abs_vs_full_batch_oracle"] = float(e.item())
    got = model.netG.state_dict()
    stat_err = 0.0
    for n, v in osdG.items():
        if n.endswith(("running_mean", "running_var", "weight_u", "weight_v")):
            stat_err = max(stat_err, (got[n].cpu() - v.detach()).abs().max().item() / max(1.0, v.abs().max().item()))
    res["running_stats_uv_rel_err_vs_oracle"] = stat_err
    res["buffers_bit_identical"] = all(_bits_equal_across_ranks(b) for b in model.netG.buffers())
    res["buffers_differing"] = [n for n, b in model.netG.named_buffers() if not _bits_equal_across_ranks(b)][:6]
    named = dict(model.netG.named_parameters())
    res["grads_bit_identical"] = all(_bits_equal_across_ranks(p_.grad) for p_ in model.netG.parameters() if p_.grad is not None)
    worst_cos, worst_rel = 1.0, 0.0
    gmax = max(osdG[n].grad.norm().item() for n in namesG if osdG[n].grad is not None)
    for n in namesG:
        r = osdG[n].grad
        if r is None:
            continue
        gq = named[n].grad
        if gq is None:
            continue
        a, b = gq.detach().cpu().double().flatten(), r.double().flatten()
        if b.norm() < 1e-3 * gmax:
            continue
        worst_cos = min(worst_cos, float(a @ b / (a.norm() * b.norm() + 1e-30)))
        worst_rel = max(worst_rel, float((a - b).norm() / b.norm()))
    res["g_grad_worst_cosine_vs_oracle"] = worst_cos
    res["g_grad_worst_rel_l2_vs_oracle"] = worst_rel
    res["g_losses"] = {k_: float(v.mean()) for k_, v in g_losses.items()}
    res["g_losses_oracle"] = {k_: float(v) for k_, v in losses_o.items()}
    optG.step()
    res["post_step_G_bit_identical"] = all(_bits_equal_across_ranks(p_) for p_ in model.netG.parameters())

    # ---------------- discriminator step
    random.seed(6)
    optD.zero_grad()
    d_losses = wrap(dict(shard), mode="discriminator")
    sum(d_losses.values()).mean().backward()
    res["d_grads_bit_identical"] = all(_bits_equal_across_ranks(p_.grad) for p_ in model.netD.parameters())
    optD.step()
    res["post_step_D_bit_identical"] = all(_bits_equal_across_ranks(p_) for p_ in model.netD.parameters())
    res["exchange_backend"] = sbn.exchange_backend()
    ex = sbn._exchange["state"]
    if ex:
        ex.check()
        res["peer_exchanges"] = ex.seq
    torch.cuda.synchronize()
    if rank == 0:
        print("NCCL_PARITY " + json.dumps(res))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
