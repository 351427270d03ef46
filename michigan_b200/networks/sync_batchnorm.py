"""Replacement for the reference's vendored `models/networks/sync_batchnorm/` package.

The reference's data-parallel runtime is single-process torch.nn.DataParallel plus a Python
thread/queue master-slave SyncBN (batchnorm.py:51-145, comm.py:18-137, replicate.py:27-94).  The
B200 design is one process per GPU: statistics are all-reduced over NCCL (2*C doubles per BN) and
gradients are all-reduced in buckets, so the master/slave machinery disappears.  The public names
are kept so that `from models.networks.sync_batchnorm import SynchronizedBatchNorm2d,
DataParallelWithCallback` (normalization.py:10, pix2pix_trainer.py:6) keeps working.
"""
import torch
import torch.distributed as dist
import torch.nn as nn

from .. import ops


def _world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def allreduce_sums(sums):
    """Sum the per-rank [sum | sum of squares] vector across the data-parallel group (in place).
    Replaces SyncMaster.run_master / SlavePipe.run_slave + ReduceAddCoalesced/Broadcast
    (batchnorm.py:105-126, comm.py:49-133)."""
    if _world() > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    return sums


class SynchronizedBatchNorm2d(nn.Module):
    """Parameter/buffer container with the reference's state-dict layout (running_mean, running_var,
    num_batches_tracked, optional affine weight/bias) and its statistics semantics
    (batchnorm.py:63-93,128-145): train -> global batch statistics over all ranks, biased variance
    for normalisation, running stats updated with momentum 0.1 and the unbiased variance,
    num_batches_tracked never incremented (the reference's forward bypasses nn.BatchNorm.forward);
    eval -> running statistics.  1/sqrt(var+eps) is used in every mode (the reference's
    single-replica / CPU path, batchnorm.py:65-68); see DESIGN.md for the clamp variant.
    """

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True):
        super().__init__()
        self.num_features, self.eps, self.momentum, self.affine = num_features, eps, momentum, affine
        if affine:
            self.weight = nn.Parameter(torch.ones(num_features))
            self.bias = nn.Parameter(torch.zeros(num_features))
        else:
            self.register_parameter("weight", None)
            self.register_parameter("bias", None)
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))

    def scale_shift(self, x_nhwc, upsample_shift=0, extra_running=()):
        """(nscale, nshift) such that x_hat = x*nscale + nshift, for an NHWC tensor.  `upsample_shift`:
        the reference normalises the 2^shift nearest-upsampled tensor - same mean/var, 4^shift more
        samples in the unbiased factor.  `extra_running`: other BN containers that see the same
        input (norm_s next to norm_0) and whose running buffers get the identical update."""
        if not self.training:
            return ops.bn_from_running(self.running_mean, self.running_var, self.eps)
        sums = allreduce_sums(ops.bn_sums(x_nhwc))
        count = (x_nhwc.numel() // x_nhwc.shape[-1]) * _world()
        cu = count * (4 ** upsample_shift)
        if cu <= 1:
            raise ValueError("BatchNorm computes unbiased standard-deviation, which requires size > 1.")
        out = ops.bn_finalize(sums, count, cu, self.eps, self.momentum, 0, self.running_mean, self.running_var,
                              want_stats=True)
        for other in extra_running:
            ops.bn_finalize(sums, count, cu, other.eps, other.momentum, 0, other.running_mean, other.running_var)
        return out

    def forward(self, x):
        raise RuntimeError("SynchronizedBatchNorm2d is consumed by the fused SPADE kernels; it has no standalone forward")

    def extra_repr(self):
        return "{num_features}, eps={eps}, momentum={momentum}, affine={affine}".format(**self.__dict__)


SynchronizedBatchNorm1d = SynchronizedBatchNorm2d
SynchronizedBatchNorm3d = SynchronizedBatchNorm2d


class DataParallelWithCallback(nn.Module):
    """Drop-in for replicate.py:50-67 under one-process-per-GPU.  `device_ids` is accepted for
    signature compatibility; the wrapped module lives on this process's device.  Calls are forwarded
    unchanged (`wrapper(data, mode=...)`), the input is this rank's shard, and gradients are averaged
    across ranks by `sync_gradients()` (called from the optimizer pre-step hook installed by
    `attach_optimizer`), which replaces DataParallel's reduce-add to GPU 0."""

    def __init__(self, module, device_ids=None, output_device=None, dim=0, bucket_mb=48):
        super().__init__()
        self.module = module
        self.device_ids = list(device_ids) if device_ids is not None else []
        self.bucket_bytes = int(bucket_mb * 1024 * 1024)

    def forward(self, *inputs, **kwargs):
        return self.module(*inputs, **kwargs)

    def sync_gradients(self, params):
        """Bucketed NCCL all-reduce (mean) of the gradients of `params`."""
        world = _world()
        if world == 1:
            return
        bucket, size = [], 0
        handles = []

        def flush():
            nonlocal bucket, size
            if not bucket:
                return
            flat = torch.cat([g.reshape(-1) for g in bucket])
            h = dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True)
            handles.append((h, flat, bucket))
            bucket, size = [], 0

        for p in params:
            if p.grad is None:
                continue
            bucket.append(p.grad)
            size += p.grad.numel() * p.grad.element_size()
            if size >= self.bucket_bytes:
                flush()
        flush()
        for h, flat, grads in handles:
            h.wait()
            flat.div_(world)
            off = 0
            for g in grads:
                n = g.numel()
                g.copy_(flat[off:off + n].view_as(g))
                off += n

    def attach_optimizer(self, optimizer):
        """Average gradients across ranks right before `optimizer.step()` (the unchanged trainer calls
        backward() then step(), pix2pix_trainer.py:42-58,66-69)."""
        params = [p for grp in optimizer.param_groups for p in grp["params"]]

        def pre_step(opt, args, kwargs):
            self.sync_gradients(params)

        optimizer.register_step_pre_hook(pre_step)
        return optimizer


def patch_replication_callback(data_parallel):
    return data_parallel


def convert_model(module):
    return module


def patch_sync_batchnorm():
    import contextlib
    return contextlib.nullcontext()
