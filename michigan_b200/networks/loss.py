"""Hinge GAN loss and GAN feature-matching loss with the reference's call signatures
(models/networks/loss.py:19-140 `GANLoss`, 144-175 `GANFeatLoss`; in-scope modes: gan_mode='hinge',
remove_background=False), evaluated by the fused reduction kernels of csrc/mg_loss.cu:

  * every call = ONE forward launch (all discriminator scales / all eight feature terms through a descriptor table,
    fp64 accumulation) and ONE backward launch, instead of ~10 eager ops per scale and per term;
  * the wide-edge weight map (loss.py:60-78) is one kernel per scale, cached per (label, size) within an iteration.
"""
import ctypes as C

import torch
import torch.nn as nn

from .. import _lib, ops


class _Term(C.Structure):
    _fields_ = [("a", C.c_void_p), ("b", C.c_void_p), ("ga", C.c_void_p), ("n", C.c_longlong), ("scale", C.c_float),
                ("sign", C.c_float), ("op", C.c_int32), ("out_slot", C.c_int32)]


OP_HINGE_D, OP_SUM, OP_L1 = 0, 1, 2
_table_cache = {}


def _dense(t):
    """A tensor whose elements occupy one contiguous block (any permutation of strides) or a contiguous copy."""
    if t.is_contiguous() or (t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last)):
        return t
    return t.contiguous()


def _device_table(terms, device):
    """ctypes term records -> device buffer (cached: the caching allocator hands the same addresses back every iteration)."""
    raw = bytes((_Term * len(terms))(*terms))
    hit = _table_cache.get(raw)
    if hit is not None and hit.device == device:
        return hit
    if len(_table_cache) > 256:
        _table_cache.clear()
    buf = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(device)
    _table_cache[raw] = buf
    return buf


class _LossFn(torch.autograd.Function):
    """spec: list of (op, a_index, b_tensor_or_None, scale, sign, slot); tensors: the differentiable `a` inputs."""

    @staticmethod
    def forward(ctx, spec, nslots, *tensors):
        assert _lib.load().mg_loss_term_bytes() == C.sizeof(_Term)
        dense = [_dense(t.detach()) for t in tensors]
        dev = dense[0].device
        terms = []
        keep = []
        for op, ai, b, scale, sign, slot in spec:
            a = dense[ai]
            bb = _dense(b.detach()) if b is not None else None
            if bb is not None:
                assert bb.numel() == a.numel(), (bb.shape, a.shape)
                keep.append(bb)
            terms.append(_Term(a.data_ptr(), bb.data_ptr() if bb is not None else None, None, a.numel(), float(scale), float(sign), op, slot))
        slots = torch.zeros(nslots, device=dev, dtype=torch.float64)
        table = _device_table(terms, dev)
        _lib.check(_lib.load().mg_loss_reduce(table.data_ptr(), len(terms), slots.data_ptr(), ops._stream()), "mg_loss_reduce")
        ctx.spec, ctx.dense, ctx.keep, ctx.nslots = spec, dense, keep, nslots
        ctx.shapes = [(t.shape, t.stride(), d.stride()) for t, d in zip(tensors, dense)]
        return slots.float()

    @staticmethod
    def backward(ctx, gslots):
        dense = ctx.dense
        dev = dense[0].device
        # one flat buffer for every gradient; a tensor that appears in several terms (never the case for these losses)
        # would need accumulation - guarded below
        offs, total = [], 0
        for d in dense:
            offs.append(total)
            total += d.numel()
        flat = torch.empty(total, device=dev, dtype=torch.float32)
        seen = set()
        terms = []
        for (op, ai, b, scale, sign, slot), kb in zip(ctx.spec, _iter_keep(ctx.spec, ctx.keep)):
            if ai in seen:
                raise RuntimeError("a tensor may appear in one loss term only")
            seen.add(ai)
            a = dense[ai]
            terms.append(_Term(a.data_ptr(), kb.data_ptr() if kb is not None else None, flat.data_ptr() + 4 * offs[ai], a.numel(),
                               float(scale), float(sign), op, slot))
        table = _device_table(terms, dev)
        g = gslots.contiguous().float()
        _lib.check(_lib.load().mg_loss_reduce_bwd(table.data_ptr(), len(terms), g.data_ptr(), ops._stream()), "mg_loss_reduce_bwd")
        grads = []
        for i, (shape, stride, dstride) in enumerate(ctx.shapes):
            if i not in seen:
                grads.append(None)
                continue
            grads.append(torch.as_strided(flat, shape, dstride, offs[i]))
        return (None, None) + tuple(grads)


def _iter_keep(spec, keep):
    it = iter(keep)
    for op, ai, b, scale, sign, slot in spec:
        yield next(it) if b is not None else None


class GANLoss(nn.Module):
    def __init__(self, gan_mode, target_real_label=1.0, target_fake_label=0.0, tensor=torch.FloatTensor, opt=None):
        super().__init__()
        if gan_mode != "hinge":
            raise NotImplementedError("michigan_b200: gan_mode '%s' (only 'hinge' is on the hot path)" % gan_mode)
        if getattr(opt, "remove_background", False):
            raise NotImplementedError("michigan_b200: --remove_background is not implemented")
        self.gan_mode = gan_mode
        self.opt = opt
        self._wcache = {}

    def get_weight_mask(self, input, mask):
        """loss.py:68-78: edges*wide_edge + (1-edges) at the logits' resolution, [N,h,w] (cached per label / size)."""
        n, _, h, w = input.shape
        m = mask.detach()
        key = (m.data_ptr(), m._version, tuple(m.shape), h, w)
        hit = self._wcache.get(key)
        if hit is not None:
            return hit
        m3 = m.reshape(m.shape[0], m.shape[-2], m.shape[-1]).contiguous().float()
        out = torch.empty((n, h, w), device=input.device, dtype=torch.float32)
        _lib.check(_lib.load().mg_edge_weight(m3.data_ptr(), out.data_ptr(), n, m3.shape[1], m3.shape[2], h, w,
                                              float(self.opt.wide_edge), ops._stream()), "mg_edge_weight")
        if len(self._wcache) > 8:
            self._wcache.clear()
        self._wcache[key] = out
        return out

    def __call__(self, input, target_is_real, for_discriminator=True, label=None):
        """loss.py:126-140: list over discriminators (each possibly a list of features, the logits last) -> mean over
        discriminators of the per-scale loss, a [1] tensor."""
        preds = input if isinstance(input, list) else [input]
        logits = [(p[-1] if isinstance(p, list) else p) for p in preds]
        num = len(logits)
        spec = []
        for i, x in enumerate(logits):
            if for_discriminator:
                wmap = self.get_weight_mask(x, label) if self.opt.wide_edge > 1.0 else None
                spec.append((OP_HINGE_D, i, wmap, -1.0 / (x.numel() * num), 1.0 if target_is_real else -1.0, 0))
            else:
                assert target_is_real, "The generator's hinge loss must be aiming for real"
                spec.append((OP_SUM, i, None, -1.0 / (x.numel() * num), 1.0, 0))
        return _LossFn.apply(spec, 1, *logits)


class GANFeatLoss(nn.Module):
    def __init__(self, opt=None):
        super().__init__()
        self.opt = opt

    def forward(self, pred_fake, pred_real, label=None):
        """loss.py:163-175: sum over discriminators and their intermediate features of L1(fake, real.detach()) * lambda_feat / num_D."""
        num_D = len(pred_fake)
        spec, tensors = [], []
        for i in range(num_D):
            for j in range(len(pred_fake[i]) - 1):
                f, r = pred_fake[i][j], pred_real[i][j].detach()
                fd = _dense(f)
                rd = _dense(r)
                if fd.stride() != rd.stride():
                    rd = r.contiguous()
                    fd = f.contiguous()
                spec.append((OP_L1, len(tensors), rd, self.opt.lambda_feat / (num_D * f.numel()), 1.0, 0))
                tensors.append(fd)
        return _LossFn.apply(spec, 1, *tensors)
