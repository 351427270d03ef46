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


# ================================================================================================ Gabor orientation loss
def gabor_bank(device, num_kernels=32, kernel_size=17):
    """The 32 Gabor kernels the reference rebuilds on every call (gabor_fn, loss.py:214-240; sigma_x 2, sigma_y 3, lambda 4,
    psi 0, theta_k = pi*k/32), with the same float32 torch arithmetic, laid out [17*17][32] for the fused kernel."""
    import math
    r = kernel_size // 2
    y = torch.arange(-r, r + 1, device=device).view(1, -1).repeat(kernel_size, 1).float()       # varies along columns
    x = torch.arange(-r, r + 1, device=device).view(-1, 1).repeat(1, kernel_size).float()       # varies along rows
    ks = []
    for k in range(num_kernels):
        theta = torch.ones(1, device=device) * (math.pi * k / num_kernels)
        x_t = x * torch.cos(theta) + y * torch.sin(theta)
        y_t = -x * torch.sin(theta) + y * torch.cos(theta)
        ks.append(torch.exp(-.5 * (x_t ** 2 / 2.0 ** 2 + y_t ** 2 / 3.0 ** 2)) * torch.cos(2 * math.pi / 4.0 * x_t + 0.0))
    return torch.stack(ks, dim=-1).reshape(kernel_size * kernel_size, num_kernels).contiguous()


class _OrientLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, bank, label2, hair):
        n, _, h, w = img.shape
        dev = img.device
        img_c = img.detach().contiguous()
        idx = torch.empty((n, h, w), device=dev, dtype=torch.uint8)
        d1 = torch.empty((n, h, w), device=dev, dtype=torch.float32)
        d2 = torch.empty((n, h, w), device=dev, dtype=torch.float32)
        sums = torch.zeros(3, device=dev, dtype=torch.float64)
        _lib.check(_lib.load().mg_orient_loss_fwd(img_c.data_ptr(), bank.data_ptr(), label2.data_ptr(), hair.data_ptr(), idx.data_ptr(),
                                                  d1.data_ptr(), d2.data_ptr(), sums.data_ptr(), n, h, w, ops._stream()), "mg_orient_loss_fwd")
        ctx.save_for_backward(bank, idx, d1, d2, sums)
        ctx.shape = (n, h, w)
        orient = (sums[0] / (2.0 * n * h * w)).float().reshape(())
        conf = (-sums[1] / sums[2]).float().reshape(())
        return orient, conf

    @staticmethod
    def backward(ctx, g_orient, g_conf):
        bank, idx, d1, d2, sums = ctx.saved_tensors
        n, h, w = ctx.shape
        wts = torch.stack([g_orient.double() / (2.0 * n * h * w), -g_conf.double() / sums[2]]).float().contiguous()
        dimg = torch.empty((n, 3, h, w), device=idx.device, dtype=torch.float32)
        _lib.check(_lib.load().mg_orient_loss_bwd(bank.data_ptr(), idx.data_ptr(), d1.data_ptr(), d2.data_ptr(), wts.data_ptr(), dimg.data_ptr(),
                                                  n, h, w, ops._stream()), "mg_orient_loss_bwd")
        return dimg, None, None, None


class L1OLoss(nn.Module):
    """loss.py:274-385 with orient_filter='gabor': (orient_loss, confidence_loss) = forward(fake_image, orientation_label,
    input_semantics).  orientation_label: the 1-channel angle map (0..255) or, with --use_ig, the 2-channel (sin 2t, cos 2t) map."""

    def __init__(self, opt, channel_in=1, channel_out=1, stride=1, padding=8):
        super().__init__()
        if "gabor" not in getattr(opt, "orient_filter", "gabor"):
            raise NotImplementedError("michigan_b200: orient_filter '%s' (the DoG variant has no kernel)" % opt.orient_filter)
        self.opt = opt
        self._bank = None

    def forward(self, fake_image0, orientation_label0, input_semantics):
        import math
        dev = fake_image0.device
        if self._bank is None or self._bank.device != dev:
            self._bank = gabor_bank(dev)
        hair = input_semantics[:, 1].contiguous().float()
        if not self.opt.use_ig:
            t = orientation_label0.float() / 255 * math.pi
            label2 = torch.cat([torch.sin(2 * t), torch.cos(2 * t)], dim=1).contiguous()
        else:
            label2 = orientation_label0.float().contiguous()
        return _OrientLossFn.apply(fake_image0, self._bank, label2.detach(), hair.detach())
