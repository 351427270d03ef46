"""Sub-networks instantiated inside SPADEBGenerator: the partial-conv reference encoder
(reference models/networks/encoder.py:160-225, partialconv2d.py:15-85) and the background encoder
(encoder.py:271-341, MaskGAN_networks.py:114-173).  Same constructors and state-dict keys; forward on
the CUDA kernels, NHWC inside."""
import random
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn as nn

from .. import ops, precision
from .base_network import BaseNetwork
from .prep import PackCache


class PartialConv2d(nn.Conv2d):
    """Parameter container with partialconv2d.py's constructor (multi_channel / return_mask kwargs)."""

    def __init__(self, *args, **kwargs):
        self.multi_channel = kwargs.pop("multi_channel", False)
        self.return_mask = kwargs.pop("return_mask", False)
        super().__init__(*args, **kwargs)
        if self.multi_channel:
            raise NotImplementedError("michigan_b200: multi-channel partial-conv masks are unused by spadeb")

    def forward(self, input, mask_in=None):
        raise RuntimeError("PartialConv2d is executed by ImageEncoder3's fused kernels")


class ImageEncoder3(BaseNetwork):
    """encoder.py:160-225 with norm_ref_encode='instance' (the default, base_options.py:94)."""

    def __init__(self, opt, sw, sh):
        super().__init__()
        kw = 3
        pw = int(np.ceil((kw - 1.0) / 2))
        ndf = opt.ngf
        self.sw, self.sh, self.opt = sw, sh, opt
        self.layer1 = PartialConv2d(3, ndf, kw, stride=2, padding=pw, return_mask=True)
        self.norm1 = nn.InstanceNorm2d(ndf, affine=False)
        self.layer2 = PartialConv2d(ndf * 1, ndf * 2, kw, stride=2, padding=pw, return_mask=True)
        self.norm2 = nn.InstanceNorm2d(ndf * 2, affine=False)
        self.layer3 = PartialConv2d(ndf * 2, ndf * 4, kw, stride=2, padding=pw, return_mask=True)
        self.norm3 = nn.InstanceNorm2d(ndf * 4, affine=False)
        self.layer4 = PartialConv2d(ndf * 4, ndf * 8, kw, stride=2, padding=pw, return_mask=True)
        self.norm4 = nn.InstanceNorm2d(ndf * 8, affine=False)
        self.layer5 = PartialConv2d(ndf * 8, ndf * 16, kw, stride=2, padding=pw, return_mask=True)
        self.norm5 = nn.InstanceNorm2d(ndf * 16, affine=False)
        self.actvn = nn.LeakyReLU(0.2, False)
        if "instance" not in opt.norm_ref_encode:
            raise NotImplementedError("michigan_b200: norm_ref_encode must be 'instance'")
        if ndf % 32 != 0:
            raise NotImplementedError("michigan_b200: ngf must be a multiple of 32")
        self._cache = PackCache()

    def forward_nhwc(self, image_ref, label_ref0, label_tag0, save=None):
        """image_ref [N,3,H,W] NCHW, label_* [N,1,H,W] -> [N,sh,sw,16*ngf] NHWC.
        save: namespace filled with what autograd.fc_bwd needs (same arithmetic either way)."""
        N, _, H, W = image_ref.shape
        mref = label_ref0.reshape(N, H, W).contiguous()
        mtag = label_tag0.reshape(N, H, W).contiguous()
        c = self._cache
        # layer 1: thin direct conv on image*mask
        x0 = ops.nchw_to_nhwc(image_ref.contiguous(), 4, pmul=mref)
        ratio, upd = ops.partial_mask(mref, 3, 2, 1)
        w1 = c.get("l1", [self.layer1.weight], lambda: ops.pack_weight_thin(self.layer1.weight.detach(), 4))
        x = ops.conv_thin(x0, w1, self.layer1.bias.detach(), self.layer1.out_channels, 3, 3, 2, 1, pscale=ratio, pmul=upd)
        if save is not None:
            save.mref, save.mtag, save.x0, save.wt1, save.layers = mref, mtag, x0, w1, []
            save.l1 = SimpleNamespace(ratio=ratio, upd=upd, y=x)
        for i in range(2, 6):
            # lrelu(IN(x)) * mask as a tensor-core operand of the next partial conv (partialconv2d.py:69)
            fmt = precision.conv_fmt(x.shape[-1])
            if save is None:
                kw_o, get_o = precision.out_spec(fmt, fmt == ops.BF16)
                xo = get_o(ops.instance_norm_act(x, ops.ACT_LRELU, 1e-5, pmul=upd, **kw_o))
            elif fmt == ops.TF32:
                a32, ss = ops.instance_norm_act_fwd(x, ops.ACT_LRELU, 1e-5, round_out=True, pmul=upd)
                xo = (ops.TF32, a32, None)
            else:
                a32, ss, hi, lo = ops.instance_norm_act_fwd(x, ops.ACT_LRELU, 1e-5, pmul=upd, out16=(fmt, True))
                xo = (fmt, hi, lo)
            ratio, upd_next = ops.partial_mask(upd, 3, 2, 1)
            layer = getattr(self, "layer%d" % i)
            wp = c.get(("l%d" % i, fmt), [layer.weight], lambda layer=layer: precision.pack_conv(layer.weight.detach(), None, fmt))
            y = precision.conv(xo, wp, layer.out_channels, 3, 3, 2, 1, bias=layer.bias.detach(), pscale=ratio, pmul=upd_next)
            if save is not None:
                save.layers.append(SimpleNamespace(layer=layer, a=a32, ss=ss, y_in=x, pm_in=upd, ratio=ratio, upd=upd_next))
            x, upd = y, upd_next
        if save is None:
            a6 = ops.instance_norm_act(x, ops.ACT_LRELU, 1e-5)
        else:
            a6, ss6 = ops.instance_norm_act_fwd(x, ops.ACT_LRELU, 1e-5)
            save.y5, save.ss6 = x, ss6
        out = ops.masked_mean_bcast(a6, mref, mtag)
        if save is not None:
            save.mhw = (out.shape[1], out.shape[2])
        if self.sh != out.shape[1]:
            out = ops.resize_bilinear(out, self.sh, self.sw)
        return out

    def forward(self, x, label_ref0, label_tag0):
        return self.forward_nhwc(x, label_ref0, label_tag0).permute(0, 3, 1, 2)


class ConvBlock(nn.Module):
    """MaskGAN_networks.py:114-173, the configuration BackgroundEncode2 uses (norm='none', reflect pad)."""

    def __init__(self, input_dim, output_dim, kernel_size, stride, padding=0, norm="none", activation="relu", pad_type="zero"):
        super().__init__()
        if norm != "none" or activation != "relu" or pad_type != "reflect":
            raise NotImplementedError("michigan_b200: ConvBlock supports norm='none', activation='relu', pad_type='reflect'")
        self.use_bias = True
        self.pad = nn.ReflectionPad2d(padding)
        self.norm = None
        self.activation = nn.ReLU(inplace=True)
        self.conv = nn.Conv2d(input_dim, output_dim, kernel_size, stride, bias=self.use_bias)
        self.padding, self.kernel_size, self.stride = padding, kernel_size, stride


class BackgroundEncode2(BaseNetwork):
    """encoder.py:271-341 (num_upsampling_layers != 'most')."""

    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        self.ngf = opt.ngf
        if opt.num_upsampling_layers == "most":
            raise NotImplementedError("michigan_b200: num_upsampling_layers='most' is not on the benchmark path")
        self.conv1 = ConvBlock(3, self.ngf, 7, 1, 3, norm="none", activation="relu", pad_type="reflect")
        self.layer1 = ConvBlock(self.ngf, 2 * self.ngf, 4, 2, 1, norm="none", activation="relu", pad_type="reflect")
        self.layer2 = ConvBlock(2 * self.ngf, 4 * self.ngf, 4, 2, 1, norm="none", activation="relu", pad_type="reflect")
        self.layer3 = ConvBlock(4 * self.ngf, 8 * self.ngf, 4, 2, 1, norm="none", activation="relu", pad_type="reflect")
        # present in the reference's state dict, never executed (encoder.py:284 vs 323-330)
        self.layer4 = ConvBlock(8 * self.ngf, 16 * self.ngf, 4, 2, 1, norm="none", activation="relu", pad_type="reflect")
        self._cache = PackCache()

    def back_mask(self, mask):
        """[N,H,W] complement of the dilated hair mask (encoder.py:286-316).  The train/eval switch is
        opt.isTrain, as in the reference; the dilation size is drawn with Python's `random` there too."""
        opt = self.opt
        hair = mask[:, 1].contiguous()
        if opt.isTrain:
            if opt.random_expand_mask:
                mh = hair.shape[1]
                th = int(mh * opt.random_expand_th)
                th = th if th % 2 == 1 else th + 1
                k = random.choice([max(th - 4, 1), max(th - 2, 1), th, th + 2, th + 4])
                return ops.maxpool_mask(hair, k, invert=True)
            return mask[:, 0].contiguous()
        if opt.expand_mask_be:
            k = opt.expand_th
            if opt.add_feat_zeros:
                th = opt.add_th
                H = W = opt.crop_size
                o = int(th / 2)
                inner = ops.maxpool_mask(hair[:, o:o + H, o:o + W].contiguous(), k, invert=False)
                e = torch.zeros_like(hair)
                e[:, o:o + H, o:o + W] = inner
                return (1 - e).contiguous()
            return ops.maxpool_mask(hair, k, invert=True)
        return mask[:, 0].contiguous()

    def forward_nhwc(self, image, mask, noise, save=None):
        """Returns ([x3,x2,x1,x0] NHWC features, back_mask [N,H,W]).  save: namespace for autograd.bg_bwd."""
        back = self.back_mask(mask)
        if self.opt.random_noise_background:
            inp = ops.nchw_to_nhwc(noise.contiguous(), 4)
        else:
            inp = ops.prep_bginput(image.contiguous(), noise.contiguous(), back)
        c = self._cache
        w1 = c.get("c1", [self.conv1.conv.weight], lambda: ops.pack_weight_thin(self.conv1.conv.weight.detach(), 4))
        x0 = ops.conv_thin(inp, w1, self.conv1.conv.bias.detach(), self.ngf, 7, 7, 1, 3, pad_mode=1, act=ops.ACT_RELU)
        if save is not None:
            save.inp, save.wt1, save.x0, save.layers = inp, w1, x0, []
        feats = [x0]
        x = x0
        for name in ("layer1", "layer2", "layer3"):
            blk = getattr(self, name)
            fmt = precision.conv_fmt(x.shape[-1])
            wp = c.get((name, fmt), [blk.conv.weight], lambda blk=blk: precision.pack_conv(blk.conv.weight.detach(), None, fmt))
            if fmt == ops.TF32:
                xp32 = ops.reflect_pad(x, 1, round_tf32=True)
                xp = (ops.TF32, xp32, None)
            else:
                xp32, hi, lo = ops.reflect_pad(x, 1, out16=(fmt, True), want_f32=save is not None)
                xp = (fmt, hi, lo)
            x = precision.conv(xp, wp, blk.conv.out_channels, 4, 4, 2, 0, bias=blk.conv.bias.detach(), act=ops.ACT_RELU)
            if save is not None:
                save.layers.append(SimpleNamespace(blk=blk, xp=xp32, y=x))
            feats.append(x)
        return feats[::-1], back

    def forward(self, image, mask, noise):
        feats, back = self.forward_nhwc(image, mask, noise)
        import torch.nn.functional as F
        bm = back.unsqueeze(1)
        sh, sw = bm.shape[2:]
        masks = [F.interpolate(bm, size=(int(sh / d), int(sw / d)), mode="nearest") for d in (8, 4, 2)] + [bm]
        return [f.permute(0, 3, 1, 2) for f in feats], masks
