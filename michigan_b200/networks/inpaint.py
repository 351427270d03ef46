"""InpaintGenerator — the frozen orientation-inpainting sub-net of `--use_ig` (the first "next"
row; reference: models/networks/generator.py:450-575, driven by Pix2PixModel.inpainting_orient,
pix2pix_model.py:407-429).

STATUS: GPU parity vs the pinned CPU oracle is green (tests/test_gpu_parity.py::test_inpaint_generator_vs_oracle,
max-abs <= 1e-3 on the [0,1] output); the index arithmetic it relies on (dilated conv as a dense conv over the four
parity sub-grids, ConvTranspose2d as the data gradient of a strided conv) is also verified on the CPU in
tests/test_host_logic.py.

Module structure = the reference's (same `nn.Sequential` indices), so the state-dict keys of
`InpaintingModel_gen.pth['generator']` load unchanged (util.py:245-257).  The network only ever runs in eval mode:
spectral norm uses the stored u, v (no power iteration), InstanceNorm has no running statistics.

Mapping onto the C ABI:
  * 7x7 reflect-padded 4 -> 64 stem ............ mg_conv_thin (pad_mode 1), as BackgroundEncode2.conv1
  * 4x4 stride-2 convs, 3x3 convs, 1x1 q/k/v .... mg_conv_igemm (split-precision operands per precision.py)
  * 3x3 dilation-2 convs ........................ the SAME dense 3x3 implicit GEMM on the four parity sub-grids of the
                                                  reflect-padded input stacked along the batch axis
  * InstanceNorm + (Leaky)ReLU .................. mg_in_stats / mg_in_apply
  * ConvTranspose2d k4 s2 p1 .................... mg_conv_igemm through ops.conv_dgrad (a transposed conv IS the data
                                                  gradient of the k4 s2 p1 conv with the same weight tensor)
  * 64 -> 3 7x7 head ............................ mg_conv_igemm with the 3 output channels padded to 32
  * 4096-token attention ........................ mg_conv_igemm (scores: Q x per-image K; output: P x per-image V^T) +
                                                  mg_softmax_rows (probabilities written directly as the next operand)
"""
import os

import torch
import torch.nn as nn

from .. import ops, precision
from .base_network import BaseNetwork


# ------------------------------------------------------------------------------------------ pure index arithmetic
def parity_stack(xp):
    """[N, 2a, 2b, C] -> [4N, a, b, C]: the four parity sub-grids (row parity major) stacked along the batch axis.
    A dilation-2 'valid' conv on xp equals the dense 'valid' conv on each sub-grid, interleaved back (parity_unstack)."""
    return torch.cat([xp[:, r::2, c::2] for r in (0, 1) for c in (0, 1)], dim=0).contiguous()


def parity_unstack(y, n):
    """Inverse interleave of parity_stack for the conv result: [4N, a, b, C] -> [N, 2a, 2b, C]."""
    a, b, ch = y.shape[1], y.shape[2], y.shape[3]
    out = torch.empty((n, 2 * a, 2 * b, ch), device=y.device, dtype=y.dtype)
    for i, (r, c) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
        out[:, r::2, c::2] = y[i * n:(i + 1) * n]
    return out


class _ResnetBlock(nn.Module):
    """Parameter container with the reference's layout (generator.py:450-465)."""

    def __init__(self, dim):
        super().__init__()
        self.conv_block = nn.Sequential(
            nn.ReflectionPad2d(2),
            nn.utils.spectral_norm(nn.Conv2d(dim, dim, 3, 1, 0, dilation=2)),
            nn.InstanceNorm2d(dim),
            nn.ReLU(True),
            nn.ReflectionPad2d(1),
            nn.utils.spectral_norm(nn.Conv2d(dim, dim, 3, 1, 0)),
            nn.InstanceNorm2d(dim))


class _SelfAttention(nn.Module):
    def __init__(self, dim, downsample=4):
        super().__init__()
        self.query_conv = nn.Conv2d(dim, dim // downsample, 1)
        self.key_conv = nn.Conv2d(dim, dim // downsample, 1)
        self.value_conv = nn.Conv2d(dim, dim, 1)


class InpaintGenerator(BaseNetwork):
    def __init__(self, opt=None, blocks=12, skips=False):
        super().__init__()
        if skips:
            raise NotImplementedError("michigan_b200: InpaintGenerator(skips=True) is not used by the reference's --use_ig path")
        sn = nn.utils.spectral_norm
        self.blocks = blocks
        self.encoder = nn.Sequential(
            nn.ReflectionPad2d(3), sn(nn.Conv2d(4, 64, 7, padding=0)), nn.InstanceNorm2d(64), nn.LeakyReLU(0.2, True),
            sn(nn.Conv2d(64, 128, 4, 2, 1)), nn.InstanceNorm2d(128), nn.LeakyReLU(0.2, True),
            sn(nn.Conv2d(128, 256, 4, 2, 1)), nn.InstanceNorm2d(256), nn.LeakyReLU(0.2, True))
        self.middle = nn.Sequential(*([_ResnetBlock(256) for _ in range(blocks)] + [_SelfAttention(256)]))
        self.decoder = nn.Sequential(
            sn(nn.ConvTranspose2d(512, 128, 4, 2, 1)), nn.InstanceNorm2d(128), nn.ReLU(True),
            sn(nn.ConvTranspose2d(128, 64, 4, 2, 1)), nn.InstanceNorm2d(64), nn.ReLU(True),
            nn.ReflectionPad2d(3), nn.Conv2d(64, 3, 7, padding=0))
        self.eval()

    # ------------------------------------------------------------------ weights
    @staticmethod
    def _inv_sigma(conv, dim=0):
        """1 / (u^T W v) with the stored u, v (eval-mode spectral norm); a 1-element device tensor."""
        w = conv.weight_orig.detach()
        mat = (w if dim == 0 else w.transpose(0, dim)).reshape(w.shape[dim], -1)
        sigma = torch.dot(conv.weight_u.detach(), torch.mv(mat, conv.weight_v.detach()))
        return (1.0 / sigma).reshape(1).float().contiguous()

    @staticmethod
    def _to_operand(y32, fmt):
        """fp32 NHWC tensor -> tensor-core operand (torch casts: plumbing, used once for the attention input)."""
        if fmt == ops.TF32:
            return (ops.TF32, y32, None)
        hi = y32.to(torch.bfloat16 if fmt == ops.BF16 else torch.float16)
        lo = (y32 - hi.float()).to(hi.dtype) if fmt == ops.BF16 else None
        return (fmt, hi, lo)

    @staticmethod
    def _conv_operand(operand, w, bias, cout, fmt):
        """1x1 conv whose result is only needed as the operand of the next tensor-core product."""
        wp = precision.pack_conv(w, None, operand[0])
        if fmt == ops.TF32:
            return (ops.TF32, precision.conv(operand, wp, cout, 1, 1, 1, 0, bias=bias, round_out=True), None)
        _, hi, lo = precision.conv(operand, wp, cout, 1, 1, 1, 0, bias=bias, out16=(fmt, fmt == ops.BF16), want_f32=False)
        return (fmt, hi, lo)

    def _conv(self, operand, w, inv_sigma, bias, cout, k, stride, pad):
        fmt = operand[0]
        return precision.conv(operand, precision.pack_conv(w, inv_sigma, fmt), cout, k, k, stride, pad, bias=bias)

    @staticmethod
    def _norm_act(y, act, fmt):
        """InstanceNorm + activation, emitted as the operand of the next tensor-core conv (+ fp32 copy)."""
        if fmt == ops.TF32:
            y32 = ops.instance_norm_act(y, act, round_out=True)
            return y32, (ops.TF32, y32, None)
        y32, hi, lo = ops.instance_norm_act(y, act, out16=(fmt, fmt == ops.BF16))
        return y32, (fmt, hi, lo)

    @staticmethod
    def _pad_operand(y32, pad, fmt):
        if fmt == ops.TF32:
            return (ops.TF32, ops.reflect_pad(y32, pad, round_tf32=True), None)
        _, hi, lo = ops.reflect_pad(y32, pad, out16=(fmt, fmt == ops.BF16), want_f32=False)
        return (fmt, hi, lo)

    # ------------------------------------------------------------------ forward
    def forward(self, x):
        """x: [N,4,H,W] fp32 CUDA (orientation RGB with the hole filled by noise + hole mask) -> [N,3,H,W] in [0,1]."""
        if self.training:
            raise RuntimeError("InpaintGenerator is a frozen network: call .eval() (pix2pix_model.py:196-198)")
        return self._forward_impl(x)      # every op below rejects CPU tensors (ops._chk): there is no CPU path

    def _forward_impl(self, x):
        enc, mid, dec = self.encoder, self.middle, self.decoder
        n = x.shape[0]
        with torch.no_grad():
            # ---- encoder
            c1 = enc[1]
            w1 = (c1.weight_orig.detach() * self._inv_sigma(c1)).contiguous()
            h = ops.conv_thin(ops.nchw_to_nhwc(x.contiguous(), 4), ops.pack_weight_thin(w1, 4), c1.bias.detach(), 64, 7, 7, 1, 3, pad_mode=1)
            fmt = precision.conv_fmt(64)
            h32, a = self._norm_act(h, ops.ACT_LRELU, fmt)
            for idx, cout in ((4, 128), (7, 256)):
                conv = enc[idx]
                h = self._conv(a, conv.weight_orig.detach(), self._inv_sigma(conv), conv.bias.detach(), cout, 4, 2, 1)
                h32, a = self._norm_act(h, ops.ACT_LRELU, fmt)
            # ---- 12 dilated residual blocks at 1/4 resolution
            for i in range(self.blocks):
                cb = mid[i].conv_block
                xp = self._pad_operand(h32, 2, fmt)
                stacked = (xp[0], parity_stack(xp[1]), parity_stack(xp[2]) if xp[2] is not None else None)
                y = self._conv(stacked, cb[1].weight_orig.detach(), self._inv_sigma(cb[1]), cb[1].bias.detach(), 256, 3, 1, 0)
                y = parity_unstack(y, n)
                y32 = ops.instance_norm_act(y, ops.ACT_RELU, round_out=(fmt == ops.TF32))
                yp = self._pad_operand(y32, 1, fmt)
                y = self._conv(yp, cb[5].weight_orig.detach(), self._inv_sigma(cb[5]), cb[5].bias.detach(), 256, 3, 1, 0)
                h32 = h32 + ops.instance_norm_act(y, ops.ACT_NONE)
            # ---- self-attention over the (H/4 * W/4) tokens
            att = mid[self.blocks]
            a = self._to_operand(h32, fmt)
            s, t = h32.shape[1], h32.shape[2]
            # softmax(Q K^T) V per image on the implicit-GEMM kernel: scores = 1x1 "conv" of Q with that image's K as the
            # weight operand ([T keys] x [64]); probabilities = mg_softmax_rows, written in operand format; output = 1x1
            # "conv" of the probabilities (T input channels) with V^T ([256] x [T]) as the weight operand.
            T = s * t
            qa = self._conv_operand(a, att.query_conv.weight.detach(), att.query_conv.bias.detach(), 64, fmt)
            k = self._conv(a, att.key_conv.weight.detach(), None, att.key_conv.bias.detach(), 64, 1, 1, 0)            # [n,s,t,64]
            v = self._conv(a, att.value_conv.weight.detach(), None, att.value_conv.bias.detach(), 256, 1, 1, 0)        # [n,s,t,256]
            vt = ops.nhwc_to_nchw(v)                                                                                 # [n,256,s,t]
            o = torch.empty((n, s, t, 256), device=x.device, dtype=torch.float32)
            for i in range(n):
                qi = (qa[0], qa[1][i:i + 1], qa[2][i:i + 1] if qa[2] is not None else None)
                wk = precision.pack_conv(k[i].reshape(T, 64, 1, 1), None, fmt)
                scores = precision.conv(qi, wk, T, 1, 1, 1, 0)                                                       # [1,s,t,T]
                pfmt, phi, plo = ops.softmax_rows(scores.view(T, T), fmt, fmt == ops.BF16)
                pa = (pfmt, phi.view(1, s, t, T), plo.view(1, s, t, T) if plo is not None else None)
                wv = precision.pack_conv(vt[i].reshape(256, T, 1, 1), None, fmt)
                ops.conv_igemm(pa[1], wv, 256, 1, 1, 1, 0, a_fmt=pfmt, x_lo=pa[2], out=o[i:i + 1])
            h32 = torch.cat([h32, o], dim=3).contiguous()
            # ---- decoder: two transposed convs (= data gradients of k4 s2 p1 convs), then the 7x7 head
            for idx, cout in ((0, 128), (3, 64)):
                ct = dec[idx]
                y = ops.conv_dgrad(h32, ct.weight_orig.detach().contiguous(), (2 * h32.shape[1], 2 * h32.shape[2]), 2, 1,
                                   inv_sigma=self._inv_sigma(ct, dim=1))
                y = y + ct.bias.detach().view(1, 1, 1, -1)
                h32 = ops.instance_norm_act(y, ops.ACT_RELU)
            head = dec[7]
            w = torch.zeros((32, 64, 7, 7), device=x.device, dtype=torch.float32)
            w[:3] = head.weight.detach()
            b = torch.zeros(32, device=x.device, dtype=torch.float32)
            b[:3] = head.bias.detach()
            y = self._conv(self._pad_operand(h32, 3, fmt), w, None, b, 32, 7, 1, 0)
            out = (torch.tanh(y[..., :3]) + 1) / 2
            return out.permute(0, 3, 1, 2).contiguous()
