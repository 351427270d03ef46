"""Training path: hand-written backward passes wired into torch.autograd as two Functions
(generator, multiscale discriminator).  Forward = the same kernels as inference with fp32-stored
TF32 operands (so that the saved activations are directly the operands of the gradient GEMMs);
backward = explicit reverse pass over the saved per-block state:

    conv  : dW = mg_conv_wgrad (tcgen05, MN-major, split-K);  dX = mg_conv_igemm on dY with flipped sub-kernels
    SPADE : mg_spade_bwd (dgamma|dbeta operand, dxhat, BN sums)  ->  gamma/beta GEMM wgrad + dgrad  ->  thin wgrad
    BN/IN : mg_bn_bwd_apply / mg_in_bwd (statistics all-reduced across ranks like the forward ones)
    SN    : mg_spectral_norm_bwd (u, v constants, as torch's spectral_norm autograd)

What is saved per SPADE: h (conv operand), 1+gamma, the normalisation vectors; `actv` is recomputed.
"""
from types import SimpleNamespace

import torch

from .. import ops
from .sync_batchnorm import _world, allreduce_sums

_RELU, _LRELU, _NONE = ops.ACT_RELU, ops.ACT_LRELU, ops.ACT_NONE


def _nhwc(g):
    """NCHW-shaped gradient (any strides) -> contiguous NHWC tensor that this backward owns (it is
    used as an accumulation target, so never alias the tensor autograd handed in)."""
    t = g.permute(0, 2, 3, 1)
    return t.clone(memory_format=torch.contiguous_format)


class _Grads:
    """param -> accumulated gradient."""

    def __init__(self):
        self.d = {}

    def add(self, p, g):
        if p is None or g is None:
            return
        g = g.reshape(p.shape)
        k = id(p)
        self.d[k] = g if k not in self.d else self.d[k] + g

    def get(self, p):
        return self.d.get(id(p))


def _thin_wt_to_oihw(dwt, kh, kw, cin):
    """[kh*kw][CinP][Cout] -> [Cout, cin, kh, kw]."""
    cp, co = dwt.shape[1], dwt.shape[2]
    return dwt.view(kh, kw, cp, co).permute(3, 2, 0, 1)[:, :cin].contiguous()


def _gb_unpack_index(c, device):
    """packed gamma|beta row order -> (gamma idx, beta idx) into the [2C] packed vector."""
    bn = ops.spade_bn(c)
    half = bn // 2
    ch = torch.arange(c, device=device)
    tile, r = ch // half, ch % half
    return tile * bn + r, tile * bn + half + r


# =============================================================================================== conv helpers
def _conv_weight(conv, inv_of):
    """(w_oihw source, inv_sigma or None, is_sn)."""
    if hasattr(conv, "weight_orig"):
        return conv.weight_orig, inv_of[conv], True
    return conv.weight, None, False


def _pack(conv, inv_of):
    w, isg, _ = _conv_weight(conv, inv_of)
    return ops.pack_weight(w.detach(), isg, True)


def _conv_param_grads(G, conv, inv_of, dy, a_operand, kh, kw, stride, pad, with_bias=True, dz_for_bias=None):
    """Weight (+bias) gradients of an implicit-GEMM conv; dy: [N,OH,OW,Cout], a_operand: its fp32 input."""
    w, isg, is_sn = _conv_weight(conv, inv_of)
    dwt = ops.unpack_wgrad(ops.conv_wgrad(dy, a_operand, kh, kw, stride, pad), tuple(w.shape))
    if is_sn:
        G.add(w, ops.spectral_norm_bwd(dwt, w.detach(), conv.weight_u, conv.weight_v, isg))
    else:
        G.add(w, dwt)
    if with_bias and getattr(conv, "bias", None) is not None:
        G.add(conv.bias, ops.chan_sum(dz_for_bias if dz_for_bias is not None else dy))


# =============================================================================================== SPADE + conv
def _spade_fwd(blk, name, src, shift, ns, nh, act, seg4, R, hw):
    sp = getattr(blk, name)
    wsh = ops.pack_mlp_shared(sp.mlp_shared[0].weight.detach())
    actv = ops.mlp_shared(seg4, wsh, sp.mlp_shared[0].bias.detach(), seg_resize=R, act=_RELU, round_out=True, out_hw=hw)
    c = src.shape[-1]
    g1 = torch.empty((src.shape[0], hw[0], hw[1], c), device=src.device, dtype=torch.float32)
    wgb = ops.pack_weight_gb(sp.mlp_gamma.weight.detach(), sp.mlp_beta.weight.detach())
    h = ops.conv_igemm(actv, wgb, c, 3, 3, 1, 1, act=act, round_out=True,
                       spade=(src, shift, ns, nh, (sp.mlp_gamma.bias.detach() + 1.0).contiguous(), sp.mlp_beta.bias.detach()), aux=g1)
    return SimpleNamespace(sp=sp, src=src, shift=shift, ns=ns, nh=nh, act=act, g1=g1, h=h, wsh=wsh, R=R, hw=hw)


def _spade_conv_bwd(G, blk, S, conv, inv_of, dy, k, pad, seg4, seg_cache=None):
    """Backward of conv(act(SPADE(src))) given dy; returns (dxhat, sums) for the BN backward of `src`."""
    _conv_param_grads(G, conv, inv_of, dy, S.h, k, k, 1, pad)
    w, isg, _ = _conv_weight(conv, inv_of)
    dh = ops.conv_dgrad(dy, w.detach(), S.hw, 1, pad, inv_sigma=isg)
    dgb, dxhat, sums = ops.spade_bwd(dh, S.h, S.g1, S.src, S.shift, S.ns, S.nh, S.act)
    del dh
    sp = S.sp
    c = S.src.shape[-1]
    actv = ops.mlp_shared(seg4, S.wsh, sp.mlp_shared[0].bias.detach(), seg_resize=S.R, act=_RELU, round_out=True, out_hw=S.hw)
    dwg, dwb = ops.unpack_wgrad_gb(ops.conv_wgrad(dgb, actv, 3, 3, 1, 1), c, 128)
    G.add(sp.mlp_gamma.weight, dwg)
    G.add(sp.mlp_beta.weight, dwb)
    gi, bi = _gb_unpack_index(c, dgb.device)
    bsum = ops.chan_sum(dgb)
    G.add(sp.mlp_gamma.bias, bsum[gi])
    G.add(sp.mlp_beta.bias, bsum[bi])
    dactv = ops.conv_igemm(dgb, ops.pack_weight_dgrad_gb(sp.mlp_gamma.weight.detach(), sp.mlp_beta.weight.detach()), 128, 3, 3, 1, 1)
    del dgb
    da = ops.act_bwd(dactv, actv, _RELU)
    del dactv, actv
    if ops.thin_wgrad_tc_enabled():
        seg32 = seg_cache.get(S.R) if seg_cache is not None else None
        if seg32 is None:
            seg32 = ops.pad_channels32(seg4, seg_resize=S.R, in_hw=S.hw)
            if seg_cache is not None:
                seg_cache.clear()          # one resolution at a time is alive (blocks run coarse -> fine in reverse)
                seg_cache[S.R] = seg32
        dwt = ops.thin_wgrad_tc(seg32, da, 3, 3, 1, 1, 4)
    else:
        dwt = ops.thin_wgrad(seg4, da, 3, 3, 1, 1, seg_resize=S.R, in_hw=S.hw)
    G.add(sp.mlp_shared[0].weight, _thin_wt_to_oihw(dwt, 3, 3, 4))
    G.add(sp.mlp_shared[0].bias, ops.chan_sum(da))
    return dxhat, allreduce_sums(sums)


# =============================================================================================== SPADEResnetBlock
def block_fwd(blk, x, xs, seg4, inv_of, blend):
    N, hs, ws, fin = x.shape
    h, w = hs << xs, ws << xs
    R = seg4.shape[1] // h
    extra = (blk.norm_s.param_free_norm,) if blk.learned_shortcut else ()
    ns0, nh0 = blk.norm_0.param_free_norm.scale_shift(x, xs, extra)[:2]
    S = SimpleNamespace(x=x, xs=xs, blend=blend, hw=(h, w), count0=N * h * w * _world())
    if blk.learned_shortcut:
        S.sps = _spade_fwd(blk, "norm_s", x, xs, ns0, nh0, _NONE, seg4, R, (h, w))
        x_s = ops.conv_igemm(S.sps.h, _pack(blk.conv_s, inv_of), blk.fout, 1, 1, 1, 0)
        res, rshift = x_s, 0
    else:
        res, rshift = x, xs
    S.sp0 = _spade_fwd(blk, "norm_0", x, xs, ns0, nh0, _LRELU, seg4, R, (h, w))
    S.dx = ops.conv_igemm(S.sp0.h, _pack(blk.conv_0, inv_of), blk.fmiddle, 3, 3, 1, 1, bias=blk.conv_0.bias.detach())
    ns1, nh1 = blk.norm_1.param_free_norm.scale_shift(S.dx)[:2]
    S.sp1 = _spade_fwd(blk, "norm_1", S.dx, 0, ns1, nh1, _LRELU, seg4, R, (h, w))
    out = ops.conv_igemm(S.sp1.h, _pack(blk.conv_1, inv_of), blk.fout, 3, 3, 1, 1, bias=blk.conv_1.bias.detach(), res=res, res_shift=rshift, blend=blend)
    return out, S


def block_bwd(G, blk, S, dout, seg4, inv_of):
    """-> (dx wrt the block input (pre-upsample), dbf wrt the blended background feature or None)."""
    seg_cache = {}
    dbf = None
    if S.blend is not None:
        _, hair, back, ms = S.blend
        dy, dbf = ops.blend_bwd(dout, hair, back, ms)
    else:
        dy = dout
    N = dy.shape[0]
    h, w = S.hw
    dxhat1, sums1 = _spade_conv_bwd(G, blk, S.sp1, blk.conv_1, inv_of, dy, 3, 1, seg4, seg_cache)
    ddx = ops.bn_bwd_apply(dxhat1, S.dx, 0, S.sp1.ns, S.sp1.nh, sums1, N * h * w * _world())
    del dxhat1
    dxhat0, sums0 = _spade_conv_bwd(G, blk, S.sp0, blk.conv_0, inv_of, ddx, 3, 1, seg4, seg_cache)
    del ddx
    dx = ops.bn_bwd_apply(dxhat0, S.x, S.xs, S.sp0.ns, S.sp0.nh, sums0, S.count0)
    del dxhat0
    if blk.learned_shortcut:
        dxhat_s, sums_s = _spade_conv_bwd(G, blk, S.sps, blk.conv_s, inv_of, dy, 1, 0, seg4, seg_cache)
        ops.bn_bwd_apply(dxhat_s, S.x, S.xs, S.sps.ns, S.sps.nh, sums_s, S.count0, dx=dx)
    else:
        ops.bn_bwd_apply(dy, S.x, S.xs, None, None, None, 1, dx=dx)   # identity shortcut through the upsample
    return dx, dbf


# =============================================================================================== encoders
def fc_fwd(fc, image_ref, label_ref0, label_tag0):
    N, _, H, W = image_ref.shape
    mref = label_ref0.reshape(N, H, W).contiguous()
    mtag = label_tag0.reshape(N, H, W).contiguous()
    S = SimpleNamespace(mref=mref, mtag=mtag, layers=[])
    S.x0 = ops.nchw_to_nhwc(image_ref.contiguous(), 4, pmul=mref)
    ratio, upd = ops.partial_mask(mref, 3, 2, 1)
    S.wt1 = ops.pack_weight_thin(fc.layer1.weight.detach(), 4)
    y = ops.conv_thin(S.x0, S.wt1, fc.layer1.bias.detach(), fc.layer1.out_channels, 3, 3, 2, 1, pscale=ratio, pmul=upd)
    S.l1 = SimpleNamespace(ratio=ratio, upd=upd, y=y)
    prev_upd = upd
    for i in range(2, 6):
        a, ss = ops.instance_norm_act_fwd(y, _LRELU, 1e-5, round_out=True, pmul=prev_upd)
        ratio, upd = ops.partial_mask(prev_upd, 3, 2, 1)
        layer = getattr(fc, "layer%d" % i)
        y_next = ops.conv_igemm(a, ops.pack_weight(layer.weight.detach(), None, True), layer.out_channels, 3, 3, 2, 1,
                                bias=layer.bias.detach(), pscale=ratio, pmul=upd)
        S.layers.append(SimpleNamespace(layer=layer, a=a, ss=ss, y_in=y, pm_in=prev_upd, ratio=ratio, upd=upd))
        y, prev_upd = y_next, upd
    a6, ss6 = ops.instance_norm_act_fwd(y, _LRELU, 1e-5)
    S.y5, S.ss6 = y, ss6
    m = ops.masked_mean_bcast(a6, mref, mtag)
    S.mhw = (m.shape[1], m.shape[2])
    out = ops.resize_bilinear(m, fc.sh, fc.sw) if fc.sh != m.shape[1] else m
    return out, S


def fc_bwd(G, fc, S, dout):
    d = ops.resize_bilinear_bwd(dout, S.mhw) if (dout.shape[1], dout.shape[2]) != S.mhw else dout
    d = ops.masked_mean_bcast_bwd(d, S.mref, S.mtag)
    dy = ops.in_bwd(d, S.y5, S.ss6, _LRELU)
    for L in reversed(S.layers):
        dz = ops.act_bwd(dy, None, _NONE, pm1=L.upd, pm2=L.ratio, round_tf32=True)   # d(acc) of (acc*ratio + b)*upd
        G.add(L.layer.bias, ops.chan_sum(ops.act_bwd(dy, None, _NONE, pm1=L.upd)))
        G.add(L.layer.weight, ops.unpack_wgrad(ops.conv_wgrad(dz, L.a, 3, 3, 2, 1), tuple(L.layer.weight.shape)))
        da = ops.conv_dgrad(dz, L.layer.weight.detach(), (L.a.shape[1], L.a.shape[2]), 2, 1)
        dy = ops.in_bwd(da, L.y_in, L.ss, _LRELU, pmul=L.pm_in)
    dz = ops.act_bwd(dy, None, _NONE, pm1=S.l1.upd, pm2=S.l1.ratio)
    G.add(fc.layer1.bias, ops.chan_sum(ops.act_bwd(dy, None, _NONE, pm1=S.l1.upd)))
    dwt1 = (ops.thin_wgrad_tc(ops.pad_channels32(S.x0), dz, 3, 3, 2, 1, 4) if ops.thin_wgrad_tc_enabled()
            else ops.thin_wgrad(S.x0, dz, 3, 3, 2, 1))
    G.add(fc.layer1.weight, _thin_wt_to_oihw(dwt1, 3, 3, 3))


def bg_fwd(bg, image, mask, noise):
    back = bg.back_mask(mask)
    inp = ops.nchw_to_nhwc(noise.contiguous(), 4) if bg.opt.random_noise_background else \
        ops.prep_bginput(image.contiguous(), noise.contiguous(), back)
    S = SimpleNamespace(inp=inp, layers=[])
    S.wt1 = ops.pack_weight_thin(bg.conv1.conv.weight.detach(), 4)
    x = ops.conv_thin(inp, S.wt1, bg.conv1.conv.bias.detach(), bg.ngf, 7, 7, 1, 3, pad_mode=1, act=_RELU)
    S.x0 = x
    feats = [x]
    for name in ("layer1", "layer2", "layer3"):
        blk = getattr(bg, name)
        xp = ops.reflect_pad(x, 1, round_tf32=True)
        y = ops.conv_igemm(xp, ops.pack_weight(blk.conv.weight.detach(), None, True), blk.conv.out_channels, 4, 4, 2, 0,
                           bias=blk.conv.bias.detach(), act=_RELU)
        S.layers.append(SimpleNamespace(blk=blk, xp=xp, y=y))
        x = y
        feats.append(y)
    return feats[::-1], back, S


def bg_bwd(G, bg, S, dfeats):
    """dfeats: grads of [x3, x2, x1, x0] (from the four background blends)."""
    d_list = dfeats[::-1]   # [dx0, dx1, dx2, dx3]
    d = d_list[3]
    for i in (2, 1, 0):
        L = S.layers[i]
        dz = ops.act_bwd(d, L.y, _RELU, round_tf32=True)
        G.add(L.blk.conv.bias, ops.chan_sum(dz))
        G.add(L.blk.conv.weight, ops.unpack_wgrad(ops.conv_wgrad(dz, L.xp, 4, 4, 2, 0), tuple(L.blk.conv.weight.shape)))
        dxp = ops.conv_dgrad(dz, L.blk.conv.weight.detach(), (L.xp.shape[1], L.xp.shape[2]), 2, 0)
        d = ops.reflect_pad_bwd(dxp, 1, dx=d_list[i])
    dz0 = ops.act_bwd(d, S.x0, _RELU)
    G.add(bg.conv1.conv.bias, ops.chan_sum(dz0))
    dwt1 = (ops.thin_wgrad_tc(ops.pad_channels32(S.inp, reflect_pad=3), dz0, 7, 7, 1, 0, 4) if ops.thin_wgrad_tc_enabled()
            else ops.thin_wgrad(S.inp, dz0, 7, 7, 1, 3, pad_mode=1))
    G.add(bg.conv1.conv.weight, _thin_wt_to_oihw(dwt1, 7, 7, 3))


# =============================================================================================== generator Function
class _GeneratorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, netG, input_ref, orient_mask, image_ref, input_tag, noise, image_tag, *params):
        opt = netG.opt
        input_tag = input_tag.contiguous()
        seg4 = ops.prep_seg(input_tag, orient_mask.contiguous())
        x, Sfc = fc_fwd(netG.fc, image_ref, input_ref[:, 1:2], input_tag[:, 1:2])
        feats, back, Sbg = bg_fwd(netG.backgroud_enc, image_tag, input_tag, noise)
        hair = input_tag[:, 1].contiguous()
        snb = netG.spectral_batch()
        inv = snb.run(netG.training)
        inv_of = {c: inv[i:i + 1].clone() for i, c in enumerate(snb.convs)}
        saved = []
        x, S = block_fwd(netG.head_0, x, 0, seg4, inv_of, None); saved.append(S)
        x, S = block_fwd(netG.G_middle_0, x, 1, seg4, inv_of, None); saved.append(S)
        x, S = block_fwd(netG.G_middle_1, x, 1, seg4, inv_of, None); saved.append(S)
        for i in range(4):
            x, S = block_fwd(getattr(netG, "up_%d" % i), x, 1, seg4, inv_of, (feats[i], hair, back, 8 >> i)); saved.append(S)
        out = ops.conv_img(x, netG.conv_img.weight.detach(), netG.conv_img.bias.detach())
        ctx.netG, ctx.seg4, ctx.inv_of, ctx.saved, ctx.Sfc, ctx.Sbg, ctx.x_last, ctx.out = netG, seg4, inv_of, saved, Sfc, Sbg, x, out
        ctx.params = params
        return out

    @staticmethod
    def backward(ctx, dout):
        netG, seg4, inv_of = ctx.netG, ctx.seg4, ctx.inv_of
        G = _Grads()
        dx, dw, db = ops.conv_img_bwd(dout.contiguous(), ctx.out, ctx.x_last, netG.conv_img.weight.detach())
        G.add(netG.conv_img.weight, dw)
        G.add(netG.conv_img.bias, db)
        names = ["head_0", "G_middle_0", "G_middle_1", "up_0", "up_1", "up_2", "up_3"]
        dfeats = [None] * 4
        for idx in range(6, -1, -1):
            blk = getattr(netG, names[idx])
            dx, dbf = block_bwd(G, blk, ctx.saved[idx], dx, seg4, inv_of)
            ctx.saved[idx] = None
            if idx >= 3:
                dfeats[idx - 3] = dbf
        fc_bwd(G, netG.fc, ctx.Sfc, dx)
        bg_bwd(G, netG.backgroud_enc, ctx.Sbg, dfeats)
        grads = tuple(G.get(p) for p in ctx.params)
        ctx.saved = ctx.Sfc = ctx.Sbg = None
        return (None,) * 7 + grads


def generator_forward_autograd(netG, input_ref, orient_mask, image_ref, input_tag, noise, image_tag):
    params = [p for p in netG.parameters()]
    return _GeneratorFn.apply(netG, input_ref, orient_mask, image_ref, input_tag, noise, image_tag, *params)


# =============================================================================================== discriminator Function
def _d_scale_fwd(D, x8, inv_of):
    S = SimpleNamespace(x8=x8, layers=[])
    conv0 = D.model0[0]
    S.wt0 = ops.pack_weight_thin(conv0.weight.detach(), 8)
    f = ops.conv_thin(x8, S.wt0, conv0.bias.detach(), conv0.out_channels, 4, 4, 2, D.padw, act=_LRELU, round_out=True)
    S.f0 = f
    outs = [f]
    for n, conv in zip(range(1, D.n_layers), D.mid_convs()):
        w, isg, _ = _conv_weight(conv, inv_of)
        raw = ops.conv_igemm(f, ops.pack_weight(w.detach(), isg, True), conv.out_channels, 4, 4, D._strides[n], D.padw)
        f_next, ss = ops.instance_norm_act_fwd(raw, _LRELU, 1e-5, round_out=True)
        S.layers.append(SimpleNamespace(conv=conv, stride=D._strides[n], f_in=f, raw=raw, ss=ss))
        f = f_next
        outs.append(f)
    last = getattr(D, "model%d" % D.n_layers)[0]
    S.f_last, S.last = f, last
    outs.append(ops.conv_to1(f, last.weight.detach(), last.bias.detach(), D.padw))
    return outs, S


def _d_scale_bwd(G, D, S, douts, inv_of, need_dimg, param_grads):
    """douts: list of NHWC grads (or None) for [f0..f3, logits]; returns dimg NCHW [B,3,H,W] or None."""
    nl = D.n_layers
    df = douts[nl - 1]
    if douts[nl] is not None:
        dxl, dw, db = ops.conv_to1_bwd(douts[nl], S.f_last, S.last.weight.detach(), D.padw, dx=df, want_dx=True)
        df = dxl
        if param_grads:
            G.add(S.last.weight, dw)
            G.add(S.last.bias, db)
    for n in range(nl - 1, 0, -1):
        L = S.layers[n - 1]
        below = douts[n - 1]
        if df is None:
            df = below
            continue
        draw = ops.in_bwd(df, L.raw, L.ss, _LRELU, round_tf32=True)
        w, isg, _ = _conv_weight(L.conv, inv_of)
        if param_grads:
            _conv_param_grads(G, L.conv, inv_of, draw, L.f_in, 4, 4, L.stride, D.padw, with_bias=False)
        df = ops.conv_dgrad(draw, w.detach(), (L.f_in.shape[1], L.f_in.shape[2]), L.stride, D.padw, inv_sigma=isg,
                            out=below, accumulate=below is not None)
    if df is None:
        return None
    conv0 = D.model0[0]
    dz0 = ops.act_bwd(df, S.f0, _LRELU)
    if param_grads:
        dwt0 = (ops.thin_wgrad_tc(ops.pad_channels32(S.x8), dz0, 4, 4, 2, D.padw, 8) if ops.thin_wgrad_tc_enabled()
                else ops.thin_wgrad(S.x8, dz0, 4, 4, 2, D.padw))
        G.add(conv0.weight, _thin_wt_to_oihw(dwt0, 4, 4, conv0.weight.shape[1]))
        G.add(conv0.bias, ops.chan_sum(dz0))
    if not need_dimg:
        return None
    B, H, W, _ = S.x8.shape
    dimg = torch.zeros((B, 3, H, W), device=dz0.device, dtype=torch.float32)
    ops.thin_dgrad3(dz0, S.wt0, dimg, 4, 4, 2, D.padw, 4)
    return dimg


class _DiscriminatorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, netD, x_nchw, *params):
        snb = netD.spectral_batch()
        inv = snb.run(netD.training)
        inv_of = {cv: inv[i:i + 1].clone() for i, cv in enumerate(snb.convs)} if inv is not None else {}
        x8 = ops.nchw_to_nhwc(x_nchw.contiguous(), 8)
        states, flat = [], []
        children = [D for _, D in netD.named_children()]
        for i, D in enumerate(children):
            outs, S = _d_scale_fwd(D, x8, inv_of)
            states.append(S)
            flat += [o.permute(0, 3, 1, 2) for o in outs]
            if i + 1 < len(children):
                x8 = ops.avgpool3s2(x8)
        ctx.netD, ctx.states, ctx.inv_of, ctx.params = netD, states, inv_of, params
        ctx.need_dimg = x_nchw.requires_grad
        ctx.param_grads = any(p.requires_grad for p in params)
        ctx.in_shape = tuple(x_nchw.shape)
        return tuple(flat)

    @staticmethod
    def backward(ctx, *gouts):
        netD = ctx.netD
        G = _Grads()
        children = [D for _, D in netD.named_children()]
        per = children[0].n_layers + 1
        dimg_total = None
        for i in range(len(children) - 1, -1, -1):
            douts = [(_nhwc(g) if g is not None else None) for g in gouts[i * per:(i + 1) * per]]
            dimg = _d_scale_bwd(G, children[i], ctx.states[i], douts, ctx.inv_of, ctx.need_dimg, ctx.param_grads)
            if ctx.need_dimg and dimg is not None:
                if dimg_total is None:
                    dimg_total = dimg
                else:
                    # gradient of the finer scale's avg-pooled input (dimg_total is at the coarser scale)
                    fine = torch.zeros((dimg.shape[0], dimg.shape[2], dimg.shape[3], 4), device=dimg.device)
                    ops.avgpool3s2_bwd(ops.nchw_to_nhwc(dimg_total, 4), fine)
                    dimg_total = dimg + ops.nhwc_to_nchw(fine, 3)
            elif ctx.need_dimg and dimg_total is not None:
                fine_hw = ctx.states[i].x8.shape
                fine = torch.zeros((fine_hw[0], fine_hw[1], fine_hw[2], 4), device=dimg_total.device)
                ops.avgpool3s2_bwd(ops.nchw_to_nhwc(dimg_total, 4), fine)
                dimg_total = ops.nhwc_to_nchw(fine, 3)
        dx = None
        if ctx.need_dimg:
            dx = torch.zeros(ctx.in_shape, device=gouts[0].device if gouts[0] is not None else None, dtype=torch.float32)
            if dimg_total is not None:
                dx[:, 4:7] = dimg_total
        grads = tuple(G.get(p) for p in ctx.params) if ctx.param_grads else (None,) * len(ctx.params)
        ctx.states = None
        return (None, dx) + grads


def discriminator_forward_autograd(netD, x_nchw):
    params = [p for p in netD.parameters()]
    flat = _DiscriminatorFn.apply(netD, x_nchw, *params)
    children = [D for _, D in netD.named_children()]
    per = children[0].n_layers + 1
    get_feats = not netD.opt.no_ganFeat_loss
    result = []
    for i in range(len(children)):
        outs = list(flat[i * per:(i + 1) * per])
        result.append(outs if get_feats else [outs[-1]])
    return result
