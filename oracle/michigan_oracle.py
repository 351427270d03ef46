"""CPU oracle for the MichiGAN hot path — TEST INFRASTRUCTURE ONLY.

A functional restatement, on the CPU, of the reference's generator / discriminator / loss path
(reference = tzt101/MichiGAN @ 3159610; file:line cited per function).  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py` may import
this module; the product (`michigan_b200/`) never does.

Where the arithmetic lives: the reference has no kernels of its own — every op is a call into
PyTorch (requirements.txt:1 `torch>=1.0.0`, unpinned; this image pins torch 2.11.0+cu128).  The
restatement therefore spells each step out with `torch.nn.functional` primitives on CPU tensors
(fp32, or fp64 when the state dict is double) instead of going through the reference's nn.Module
graph, options, DataParallel or hooks.

Pinning: the reference ships no tests or golden vectors for this path.  The
oracle is pinned against the reference itself, imported and run in the build container:
`tests/golden/make_golden.py` drives the unmodified reference modules and this oracle on identical
weights/inputs, asserts agreement, and commits the reference's outputs as fixtures under
`tests/golden/` which `tests/test_oracle_golden.py` re-checks everywhere (no reference needed).
"""
import math
import random as _pyrandom
from types import SimpleNamespace

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------- options
def default_opt(**kw):
    """The subset of the reference's argparse namespace the path reads (options/base_options.py:22-131,
    options/train_options.py:18-78), at the README train/inference flag values."""
    o = SimpleNamespace(
        ngf=64, ndf=64, crop_size=512, aspect_ratio=1.0, num_upsampling_layers="more",
        label_nc=2, orient_nc=2, output_nc=3, semantic_nc=2, use_ig=False, isTrain=True,
        add_feat_zeros=False, add_th=64, noise_background=True, random_expand_mask=True,
        random_expand_th=0.05, expand_mask_be=True, expand_th=5, random_noise_background=False,
        bf_direct_add=False, norm_ref_encode="instance", num_D=2, n_layers_D=4,
        no_ganFeat_loss=False, lambda_feat=1.0, wide_edge=2.0, gan_mode="hinge",
        remove_background=False,
    )
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def lrelu(x):
    return F.leaky_relu(x, 0.2)


# ----------------------------------------------------------------------------------------------- spectral norm
def spectral_weight(sd, prefix, training, eps=1e-12):
    """W_orig / sigma with torch's old-style spectral_norm semantics (torch/nn/utils/spectral_norm.py
    `SpectralNorm.compute_weight`), as applied at architecture.py:38-42 and normalization.py:28-29.
    Training: one in-place power iteration on the stored u, v before use (also under no_grad)."""
    w = sd[prefix + ".weight_orig"]
    u = sd[prefix + ".weight_u"]
    v = sd[prefix + ".weight_v"]
    mat = w.reshape(w.shape[0], -1)
    if training:
        with torch.no_grad():
            v_new = F.normalize(torch.mv(mat.t(), u), dim=0, eps=eps)
            u_new = F.normalize(torch.mv(mat, v_new), dim=0, eps=eps)
            v.copy_(v_new)
            u.copy_(u_new)
        u = u.clone()
        v = v.clone()
    sigma = torch.dot(u, torch.mv(mat, v))
    return w / sigma


def conv_weight(sd, prefix, training):
    if prefix + ".weight_orig" in sd:
        return spectral_weight(sd, prefix, training)
    return sd[prefix + ".weight"]


# ----------------------------------------------------------------------------------------------- batch norm
def param_free_bn(x, sd, prefix, training, world_sums=None, record=None, momentum=0.1, eps=1e-5):
    """SynchronizedBatchNorm2d(affine=False) (sync_batchnorm/batchnorm.py:63-93,128-145).

    Single-replica / eval: F.batch_norm semantics (batchnorm.py:65-68): train -> biased batch variance,
    1/sqrt(var+eps), running stats momentum 0.1 with the UNBIASED variance; eval -> running stats.
    num_batches_tracked is never incremented by the reference (forward bypasses nn.BatchNorm.forward).
    `world_sums(s, ss, n)` (optional) emulates the data-parallel master: it receives this replica's
    sum / square-sum / count and returns the global ones; that path uses clamp(var, eps)^-0.5
    (batchnorm.py:128-145).
    """
    rm = sd[prefix + ".running_mean"]
    rv = sd[prefix + ".running_var"]
    if not training:
        return (x - rm.view(1, -1, 1, 1)) / torch.sqrt(rv.view(1, -1, 1, 1) + eps)
    n = x.shape[0] * x.shape[2] * x.shape[3]
    if world_sums is None:
        mean = x.mean(dim=(0, 2, 3))
        var = x.var(dim=(0, 2, 3), unbiased=False)
        inv_std = 1.0 / torch.sqrt(var + eps)
    else:
        s = x.sum(dim=(0, 2, 3))
        ss = (x * x).sum(dim=(0, 2, 3))
        s, ss, n = world_sums(s, ss, n)
        mean = s / n
        sumvar = ss - s * mean
        var = sumvar / n
        inv_std = var.clamp(min=eps) ** -0.5
    with torch.no_grad():
        unbiased = var.detach() * n / (n - 1)
        rm.copy_((1 - momentum) * rm + momentum * mean.detach())
        rv.copy_((1 - momentum) * rv + momentum * unbiased)
    if record is not None:
        record[prefix] = (mean.detach().clone(), var.detach().clone(), n)
    return (x - mean.view(1, -1, 1, 1)) * inv_std.view(1, -1, 1, 1)


# ----------------------------------------------------------------------------------------------- SPADE
def spade(x, seg, sd, prefix, training, **bn_kw):
    """SPADE.forward (normalization.py:101-118): param-free BN, nearest-resized seg -> 3x3 conv + ReLU
    (128 hidden) -> gamma, beta 3x3 convs -> normalized * (1 + gamma) + beta."""
    normalized = param_free_bn(x, sd, prefix + ".param_free_norm", training, **bn_kw)
    s = F.interpolate(seg, size=x.shape[2:], mode="nearest")
    actv = F.relu(F.conv2d(s, sd[prefix + ".mlp_shared.0.weight"], sd[prefix + ".mlp_shared.0.bias"], padding=1))
    gamma = F.conv2d(actv, sd[prefix + ".mlp_gamma.weight"], sd[prefix + ".mlp_gamma.bias"], padding=1)
    beta = F.conv2d(actv, sd[prefix + ".mlp_beta.weight"], sd[prefix + ".mlp_beta.bias"], padding=1)
    return normalized * (1 + gamma) + beta


def spade_resnet_block(x, seg, sd, prefix, training, **bn_kw):
    """SPADEResnetBlock.forward (architecture.py:67-85).  The learned shortcut exists iff fin != fout
    (architecture.py:27), i.e. iff the state dict holds `<prefix>.conv_s.*`; it has NO activation."""
    if prefix + ".conv_s.weight_orig" in sd or prefix + ".conv_s.weight" in sd:
        x_s = F.conv2d(spade(x, seg, sd, prefix + ".norm_s", training, **bn_kw), conv_weight(sd, prefix + ".conv_s", training))
    else:
        x_s = x
    dx = F.conv2d(lrelu(spade(x, seg, sd, prefix + ".norm_0", training, **bn_kw)), conv_weight(sd, prefix + ".conv_0", training),
                  sd[prefix + ".conv_0.bias"], padding=1)
    dx = F.conv2d(lrelu(spade(dx, seg, sd, prefix + ".norm_1", training, **bn_kw)), conv_weight(sd, prefix + ".conv_1", training),
                  sd[prefix + ".conv_1.bias"], padding=1)
    return x_s + dx


# ----------------------------------------------------------------------------------------------- encoders
def partial_conv(x, mask, w, b, stride=2, padding=1):
    """PartialConv2d.forward, single-channel mask, return_mask=True (partialconv2d.py:46-85)."""
    kh, kw = w.shape[2], w.shape[3]
    with torch.no_grad():
        ones = torch.ones(1, 1, kh, kw, dtype=x.dtype)
        update_mask = F.conv2d(mask, ones, None, stride=stride, padding=padding)
        mask_ratio = (kh * kw) / (update_mask + 1e-8)
        update_mask = torch.clamp(update_mask, 0, 1)
        mask_ratio = mask_ratio * update_mask
    raw = F.conv2d(x * mask, w, b, stride=stride, padding=padding)
    bv = b.view(1, -1, 1, 1)
    out = ((raw - bv) * mask_ratio + bv) * update_mask
    return out, update_mask


def image_encoder3(image_ref, label_ref0, label_tag0, sd, prefix, sh, sw):
    """ImageEncoder3.forward, norm_ref_encode='instance' (encoder.py:190-225)."""
    x, mask = partial_conv(image_ref, label_ref0, sd[prefix + ".layer1.weight"], sd[prefix + ".layer1.bias"])
    x = F.instance_norm(x)
    for i in range(2, 6):
        x, mask = partial_conv(lrelu(x), mask, sd[prefix + ".layer%d.weight" % i], sd[prefix + ".layer%d.bias" % i])
        x = F.instance_norm(x)
    x = lrelu(x)
    xh, xw = x.shape[2:]
    label_ref = F.interpolate(label_ref0, size=(xh, xw), mode="nearest")
    label_tag = F.interpolate(label_tag0, size=(xh, xw), mode="nearest")
    outs = []
    for b in range(x.shape[0]):
        tmps = x[b] * label_ref[b]
        denom = torch.clamp(label_ref[b].sum(), min=1)  # max(torch.sum(label_ref[b]), 1), encoder.py:218
        tmps = tmps.sum(dim=(1, 2), keepdim=True) / denom
        outs.append(tmps.expand_as(x[b]) * label_tag[b])
    out = torch.stack(outs, 0)
    if sh != xh:
        out = F.interpolate(out, size=(sh, sw), mode="bilinear")
    return out


def conv_block(x, sd, prefix, k, stride, pad):
    """ConvBlock(norm='none', activation='relu', pad_type='reflect') (MaskGAN_networks.py:114-173)."""
    x = F.pad(x, (pad, pad, pad, pad), mode="reflect")
    return F.relu(F.conv2d(x, sd[prefix + ".conv.weight"], sd[prefix + ".conv.bias"], stride=stride))


def background_back_mask(mask, opt, training, rng_k=None):
    """The dilated-hair complement used by BackgroundEncode2 (encoder.py:286-316).  `rng_k` overrides
    the reference's `random.choice` of the dilation kernel in training."""
    hair = mask[:, 1:2]
    if training and opt.isTrain:
        if opt.random_expand_mask:
            mh = hair.shape[2]
            th = int(mh * opt.random_expand_th)
            th = th if th % 2 == 1 else th + 1
            k = rng_k if rng_k is not None else _pyrandom.choice([max(th - 4, 1), max(th - 2, 1), th, th + 2, th + 4])
            return 1 - F.max_pool2d(hair, kernel_size=k, stride=1, padding=int(k / 2)), k
        return mask[:, 0:1], None
    if opt.expand_mask_be:
        k = opt.expand_th
        p = int(k / 2)
        if opt.add_feat_zeros:
            th = opt.add_th
            H = W = opt.crop_size
            o = int(th / 2)
            e = hair * 0
            e[:, :, o:o + H, o:o + W] = F.max_pool2d(hair[:, :, o:o + H, o:o + W], kernel_size=k, stride=1, padding=p)
        else:
            e = F.max_pool2d(hair, kernel_size=k, stride=1, padding=p)
        return 1 - e, k
    return mask[:, 0:1], None


def background_encode2(image, mask, noise, sd, prefix, opt, rng_k=None):
    """BackgroundEncode2.forward for num_upsampling_layers != 'most' (encoder.py:284-341).  Note the
    reference switches on opt.isTrain, not module.training (encoder.py:286)."""
    back_mask, k = background_back_mask(mask, opt, opt.isTrain, rng_k)
    inp = noise if opt.random_noise_background else image * back_mask + noise * (1 - back_mask)
    x0 = conv_block(inp, sd, prefix + ".conv1", 7, 1, 3)
    x1 = conv_block(x0, sd, prefix + ".layer1", 4, 2, 1)
    x2 = conv_block(x1, sd, prefix + ".layer2", 4, 2, 1)
    x3 = conv_block(x2, sd, prefix + ".layer3", 4, 2, 1)
    sh, sw = back_mask.shape[2:]
    bm = [F.interpolate(back_mask, size=(int(sh / d), int(sw / d)), mode="nearest") for d in (2, 4, 8)]
    return [x3, x2, x1, x0], [bm[2], bm[1], bm[0], back_mask]


def latent_size(opt):
    """SPADEBGenerator.compute_latent_vector_size (generator.py:79-96)."""
    n_up = {"normal": 5, "more": 6, "most": 7}[opt.num_upsampling_layers]
    sw = (opt.crop_size + opt.add_th) // (2 ** n_up) if opt.add_feat_zeros else opt.crop_size // (2 ** n_up)
    sh = round(sw / opt.aspect_ratio)
    return sw, sh


def orient_channels(orient_mask, hair, opt):
    """generator.py:129-135 / pix2pix_model.py:548-553."""
    if opt.use_ig:
        return orient_mask
    th = orient_mask / 255.0 * math.pi
    return torch.cat([torch.sin(2 * th), torch.cos(2 * th)], dim=1) * hair


# ----------------------------------------------------------------------------------------------- generator
def generator_forward(sd, opt, input_ref, input_tag, image_ref, image_tag, orient_mask, noise, training,
                      rng_k=None, taps=None, **bn_kw):
    """SPADEBGenerator.forward, use_encoder + partialconv encoder + noise_background, 'more' upsampling
    (generator.py:107-230).  `taps` (dict) receives every block output for per-block parity checks."""
    sw, sh = latent_size(opt)
    ins_ref = input_ref[:, 1:2]
    ins_tag = input_tag[:, 1:2]
    x = image_encoder3(image_ref, ins_ref, ins_tag, sd, "fc", sh, sw)
    if taps is not None:
        taps["fc"] = x
    hair = input_tag[:, 1:2]
    seg = torch.cat([input_tag, orient_channels(orient_mask, hair, opt)], dim=1)
    back_feats, back_masks = background_encode2(image_tag, input_tag, noise, sd, "backgroud_enc", opt, rng_k)
    if taps is not None:
        for i, bfeat in enumerate(back_feats):
            taps["bg%d" % i] = bfeat
    H, W = hair.shape[2:]
    hair_masks = [F.interpolate(hair, size=(int(H / d), int(W / d)), mode="nearest") for d in (8, 4, 2)] + [hair]

    def up(t):
        return F.interpolate(t, scale_factor=2, mode="nearest")

    x = spade_resnet_block(x, seg, sd, "head_0", training, **bn_kw)
    if taps is not None:
        taps["head_0"] = x
    x = spade_resnet_block(up(x), seg, sd, "G_middle_0", training, **bn_kw)
    if taps is not None:
        taps["G_middle_0"] = x
    if opt.num_upsampling_layers in ("more", "most"):
        x = up(x)
    x = spade_resnet_block(x, seg, sd, "G_middle_1", training, **bn_kw)
    if taps is not None:
        taps["G_middle_1"] = x
    for i in range(4):
        x = spade_resnet_block(up(x), seg, sd, "up_%d" % i, training, **bn_kw)
        if taps is not None:
            taps["up_%d_pre" % i] = x
        if opt.bf_direct_add:
            x = back_feats[i] + x
        else:
            x = back_feats[i] * (1 - hair_masks[i]) + x * (1 - back_masks[i])
        if taps is not None:
            taps["up_%d" % i] = x
    x = F.conv2d(lrelu(x), sd["conv_img.weight"], sd["conv_img.bias"], padding=1)
    return torch.tanh(x)


def zeros_padding(t, th):
    """Pix2PixModel.zeros_padding (pix2pix_model.py:495-502)."""
    N, C, H, W = t.shape
    out = torch.zeros(N, C, H + th, W + th, dtype=t.dtype)
    o = int(th / 2)
    out[:, :, o:o + H, o:o + W] = t
    return out


def generate_fake(sdG, opt, data, training, **kw):
    """Pix2PixModel.generate_fake without VAE/blender (pix2pix_model.py:505-541); `data` holds the
    already-preprocessed tensors (one-hot input_ref/input_tag etc., pix2pix_model.py:209-254)."""
    t = dict(data)
    if opt.add_feat_zeros:
        for k in ("input_ref", "image_ref", "orient_mask", "input_tag", "image_tag", "noise"):
            t[k] = zeros_padding(t[k], opt.add_th)
    return generator_forward(sdG, opt, t["input_ref"], t["input_tag"], t["image_ref"], t["image_tag"], t["orient_mask"],
                             t["noise"], training, **kw)


def one_hot(label, nc=2):
    """preprocess_input's scatter_ (pix2pix_model.py:228-243)."""
    bs, _, h, w = label.shape
    return torch.zeros(bs, nc, h, w, dtype=torch.float32).scatter_(1, label.long(), 1.0)


# ----------------------------------------------------------------------------------------------- discriminator
def nlayer_discriminator(x, sd, prefix, training, n_layers=4):
    """NLayerDiscriminator.forward returning every intermediate (discriminator.py:74-120) with
    norm_D='spectralinstance' (normalization.py:18-54): model0 conv+bias+lrelu; model1..n-1 SN conv
    (bias stripped) + InstanceNorm2d(affine=False) + lrelu; last conv -> 1 channel."""
    outs = []
    x = lrelu(F.conv2d(x, sd[prefix + ".model0.0.weight"], sd[prefix + ".model0.0.bias"], stride=2, padding=2))
    outs.append(x)
    for n in range(1, n_layers):
        stride = 1 if n == n_layers - 1 else 2
        w = spectral_weight(sd, prefix + ".model%d.0.0" % n, training)
        x = lrelu(F.instance_norm(F.conv2d(x, w, None, stride=stride, padding=2)))
        outs.append(x)
    p = prefix + ".model%d.0" % n_layers
    x = F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], stride=1, padding=2)
    outs.append(x)
    return outs


def multiscale_discriminator(x, sd, opt, training):
    """MultiscaleDiscriminator.forward (discriminator.py:46-63)."""
    result = []
    for i in range(opt.num_D):
        result.append(nlayer_discriminator(x, sd, "discriminator_%d" % i, training, opt.n_layers_D))
        x = F.avg_pool2d(x, kernel_size=3, stride=2, padding=[1, 1], count_include_pad=False)
    return result


def discriminate(sdD, opt, input_tag, fake_image, real_image, orient_mask, training):
    """Pix2PixModel.discriminate + divide_pred (pix2pix_model.py:546-594)."""
    o = orient_channels(orient_mask, input_tag[:, 1:2], opt)
    fake_concat = torch.cat([input_tag, o, fake_image], dim=1)
    real_concat = torch.cat([input_tag, o, real_image], dim=1)
    out = multiscale_discriminator(torch.cat([fake_concat, real_concat], dim=0), sdD, opt, training)
    fake = [[t[: t.shape[0] // 2] for t in p] for p in out]
    real = [[t[t.shape[0] // 2:] for t in p] for p in out]
    return fake, real


# ----------------------------------------------------------------------------------------------- losses
def wide_edge_weight(inp, mask, wide_edge):
    """GANLoss.get_wide_edges/get_weight_mask (loss.py:60-78)."""
    n, c, h, w = inp.shape
    label = F.interpolate(mask, size=(h, w), mode="nearest")
    k = max(1, int(h * 0.06))
    p = int(k / 2)
    out = F.max_pool2d(label, kernel_size=k, stride=1, padding=p)
    out2 = 1 - F.max_pool2d(1 - label, kernel_size=k, stride=1, padding=p)
    edges = F.interpolate(out - out2, size=(h, w), mode="nearest")
    return edges * wide_edge + (1 - edges)


def gan_loss_hinge(preds, target_is_real, for_discriminator, label, opt):
    """GANLoss.__call__ / loss for gan_mode='hinge', remove_background=False (loss.py:80-140)."""
    total = 0
    for pred in preds:
        x = pred[-1] if isinstance(pred, list) else pred
        if for_discriminator:
            minval = torch.min((x - 1) if target_is_real else (-x - 1), torch.zeros_like(x))
            if opt.wide_edge > 1.0:
                minval = minval * wide_edge_weight(x, label, opt.wide_edge)
            loss = -torch.mean(minval)
        else:
            loss = -torch.mean(x)
        total = total + loss.view(1, -1).mean(dim=1)
    return total / len(preds)


def gan_feat_loss(pred_fake, pred_real, opt):
    """GANFeatLoss.forward, remove_background=False (loss.py:163-175)."""
    num_D = len(pred_fake)
    loss = torch.zeros(1, dtype=pred_fake[0][0].dtype)
    for i in range(num_D):
        for j in range(len(pred_fake[i]) - 1):
            loss = loss + F.l1_loss(pred_fake[i][j], pred_real[i][j].detach()) * opt.lambda_feat / num_D
    return loss


def compute_generator_loss(sdG, sdD, opt, data, **kw):
    """Pix2PixModel.compute_generator_loss with the in-scope losses (GAN hinge + GAN_Feat)
    (pix2pix_model.py:257-297); nets in train mode.  Returns (losses dict, fake image)."""
    fake = generate_fake(sdG, opt, data, True, **kw)
    pred_fake, pred_real = discriminate(sdD, opt, data["input_tag"], fake, data["image_tag"], data["orient_mask"], True)
    label_tag = data["input_tag"][:, 1:2]
    losses = {"GAN": gan_loss_hinge(pred_fake, True, False, label_tag, opt)}
    ref_tag_diff = torch.sum(data["input_tag"][:, 1] - data["input_ref"][:, 1])
    if not opt.no_ganFeat_loss and ref_tag_diff == 0:
        losses["GAN_Feat"] = gan_feat_loss(pred_fake, pred_real, opt)
    return losses, fake


def compute_discriminator_loss(sdG, sdD, opt, data, **kw):
    """Pix2PixModel.compute_discriminator_loss (pix2pix_model.py:367-398): G forward under no_grad but in
    train mode (BN batch stats, running-stat and u/v updates still happen), then hinge D losses."""
    with torch.no_grad():
        fake = generate_fake(sdG, opt, data, True, **kw)
    fake = fake.detach()
    pred_fake, pred_real = discriminate(sdD, opt, data["input_tag"], fake, data["image_tag"], data["orient_mask"], True)
    label_tag = data["input_tag"][:, 1:2]
    return {
        "D_Fake": gan_loss_hinge(pred_fake, False, True, label_tag, opt),
        "D_real": gan_loss_hinge(pred_real, True, True, label_tag, opt),
    }


def trainer_loss(losses):
    """Pix2PixTrainer: sum(losses.values()).mean() (pix2pix_trainer.py:42,66)."""
    return sum(losses.values()).mean()


# ----------------------------------------------------------------------------------------------- orientation inpainting
# SURVEY.md §8 row a16 ("next"): InpaintGenerator (generator.py:450-575) + Pix2PixModel.inpainting_orient
# (pix2pix_model.py:407-429).  The network is frozen and always runs in eval mode (pix2pix_model.py:196-198).
def spectral_weight_eval(sd, prefix, dim=0):
    """Eval-mode spectral norm: W_orig / (u^T W_mat v) with the stored u, v and no power iteration
    (torch/nn/utils/spectral_norm.py `compute_weight(do_power_iteration=False)`); `dim` = 1 for ConvTranspose2d."""
    w = sd[prefix + ".weight_orig"]
    u = sd[prefix + ".weight_u"]
    v = sd[prefix + ".weight_v"]
    mat = w if dim == 0 else w.transpose(0, dim)
    mat = mat.reshape(mat.shape[0], -1)
    sigma = torch.dot(u, torch.mv(mat, v))
    return w / sigma


def _in(x):
    return F.instance_norm(x, eps=1e-5)   # nn.InstanceNorm2d(dim): affine=False, no running stats


def inpaint_resnet_block(x, sd, prefix):
    """ResnetBlock (generator.py:450-465): x + IN(conv3x3(reflpad1(ReLU(IN(conv3x3 dil2(reflpad2(x)))))))."""
    h = F.conv2d(F.pad(x, (2, 2, 2, 2), mode="reflect"), spectral_weight_eval(sd, prefix + ".conv_block.1"),
                 sd[prefix + ".conv_block.1.bias"], dilation=2)
    h = F.relu(_in(h))
    h = F.conv2d(F.pad(h, (1, 1, 1, 1), mode="reflect"), spectral_weight_eval(sd, prefix + ".conv_block.5"),
                 sd[prefix + ".conv_block.5.bias"])
    return x + _in(h)


def inpaint_self_attention(x, sd, prefix):
    """SelfAttention (generator.py:468-487): softmax(q^T k) over keys, value projected back, concatenated to x."""
    n, c, a, b = x.shape
    q = F.conv2d(x, sd[prefix + ".query_conv.weight"], sd[prefix + ".query_conv.bias"]).view(n, -1, a * b).permute(0, 2, 1)
    k = F.conv2d(x, sd[prefix + ".key_conv.weight"], sd[prefix + ".key_conv.bias"]).view(n, -1, a * b)
    attn = torch.softmax(torch.bmm(q, k), dim=-1)
    v = F.conv2d(x, sd[prefix + ".value_conv.weight"], sd[prefix + ".value_conv.bias"]).view(n, -1, a * b)
    out = torch.bmm(v, attn.permute(0, 2, 1)).view(n, c, a, b)
    return torch.cat([x, out], dim=1)


def inpaint_generator(x, sd, blocks=12):
    """InpaintGenerator.forward, skips=False (generator.py:490-575): [N,4,H,W] -> [N,3,H,W] in [0,1]."""
    h = F.conv2d(F.pad(x, (3, 3, 3, 3), mode="reflect"), spectral_weight_eval(sd, "encoder.1"), sd["encoder.1.bias"])
    h = lrelu(_in(h))
    h = lrelu(_in(F.conv2d(h, spectral_weight_eval(sd, "encoder.4"), sd["encoder.4.bias"], stride=2, padding=1)))
    h = lrelu(_in(F.conv2d(h, spectral_weight_eval(sd, "encoder.7"), sd["encoder.7.bias"], stride=2, padding=1)))
    for i in range(blocks):
        h = inpaint_resnet_block(h, sd, "middle.%d" % i)
    h = inpaint_self_attention(h, sd, "middle.%d" % blocks)
    h = F.conv_transpose2d(h, spectral_weight_eval(sd, "decoder.0", dim=1), sd["decoder.0.bias"], stride=2, padding=1)
    h = F.relu(_in(h))
    h = F.conv_transpose2d(h, spectral_weight_eval(sd, "decoder.3", dim=1), sd["decoder.3.bias"], stride=2, padding=1)
    h = F.relu(_in(h))
    h = F.conv2d(F.pad(h, (3, 3, 3, 3), mode="reflect"), sd["decoder.7.weight"], sd["decoder.7.bias"])
    return (torch.tanh(h) + 1) / 2


def inpainting_orient(sd_ig, crop_size, hole, orient_rgb, noise, mask):
    """Pix2PixModel.inpainting_orient (pix2pix_model.py:407-429): fill the hole of the orientation RGB map with the
    frozen inpainting net (run at 256x256, nearest resize both ways) and convert it to the 2-channel orientation
    (cos 2theta, sin 2theta swapped into the generator's order) masked by the hair mask.  Returns (output, orient)."""
    inp = torch.cat([orient_rgb * (1 - hole) + noise * hole, hole], dim=1)
    if crop_size != 256:
        inp = F.interpolate(inp, size=(256, 256), mode="nearest")
    out = inpaint_generator(inp, sd_ig)
    if crop_size != 256:
        out = F.interpolate(out, size=(crop_size, crop_size), mode="nearest")
    out = out * hole + orient_rgb * (1 - hole)
    o2 = (out[:, :-1] - 0.5) * 2
    orient = torch.stack([o2[:, 1], o2[:, 0]], dim=1) * mask
    return out, orient


# ----------------------------------------------------------------------------------------------- Gabor orientation loss
# SURVEY.md §8f row 2 ("next"), first slice: L1OLoss with orient_filter='gabor' (models/networks/loss.py:214-240 gabor_fn,
# 274-318 calOrientationGabor, 329-372 forward).  The reference class hard-codes .cuda() (loss.py:218-232,284,297), so on the
# CPU-only build box it cannot be imported-and-run; tests/test_gpu_orient.py pins this restatement against the reference's own
# class on the GPU box (baseline/_ref).
def gabor_kernels(num_kernels=32, kernel_size=17):
    r = kernel_size // 2
    y = torch.arange(-r, r + 1).view(1, -1).repeat(kernel_size, 1).float()
    x = torch.arange(-r, r + 1).view(-1, 1).repeat(1, kernel_size).float()
    ks = []
    for k in range(num_kernels):
        theta = torch.ones(1) * (math.pi * k / num_kernels)
        x_t = x * torch.cos(theta) + y * torch.sin(theta)
        y_t = -x * torch.sin(theta) + y * torch.cos(theta)
        ks.append(torch.exp(-.5 * (x_t ** 2 / 2.0 ** 2 + y_t ** 2 / 3.0 ** 2)) * torch.cos(2 * math.pi / 4.0 * x_t))
    return torch.stack(ks, 0).unsqueeze(1)          # [32,1,17,17]


def orient_loss_gabor(fake_image0, orientation_label0, input_semantics, use_ig=False):
    """-> (orient_loss, confidence_loss), loss.py:329-372 with 'gabor' in opt.orient_filter."""
    hair = input_semantics[:, 1:2]
    fake = (fake_image0 + 1) / 2.0 * 255
    gray = (0.299 * fake[:, 0] + 0.587 * fake[:, 1] + 0.144 * fake[:, 2]).unsqueeze(1)
    res = F.conv2d(gray, gabor_kernels().to(gray.dtype), stride=1, padding=8)
    res = res * (res >= 0).to(res.dtype)                      # resTensor[resTensor < 0] = 0 (in place: zero gradient there)
    max_idx = torch.argmax(res, dim=1).float()
    conf = (torch.tanh(torch.max(res, dim=1)[0]) + 1) / 2.0
    conf = conf.unsqueeze(1)
    ang = (max_idx * math.pi / 32).unsqueeze(1)
    two = torch.cat([torch.sin(2 * ang), torch.cos(2 * ang)], dim=1) * conf
    if not use_ig:
        lab = orientation_label0 / 255 * math.pi
        lab2 = torch.cat([torch.sin(2 * lab), torch.cos(2 * lab)], dim=1)
    else:
        lab2 = orientation_label0
    orient_loss = F.l1_loss(two * hair, (lab2 * hair).detach())
    confc = torch.clamp(conf, 0.001, 1)
    confidence_loss = -torch.sum(torch.log(confc) * hair) / torch.sum(hair)
    return orient_loss, confidence_loss
