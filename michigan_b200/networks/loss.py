"""Hinge GAN loss and GAN feature-matching loss with the reference's call signatures
(models/networks/loss.py:19-140, 144-175; the in-scope modes: gan_mode='hinge',
remove_background=False).  These are small reductions over the discriminator's outputs."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class GANLoss(nn.Module):
    def __init__(self, gan_mode, target_real_label=1.0, target_fake_label=0.0, tensor=torch.FloatTensor, opt=None):
        super().__init__()
        if gan_mode != "hinge":
            raise NotImplementedError("michigan_b200: gan_mode '%s' (only 'hinge' is on the hot path)" % gan_mode)
        if getattr(opt, "remove_background", False):
            raise NotImplementedError("michigan_b200: --remove_background is not implemented")
        self.gan_mode = gan_mode
        self.opt = opt

    def get_wide_edges(self, t, th=0.06):
        n, c, h, w = t.size()
        k = max(1, int(h * th))
        p = int(k / 2)
        out = F.max_pool2d(t, kernel_size=k, stride=1, padding=p)
        out2 = 1 - F.max_pool2d(1 - t, kernel_size=k, stride=1, padding=p)
        return F.interpolate(out - out2, size=(h, w), mode="nearest")

    def get_weight_mask(self, input, mask):
        n, c, h, w = input.size()
        label = F.interpolate(mask, size=(h, w), mode="nearest")
        edges = self.get_wide_edges(label)
        return edges * self.opt.wide_edge + (1 - edges)

    def loss(self, input, target_is_real, for_discriminator=True, label=None):
        if for_discriminator:
            minval = torch.clamp((input - 1) if target_is_real else (-input - 1), max=0)
            if self.opt.wide_edge > 1.0:
                minval = minval * self.get_weight_mask(input, label)
            return -torch.mean(minval)
        assert target_is_real, "The generator's hinge loss must be aiming for real"
        return -torch.mean(input)

    def __call__(self, input, target_is_real, for_discriminator=True, label=None):
        if isinstance(input, list):
            loss = 0
            for pred_i in input:
                if isinstance(pred_i, list):
                    pred_i = pred_i[-1]
                loss_tensor = self.loss(pred_i, target_is_real, for_discriminator, label.detach())
                bs = 1 if len(loss_tensor.size()) == 0 else loss_tensor.size(0)
                loss = loss + torch.mean(loss_tensor.view(bs, -1), dim=1)
            return loss / len(input)
        return self.loss(input, target_is_real, for_discriminator, label.detach())


class GANFeatLoss(nn.Module):
    def __init__(self, opt=None):
        super().__init__()
        self.opt = opt

    def forward(self, pred_fake, pred_real, label=None):
        num_D = len(pred_fake)
        total = torch.zeros(1, device=pred_fake[0][0].device)
        for i in range(num_D):
            for j in range(len(pred_fake[i]) - 1):
                total = total + F.l1_loss(pred_fake[i][j], pred_real[i][j].detach()) * self.opt.lambda_feat / num_D
        return total
