"""SPADE and the non-SPADE norm wrapper — same classes/constructors/state-dict keys as the
reference's models/networks/normalization.py:18-118; the arithmetic runs in the fused CUDA kernels
(see architecture.py / discriminator.py)."""
import re

import torch.nn as nn
import torch.nn.utils.spectral_norm as spectral_norm

from .sync_batchnorm import SynchronizedBatchNorm2d


def get_nonspade_norm_layer(opt, norm_type="instance"):
    """normalization.py:18-54.  Only 'spectralinstance' / 'instance' / 'none' variants have kernels."""
    def get_out_channel(layer):
        return getattr(layer, "out_channels", None) or layer.weight.size(0)

    def add_norm_layer(layer):
        nonlocal norm_type
        if norm_type.startswith("spectral"):
            layer = spectral_norm(layer)
            subnorm_type = norm_type[len("spectral"):]
        else:
            subnorm_type = norm_type
        if subnorm_type == "none" or len(subnorm_type) == 0:
            return layer
        if getattr(layer, "bias", None) is not None:
            delattr(layer, "bias")
            layer.register_parameter("bias", None)
        if subnorm_type == "instance":
            norm_layer = nn.InstanceNorm2d(get_out_channel(layer), affine=False)
        else:
            raise NotImplementedError("michigan_b200: norm_D/E '%s' has no CUDA kernel (supported: [spectral]instance)"
                                      % subnorm_type)
        return nn.Sequential(layer, norm_layer)

    return add_norm_layer


class SPADE(nn.Module):
    """normalization.py:70-118.  Holds `param_free_norm`, `mlp_shared.0`, `mlp_gamma`, `mlp_beta`; the
    forward (BN -> seg conv -> gamma/beta convs -> modulation) is fused into the consumer conv by
    SPADEResnetBlock."""

    def __init__(self, config_text, norm_nc, label_nc, use_weight_norm=False):
        super().__init__()
        assert config_text.startswith("spade")
        parsed = re.search(r"spade(\D+)(\d)x\d", config_text)
        param_free_norm_type = str(parsed.group(1))
        ks = int(parsed.group(2))
        if param_free_norm_type != "syncbatch" or ks != 3 or use_weight_norm:
            raise NotImplementedError("michigan_b200 implements norm_G=spectralspadesyncbatch3x3 (the netG=spadeb default, "
                                      "generator.py:21-24); got '%s'" % config_text)
        if label_nc != 4:
            raise NotImplementedError("michigan_b200: SPADE segmap must have 4 channels (2 label + 2 orientation); got %d"
                                      % label_nc)
        self.param_free_norm = SynchronizedBatchNorm2d(norm_nc, affine=False)
        nhidden = 128
        pw = ks // 2
        self.mlp_shared = nn.Sequential(nn.Conv2d(label_nc, nhidden, kernel_size=ks, padding=pw), nn.ReLU())
        self.mlp_gamma = nn.Conv2d(nhidden, norm_nc, kernel_size=ks, padding=pw)
        self.mlp_beta = nn.Conv2d(nhidden, norm_nc, kernel_size=ks, padding=pw)
        self.use_weight_norm = use_weight_norm
        self.norm_nc = norm_nc

    def forward(self, x, segmap):
        raise RuntimeError("SPADE is fused into SPADEResnetBlock's kernels; call the block")
