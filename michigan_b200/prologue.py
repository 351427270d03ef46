"""GPU input prologue: batch-wise synthesis of the generator's auxiliary inputs on the device, replacing
the per-sample CPU work of the reference's Dataset.__getitem__ (data/pix2pix_dataset.py:66-200 -> data/base_dataset.py:335-396):
background noise, orientation-RGB map and the random hole of --use_ig.  Random draws come from torch's device generator
(Philox) - the same distributions as the reference's numpy / `random` draws, not the same streams; the arithmetic applied
to the draws is checked against the reference functions on identical draws (tests/test_gpu_kernels.py)."""
import ctypes as C

import torch

from . import _lib, ops


def noise_octave_sizes(h, w):
    """Octave shapes of generate_noise (base_dataset.py:387-396): halve while both sides are >= 8."""
    sizes = []
    while w >= 8 and h >= 8:
        sizes.append((h, w))
        w //= 2
        h //= 2
    return sizes


def noise_from_fields(fields, n, h, w):
    """fields: list of [n, h>>l, w>>l, 3] float32 CUDA tensors (draws of N(0.5, 0.25^2)) -> [n,3,h,w] noise image."""
    for l, f in enumerate(fields):
        ops._chk(f, "field%d" % l)
        if tuple(f.shape) != (n, h >> l, w >> l, 3):
            raise ValueError("octave %d must be [%d,%d,%d,3], got %s" % (l, n, h >> l, w >> l, tuple(f.shape)))
    out = torch.empty((n, 3, h, w), device=fields[0].device, dtype=torch.float32)
    ptrs = (C.c_void_p * len(fields))(*[f.data_ptr() for f in fields])
    _lib.check(_lib.load().mg_noise_pyramid(ptrs, len(fields), out.data_ptr(), n, h, w, ops._stream()), "mg_noise_pyramid")
    return out


def generate_noise(n, h, w, device, generator=None):
    """Batch version of base_dataset.generate_noise: multi-octave Gaussian noise, mean 0.5."""
    fields = [torch.randn((n, hh, ww, 3), device=device, generator=generator) * 0.25 + 0.5 for hh, ww in noise_octave_sizes(h, w)]
    return noise_from_fields(fields, n, h, w)


def orient_rgb(orient, label):
    """orient [N,1,H,W] (0..255), label [N,1,H,W] {0,1} -> [N,3,H,W] (trans_orient_to_rgb + ToTensor + mask)."""
    ops._chk(orient, "orient"); ops._chk(label, "label")
    n, _, h, w = orient.shape
    out = torch.empty((n, 3, h, w), device=orient.device, dtype=torch.float32)
    _lib.check(_lib.load().mg_orient_rgb(orient.data_ptr(), label.data_ptr(), out.data_ptr(), n, h, w, ops._stream()), "mg_orient_rgb")
    return out


def hole_mask(mask, orient_mask, th_u=None, idx_u=None, generator=None):
    """generate_hole for a batch: mask, orient_mask [N,1,H,W]; th_u ~ U(0.5,1.2), idx_u ~ U[0,1) drawn here unless given."""
    ops._chk(mask, "mask"); ops._chk(orient_mask, "orient_mask")
    n, _, h, w = mask.shape
    dev = mask.device
    if th_u is None:
        th_u = torch.rand(n, device=dev, generator=generator) * 0.7 + 0.5
    if idx_u is None:
        idx_u = torch.rand(n, device=dev, generator=generator)
    ops._chk(th_u, "th_u"); ops._chk(idx_u, "idx_u")
    out = torch.empty_like(mask)
    _lib.check(_lib.load().mg_hole_mask(mask.data_ptr(), orient_mask.data_ptr(), th_u.data_ptr(), idx_u.data_ptr(), out.data_ptr(), n, h, w,
                                        ops._stream()), "mg_hole_mask")
    return out
