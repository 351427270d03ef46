"""Tensor-core operand policy.

  "mixed16" (default)  SPADE gamma/beta GEMMs (66 % of the generator FLOPs): fp16 operands, one pass -
                       same 11-bit significand as TF32 at twice the issue rate; their inputs (ReLU of
                       a 3x3 conv of a [-1,1] segmap) and weights are far inside the fp16 range.
                       conv_0 / conv_1 / conv_s, background- and reference-encoder convs: bf16 hi+lo
                       split of both operands, three passes (A_hi*W_hi + A_lo*W_hi + A_hi*W_lo, fp32
                       accumulate) = ~16 significand bits with bf16's full exponent range.  These convs
                       feed batch-norm statistics and dominate the output error in TF32.
                       Discriminator convs: fp16, one pass.
  "tf32"               every tensor-core conv reads fp32 storage as TF32, one pass.

Set with MICHIGAN_B200_PRECISION or precision.set_mode().
"""
import os

from . import ops

_mode = os.environ.get("MICHIGAN_B200_PRECISION", "mixed16")


def set_mode(mode):
    global _mode
    if mode not in ("mixed16", "tf32"):
        raise ValueError("precision mode must be 'mixed16' or 'tf32'")
    _mode = mode


def mode():
    return _mode


def gb_fmt(cin=128):
    """Operand format of the SPADE gamma/beta GEMM and of the discriminator convs.  16-bit operands need
    Cin % 64 == 0 (one 128 B swizzle row = 64 channels); other layers (ngf/ndf = 32 test nets) use TF32."""
    return ops.F16 if (_mode == "mixed16" and cin % 64 == 0) else ops.TF32


def gb_policy(ratio, training=True):
    """(fmt, split) of the SPADE gamma/beta GEMM of a block whose feature map is 1/`ratio` of the segmap resolution.

    Train-mode statistics (the benchmarked forward, every training iteration): the low-resolution blocks (head_0,
    G_middle_0/1, up_0: ratio >= 8, 6 % of the gamma/beta FLOPs) feed every later batch-norm, so their GEMMs use the three-pass
    bf16 split, the others one-pass fp16; emulation on the 512x512 net: image max-abs error 1.0e-3 -> 4.9e-4.
    The criterion is the block's DEPTH (resolution ratio), not its absolute height: with --add_feat_zeros the same blocks run
    at 9..72 instead of 8..64 pixels (keying on h <= 64 left up_0 in one-pass fp16 at 576x576: 2.5e-3).

    Eval mode (running statistics, inference.py): batch-norm no longer re-normalises each layer's output, so the operand rounding
    of the one-pass GEMMs propagates instead of being absorbed by the next layer's statistics - measured on the 576x576
    inference geometry 0.9e-3 .. 1.1e-3 max-abs (mean 2e-5) with the training policy.  Every gamma/beta GEMM therefore uses the
    three-pass split in eval mode (single-image inference is latency-, not throughput-bound)."""
    if _mode != "mixed16":
        return ops.TF32, False
    if not training:
        return ops.BF16, True
    return (ops.BF16, True) if ratio >= 8 else (ops.F16, False)


def grad_fmt():
    """Operand format of the SPADE gamma|beta GRADIENT GEMMs (data gradient w.r.t. actv and weight gradient: 2/3 of the backward
    FLOPs): bf16 operands, fp32 accumulation.  They only feed the mlp_shared / mlp_gamma / mlp_beta parameter gradients, i.e. sums
    over 10^5..10^6 pixels in which bf16's 2^-9 operand rounding averages out (measured gradient cosine vs the fp32 reference in
    tests/test_gpu_parity.py); dgamma|dbeta is produced directly in bf16 by mg_spade_bwd.  TF32 in "tf32" mode."""
    return ops.BF16 if _mode == "mixed16" else ops.TF32


def conv_grad_fmt():
    """Operand format of the gradient GEMMs of conv_0 / conv_1 / conv_s (data and weight gradients): bf16 operands, fp32
    accumulation (dY converted once per layer, the forward's bf16 hi operand reused as the weight-gradient input); measured
    against the reference trainer's gradients: worst per-tensor cosine 0.9991 with bf16 vs 0.9989 with TF32 (the difference to
    the reference is dominated by the piecewise-linear losses, not by operand rounding) and -3.5 % step time.
    MICHIGAN_B200_GRAD16=gb keeps these GEMMs in TF32."""
    import os
    return ops.BF16 if (_mode == "mixed16" and os.environ.get("MICHIGAN_B200_GRAD16", "all") == "all") else ops.TF32


def conv_fmt(cin):
    """Operand format of the split-precision convs (BF16 -> three passes) or TF32 (one pass)."""
    return ops.BF16 if (_mode == "mixed16" and cin % 64 == 0) else ops.TF32


def out_spec(fmt, split):
    """(kwargs for a producer kernel, extractor) so that its output is a tensor-core operand of `fmt`."""
    if fmt == ops.TF32:
        return dict(round_out=True), (lambda r: (ops.TF32, r, None))
    return dict(out16=(fmt, split), want_f32=False), (lambda r: (fmt, r[1], r[2]))


def pack_conv(w, inv_sigma, fmt):
    if fmt == ops.TF32:
        return ops.pack_weight(w, inv_sigma, True)
    return ops.pack_weight16(w, inv_sigma, fmt, split=(fmt == ops.BF16))


def conv(operand, wpack, cout, kh, kw, stride, pad, **kw_):
    fmt, hi, lo = operand
    return ops.conv_igemm(hi, wpack, cout, kh, kw, stride, pad, a_fmt=fmt, x_lo=lo, **kw_)
