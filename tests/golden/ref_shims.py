"""Import the UNMODIFIED reference (tzt101/MichiGAN at /root/reference) on CPU in the build container.

Only used by make_golden.py (fixture generation) and by tests that are skipped when /root/reference
is absent (it does not exist on the GPU box).  The shims touch no arithmetic:
  1. stub `matplotlib` / `matplotlib.pyplot` (imported, unused: generator.py:9, normalization.py:13);
  2. stub `dominate` (util/html.py:7-8, pulled in by util/visualizer.py);
  3. `--gpu_ids -1` (CPU);
  4. training only: Adam betas as floats (pix2pix_model.py:141 passes the int 0, rejected by torch 2.11);
  5. training only: `networks.StyleContentLoss` replaced by a zero stub *before* Pix2PixModel is built
     (it would download VGG19 and call .cuda(), loss.py:659), plus --no_vgg_loss --no_orient_loss
     --no_lab_loss (loss classes that hard-code .cuda() / break on modern torch).
"""
import os
import sys
import types

REF = os.environ.get("MICHIGAN_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF, "models", "networks"))


def import_reference():
    if not available():
        raise RuntimeError("reference not found at %s" % REF)
    for name in ("matplotlib", "matplotlib.pyplot", "dominate", "dominate.tags"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            sys.modules[name] = m
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["dominate"].tags = sys.modules["dominate.tags"]
    sys.modules["dominate"].document = object
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import warnings
    warnings.filterwarnings("ignore", category=SyntaxWarning)
    import models.networks as networks  # noqa: F401
    return networks


TRAIN_FLAGS = ("--use_encoder --wide_edge 2 --noise_background --random_expand_mask --no_confidence_loss --no_style_loss "
               "--no_rgb_loss --no_content_loss --no_background_loss --no_vgg_loss --no_orient_loss --no_lab_loss "
               "--gpu_ids -1 --checkpoints_dir /tmp/mg_ref_ckpt --no_html").split()
TEST_FLAGS = ("--use_encoder --noise_background --expand_mask_be --expand_th 5 --gpu_ids -1 "
              "--checkpoints_dir /tmp/mg_ref_ckpt").split()


def ref_options(train=True, extra=()):
    """The reference's own option parser (options/base_options.py:212-242) on a patched argv."""
    import_reference()
    import contextlib
    import io
    argv = sys.argv
    sys.argv = ["x"] + (TRAIN_FLAGS if train else TEST_FLAGS) + list(extra)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            if train:
                from options.train_options import TrainOptions
                opt = TrainOptions().parse()
            else:
                from options.test_options import TestOptions
                opt = TestOptions().parse()
    finally:
        sys.argv = argv
    return opt


def patch_training():
    """Shims 4 and 5."""
    import torch
    networks = import_reference()

    class _ZeroStyleContent(torch.nn.Module):
        def __init__(self, opt=None):
            super().__init__()

        def forward(self, *a, **k):
            return 0, 0

    networks.StyleContentLoss = _ZeroStyleContent
    if not getattr(torch.optim.Adam, "_mg_patched", False):
        orig = torch.optim.Adam.__init__

        def init(self, params, lr=1e-3, betas=(0.9, 0.999), **kw):
            orig(self, params, lr=lr, betas=(float(betas[0]), float(betas[1])), **kw)

        torch.optim.Adam.__init__ = init
        torch.optim.Adam._mg_patched = True
