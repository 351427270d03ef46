"""michigan_b200 — a B200-native (sm_100a) implementation of MichiGAN's data-parallel hot path:
the SPADE-conditioned generator (`--netG spadeb`), the multiscale PatchGAN discriminator and
synchronized batch-norm, behind the reference's operator surface and checkpoint layout.

    python -m michigan_b200.build          # nvcc -> michigan_b200/lib/libmichigan_sm100.so (C ABI)
    michigan_b200.install(reference_root)  # plug the networks into the reference's train.py/inference.py
"""
__version__ = "0.1.0"


def install(reference_root=None):
    from .install import install as _install
    return _install(reference_root)
