"""Drop the B200 operator classes into an (unmodified) MichiGAN checkout.

The reference resolves networks by *name* at run time: `find_class_in_module(name + 'generator',
'models.networks.generator')` scans the module's __dict__ (util/util.py:180-192,
models/networks/__init__.py:16-24), and the trainer imports `DataParallelWithCallback` from
`models.networks.sync_batchnorm` (trainers/pix2pix_trainer.py:6).  Rebinding those names before the
options are parsed is therefore enough for `train.py` / `inference.py` to run on the CUDA kernels
with no source change (see INTEGRATION.md for the launcher).
"""
import importlib
import os
import sys


def install(reference_root=None):
    """Import the reference's `models.networks` package and rebind the hot-path classes."""
    if reference_root:
        reference_root = os.path.abspath(reference_root)
        if reference_root not in sys.path:
            sys.path.insert(0, reference_root)
    from . import networks as mine
    gen = importlib.import_module("models.networks.generator")
    dis = importlib.import_module("models.networks.discriminator")
    nrm = importlib.import_module("models.networks.normalization")
    arc = importlib.import_module("models.networks.architecture")
    enc = importlib.import_module("models.networks.encoder")
    sbn = importlib.import_module("models.networks.sync_batchnorm")
    base = importlib.import_module("models.networks.base_network")
    pkg = importlib.import_module("models.networks")
    # the factory asserts issubclass(cls, <reference BaseNetwork>) (models/networks/__init__.py:21-22):
    # derive the installed network classes from both bases (same names, lookup is by module attribute)
    def both(cls):
        if issubclass(cls, base.BaseNetwork):
            return cls
        return type(cls.__name__, (cls, base.BaseNetwork), {"__doc__": cls.__doc__, "__module__": cls.__module__})

    G = both(mine.SPADEBGenerator)
    MSD = both(mine.MultiscaleDiscriminator)
    NLD = both(mine.NLayerDiscriminator)
    gen.SPADEBGenerator = G
    gen.SPADEResnetBlock = mine.SPADEResnetBlock
    gen.ImageEncoder3 = mine.ImageEncoder3
    gen.BackgroundEncode2 = mine.BackgroundEncode2
    dis.MultiscaleDiscriminator = MSD
    dis.NLayerDiscriminator = NLD
    nrm.SPADE = mine.SPADE
    nrm.SynchronizedBatchNorm2d = mine.SynchronizedBatchNorm2d
    arc.SPADEResnetBlock = mine.SPADEResnetBlock
    arc.SPADE = mine.SPADE
    enc.ImageEncoder3 = mine.ImageEncoder3
    enc.BackgroundEncode2 = mine.BackgroundEncode2
    for name in ("SynchronizedBatchNorm1d", "SynchronizedBatchNorm2d", "SynchronizedBatchNorm3d", "DataParallelWithCallback",
                 "patch_replication_callback", "convert_model", "patch_sync_batchnorm"):
        setattr(sbn, name, getattr(mine.sync_batchnorm, name))
    pkg.SPADEBGenerator, pkg.MultiscaleDiscriminator, pkg.NLayerDiscriminator = G, MSD, NLD
    return mine
