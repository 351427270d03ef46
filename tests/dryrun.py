"""Test infrastructure: run the Python glue of the CUDA path on the CPU with a no-op library.

`dry_run()` replaces the ctypes library object by one whose every `mg_*` entry point returns 0 without doing anything,
and lets `ops` accept CPU tensors.  Outputs therefore hold whatever `torch.empty/zeros` returned - the VALUES are
meaningless - but every wrapper still allocates its results with the real shapes and dtypes, so a whole forward/backward
of the generator and discriminator exercises the host code (argument plumbing, saved-state field names, gradient
bookkeeping, shapes) on the GPU-less build box before GPU minutes are spent.  This is not a CPU path of the product:
nothing outside tests/ can enable it.
"""
import contextlib
from types import SimpleNamespace

import torch


class _NoopLib:
    def __init__(self, real):
        self._real = real
        self.calls = []

    def __getattr__(self, name):
        if name in ("mg_last_error", "mg_version", "mg_launch_count", "mg_get_tuning", "mg_set_tuning", "mg_loss_term_bytes",
                    "mg_peer_buffer_bytes", "mg_peer_max_elems"):    # host-only queries
            return getattr(self._real, name)
        if not name.startswith("mg_"):
            raise AttributeError(name)
        getattr(self._real, name)   # AttributeError for an entry point the real library does not export

        def call(*args):
            self.calls.append(name)
            return 0
        return call


@contextlib.contextmanager
def dry_run():
    from michigan_b200 import _lib, ops
    real = _lib.load()
    fake = _NoopLib(real)
    saved = (_lib._lib, ops._chk, ops._stream)

    def chk(t, name, dtype=torch.float32):
        if t is None:
            return
        if t.dtype != dtype:
            raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
        if not t.is_contiguous():
            raise ValueError("%s must be contiguous" % name)

    _lib._lib, ops._chk, ops._stream = fake, chk, (lambda: 0)
    # deterministic, finite "results": empty() -> zeros() so that nothing downstream sees NaN garbage
    real_empty, real_empty_like = torch.empty, torch.empty_like
    torch.empty = lambda *a, **k: torch.zeros(*a, **k)
    torch.empty_like = lambda *a, **k: torch.zeros_like(*a, **k)
    # the stand-alone Pix2PixModel insists on a CUDA device: keep everything on the CPU for the dry run
    from michigan_b200 import networks, pix2pix_model
    saved_dev, saved_create = pix2pix_model._device, networks.create_network
    pix2pix_model._device = lambda: torch.device("cpu")

    def create_network(cls, opt):
        net = cls(opt)
        net.print_network()
        net.init_weights(opt.init_type, opt.init_variance)
        return net

    networks.create_network = create_network
    try:
        yield SimpleNamespace(lib=fake)
    finally:
        _lib._lib, ops._chk, ops._stream = saved
        torch.empty, torch.empty_like = real_empty, real_empty_like
        pix2pix_model._device, networks.create_network = saved_dev, saved_create
