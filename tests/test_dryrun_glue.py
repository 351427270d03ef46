"""Host-code smoke tests on the CPU-only build box: the whole Python glue of the CUDA path (forward in both modes, the
hand-written backward, the trainer iteration) runs against a no-op library (tests/dryrun.py).  Values are meaningless;
what is checked is that the code executes, calls only exported entry points, and produces tensors / gradients of the
right shapes.  Numerics are the `-m gpu` tests' job."""
import random

import pytest
import torch

from dryrun import dry_run


def _nets(size=64, ngf=64):
    from michigan_b200 import networks
    from michigan_b200.options import make_opt
    opt = make_opt(is_train=True, ngf=ngf, ndf=64, crop_size=size, gpu_ids=[])
    G = networks.SPADEBGenerator(opt).train()
    D = networks.MultiscaleDiscriminator(opt).train()
    return opt, G, D


def _inputs(n, size):
    from helpers import preprocessed
    _, pre = preprocessed(dict(batch=n, size=size, data_seed=1))
    return pre


@pytest.mark.parametrize("mode", ["mixed16", "tf32"])
def test_generator_and_discriminator_forward_backward_glue(mode):
    from michigan_b200 import precision
    old = precision.mode()
    precision.set_mode(mode)
    try:
        opt, G, D = _nets()
        pre = _inputs(2, 64)
        random.seed(0)
        with dry_run() as dr:
            with torch.no_grad():
                out = G(pre["input_ref"], orient_mask=pre["orient_mask"], image_ref=pre["image_ref"], input_tag=pre["input_tag"],
                        noise=pre["noise"], image_tag=pre["image_tag"])
            assert tuple(out.shape) == (2, 3, 64, 64)
            n_nograd = len(dr.lib.calls)
            fake = G(pre["input_ref"], orient_mask=pre["orient_mask"], image_ref=pre["image_ref"], input_tag=pre["input_tag"],
                     noise=pre["noise"], image_tag=pre["image_tag"])
            assert fake.requires_grad and tuple(fake.shape) == (2, 3, 64, 64)
            cond = torch.zeros(2, 4, 64, 64)
            x = torch.cat([torch.cat([cond, fake], 1), torch.cat([cond, pre["image_tag"]], 1)], 0)
            outs = D(x)
            assert len(outs) == 2 and all(len(o) == 5 for o in outs)
            loss = sum((t * t).mean() for o in outs for t in o) + fake.mean()
            loss.backward()
            assert len(dr.lib.calls) > 3 * n_nograd
        for name, p in list(G.named_parameters()) + list(D.named_parameters()):
            if name.startswith("backgroud_enc.layer4"):
                assert p.grad is None      # present in the state dict, never executed (encoder.py:284)
            else:
                assert p.grad is not None and p.grad.shape == p.shape, name
        # eval mode forward (running statistics, stored u/v) and the reference's NCHW block signature
        with dry_run():
            with torch.no_grad():
                G.eval()
                G(pre["input_ref"], orient_mask=pre["orient_mask"], image_ref=pre["image_ref"], input_tag=pre["input_tag"],
                  noise=pre["noise"], image_tag=pre["image_tag"])
                D.eval()
                D(x.detach())
    finally:
        precision.set_mode(old)


def test_discriminator_only_backward_glue():
    """D step: gradients for D's parameters only, input without grad; G step through a frozen D: input gradient only."""
    opt, G, D = _nets()
    x = torch.rand(4, 7, 64, 64)
    with dry_run():
        outs = D(x)
        sum(t.mean() for o in outs for t in o).backward()
        assert all(p.grad is not None for p in D.parameters())
        for p in D.parameters():
            p.grad = None
        D.requires_grad_(False)
        xg = x.clone().requires_grad_()
        outs = D(xg)
        sum(t.mean() for o in outs for t in o).backward()
        assert xg.grad is not None and xg.grad.shape == xg.shape
        assert all(p.grad is None for p in D.parameters())


def test_standalone_model_train_iteration_glue(tmp_path):
    """michigan_b200.Pix2PixModel + train_iteration (what bench.py's train-step leg and the NCCL parity worker drive):
    generator step with the discriminator frozen, discriminator step, checkpoint layout; loss flags must be explicit."""
    import os
    from michigan_b200 import checkpoint
    from michigan_b200.options import make_opt
    from michigan_b200.pix2pix_model import Pix2PixModel, train_iteration
    from michigan_b200.synth import synthetic_batch
    with dry_run():
        opt = make_opt(is_train=True, ngf=64, ndf=64, crop_size=64, batchSize=2, checkpoints_dir=str(tmp_path), name="dry")
        with pytest.raises(NotImplementedError):
            Pix2PixModel(make_opt(is_train=True, ngf=64, ndf=64, crop_size=64, no_vgg_loss=False))
        model = Pix2PixModel(opt).train()
        optG, optD = model.create_optimizers(opt)
        data = synthetic_batch(2, 64, 1)
        random.seed(0)
        g, d, img = train_iteration(model, optG, optD, dict(data))
        assert set(g) == {"GAN", "GAN_Feat"} and set(d) == {"D_Fake", "D_real"} and tuple(img.shape) == (2, 3, 64, 64)
        assert all(p.requires_grad for p in model.netD.parameters())       # un-frozen again after the generator step
        model.save("latest")
        checkpoint.wait_pending()
        sd = torch.load(os.path.join(tmp_path, "dry", "latest_net_G.pth"))
        assert list(sd.keys()) == list(model.netG.state_dict().keys())
        out = model(dict(data), mode="inference")
        assert tuple(out.shape) == (2, 3, 64, 64) and not out.requires_grad


def test_use_ig_and_orientation_loss_glue(tmp_path):
    """BASELINE.json configs[4]'s model: --use_ig (frozen InpaintGenerator loaded from <checkpoints_dir>/<name>/InpaintingModel_gen.pth,
    2-channel orientation input derived from its output) plus the Gabor orientation / confidence losses."""
    import os
    from michigan_b200.networks import InpaintGenerator
    from michigan_b200.options import make_opt
    from michigan_b200.pix2pix_model import Pix2PixModel, train_iteration
    from michigan_b200.synth import fill_state_dict, synthetic_batch
    os.makedirs(tmp_path / "ig")
    ig = InpaintGenerator()
    fill_state_dict(ig.state_dict(), 2)
    torch.save({"generator": ig.state_dict()}, tmp_path / "ig" / "InpaintingModel_gen.pth")
    with dry_run():
        opt = make_opt(is_train=True, ngf=64, ndf=64, crop_size=256, batchSize=1, use_ig=True, checkpoints_dir=str(tmp_path), name="ig",
                       ig_model_name="InpaintingModel_gen.pth", netIG="inpaint", no_orient_loss=False, no_confidence_loss=False)
        model = Pix2PixModel(opt).train()
        assert not model.netIG.training and model.netG.training
        optG, optD = model.create_optimizers(opt)
        assert len(list(optG.param_groups[0]["params"])) == len(list(model.netG.parameters()))      # the frozen net is not optimised
        data = synthetic_batch(1, 256, 3, use_ig=True)
        random.seed(0)
        g, d, img = train_iteration(model, optG, optD, dict(data))
        assert set(g) == {"GAN", "GAN_Feat", "ORIENT", "CONFIDENCE"} and tuple(img.shape) == (1, 3, 256, 256)
