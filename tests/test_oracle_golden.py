"""The CPU oracle against the golden fixtures produced by the live reference
(tests/golden/make_golden.py).  Runs without the reference and without a GPU."""
import numpy as np
import torch

import michigan_oracle as orc
from helpers import assert_summary_close, load_golden, preprocessed, reference_layout_state

TOL = 3e-5


def _close(name, got, ref, tol=TOL):
    ref = torch.from_numpy(np.asarray(ref))
    err = (got.detach() - ref).abs().max().item()
    assert err <= tol * max(1.0, ref.abs().max().item()), "%s: max err %g" % (name, err)


def test_generator_train_mode_matches_reference():
    z, cfg = load_golden()
    sd = reference_layout_state("G", cfg, cfg["seed_G"])
    _, pre = preprocessed(cfg)
    opt = orc.default_opt(ngf=cfg["ngf"], crop_size=cfg["size"], isTrain=True)
    taps = {}
    with torch.no_grad():
        out = orc.generate_fake(sd, opt, pre, True, rng_k=int(z["rng_k"][0]), taps=taps)
    _close("G train output", out, z["g_train_out"])
    for name, t in taps.items():
        assert_summary_close("tap " + name, t, z["tap/" + name], 3e-5)
    # train-mode side effects: running stats (unbiased var, momentum 0.1), u/v power iteration
    for k in z.files:
        if k.startswith("g_post/"):
            _close(k, sd[k[len("g_post/"):]], z[k], 3e-5)
    assert all(int(v) == 0 for k, v in sd.items() if k.endswith("num_batches_tracked"))


def test_generator_eval_mode_matches_reference():
    z, cfg = load_golden()
    sd = reference_layout_state("G", cfg, cfg["seed_G"])
    for k in z.files:
        if k.startswith("g_post/"):
            sd[k[len("g_post/"):]] = torch.from_numpy(z[k]).clone()
    before = {k: v.clone() for k, v in sd.items()}
    _, pre = preprocessed(cfg)
    opt = orc.default_opt(ngf=cfg["ngf"], crop_size=cfg["size"], isTrain=False)
    with torch.no_grad():
        out = orc.generate_fake(sd, opt, pre, False)
    _close("G eval output", out, z["g_eval_out"])
    assert all(torch.equal(before[k], sd[k]) for k in sd), "eval forward must not modify state"


def _d_input(z, cfg, pre, opt):
    fake = torch.from_numpy(z["g_train_out"])
    o = orc.orient_channels(pre["orient_mask"], pre["input_tag"][:, 1:2], opt)
    return torch.cat([torch.cat([pre["input_tag"], o, fake], 1), torch.cat([pre["input_tag"], o, pre["image_tag"]], 1)], 0)


def test_discriminator_matches_reference():
    z, cfg = load_golden()
    sd = reference_layout_state("D", cfg, cfg["seed_D"])
    _, pre = preprocessed(cfg)
    opt = orc.default_opt(ngf=cfg["ngf"], ndf=cfg["ndf"], crop_size=cfg["size"])
    with torch.no_grad():
        out = orc.multiscale_discriminator(_d_input(z, cfg, pre, opt), sd, opt, True)
    for i in range(2):
        for j in range(5):
            if j == 4:
                _close("D[%d][4]" % i, out[i][j], z["d/%d/4" % i])
            else:
                assert_summary_close("D[%d][%d]" % (i, j), out[i][j], z["d/%d/%d" % (i, j)], 3e-5)
    for k in z.files:
        if k.startswith("d_post/"):
            _close(k, sd[k[len("d_post/"):]], z[k], 3e-5)


def test_train_iteration_losses_and_grads_match_reference():
    z, cfg = load_golden()
    sdG = reference_layout_state("G", cfg, cfg["seed_G"])
    sdD = reference_layout_state("D", cfg, cfg["seed_D"])
    _, pre = preprocessed(cfg)
    opt = orc.default_opt(ngf=cfg["ngf"], ndf=cfg["ndf"], crop_size=cfg["size"], isTrain=True)
    names = [k for k in sdG if not k.endswith(("running_mean", "running_var", "num_batches_tracked", "weight_u", "weight_v"))]
    for n in names:
        sdG[n].requires_grad_(True)
    losses, _ = orc.compute_generator_loss(sdG, sdD, opt, pre, rng_k=int(z["rng_k"][0]))
    orc.trainer_loss(losses).backward()
    assert abs(float(losses["GAN"]) - z["g_losses"][0]) < 2e-5
    assert abs(float(losses["GAN_Feat"]) - z["g_losses"][1]) < 2e-5
    for k in z.files:
        if k.startswith("g_grad/"):
            assert_summary_close(k, sdG[k[len("g_grad/"):]].grad, z[k], 5e-3, stride=101)
    assert sdG["backgroud_enc.layer4.conv.weight"].grad is None  # dead weight (encoder.py:284)


def test_bn_data_parallel_path_matches_single_replica():
    """batchnorm.py:128-145 (sum / square-sum / clamp) vs the F.batch_norm path on one replica, and the
    two-replica reduction: sharding the batch must not change the statistics."""
    torch.manual_seed(0)
    x = torch.randn(4, 8, 6, 6) * 2 + 0.5
    mk = lambda: {"p.running_mean": torch.zeros(8), "p.running_var": torch.ones(8)}
    sd1, sd2 = mk(), mk()
    y1 = orc.param_free_bn(x, sd1, "p", True)
    parts = []

    def world(s, ss, n):  # emulates SyncMaster: sums of both halves
        xs = x[2:] if len(parts) == 0 else x[:2]
        parts.append(1)
        return s + xs.sum(dim=(0, 2, 3)), ss + (xs * xs).sum(dim=(0, 2, 3)), n * 2

    y2 = orc.param_free_bn(x[:2], sd2, "p", True, world_sums=world)
    assert torch.allclose(y1[:2], y2, atol=1e-5)
    assert torch.allclose(sd1["p.running_var"], sd2["p.running_var"], atol=1e-6)
