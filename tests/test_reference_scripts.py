"""The drop-in boundary on the GPU: the reference's UNMODIFIED inference.py and train.py (staged under baseline/_ref by
tools/make_baseline_ref.py; /root/reference does not exist on the GPU box) run through michigan_b200.launch.

  * BASELINE.json configs[0]: `inference.py --netG spadeb --crop_size 512 --add_feat_zeros` on datasets/FFHQ_single with a
    checkpoint of deterministic weights written in the reference's layout; the tensor the script computes
    (`generated`, 576x576) is compared with the CPU oracle on the very `data` dict the reference's loader produced
    (max-abs <= 1e-3), and the JPEG it saves with the JPEG of the oracle image;
  * train.py: one epoch over the 3-sample demo set at batch 2 (1 GPU), and under torchrun on 2 GPUs when available.
"""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

import michigan_oracle as orc
from helpers import max_mean_abs, preprocessed, reference_layout_state

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models", "networks")),
                                                  reason="baseline/_ref not staged (python tools/make_baseline_ref.py)")]

DRIVER = r"""
import sys, torch, random
import numpy as np
np.random.seed(20260924); random.seed(20260924)      # the reference's loader draws its noise image with np.random (unseeded there)
sys.path[:0] = [{root!r}]
from michigan_b200 import launch, _lib
n0 = _lib.launch_count()
g = launch.main([{ref!r}] + {argv!r}, run_name="__main__")
if {dump!r}:
    torch.save({{"generated": g["generated"].detach().cpu(), "data": {{k: v.detach().cpu() for k, v in g["data"].items() if torch.is_tensor(v)}},
                "launches": _lib.launch_count() - n0}}, {dump!r})
print("DROPIN-OK launches=%d" % (_lib.launch_count() - n0))
"""

INFER = ("inference.py --name cfg1 --inference_ref_name 67172 --inference_tag_name 67172 --inference_orient_name 67172 --netG spadeb "
         "--which_epoch latest --use_encoder --noise_background --expand_mask_be --expand_th 5 --load_size 512 --crop_size 512 "
         "--add_feat_zeros --data_dir ./datasets/FFHQ_single --checkpoints_dir").split()
TRAIN = ("train.py --name tr --batchSize 2 --no_confidence_loss --no_style_loss --no_rgb_loss --no_content_loss --use_encoder "
         "--wide_edge 2 --no_background_loss --noise_background --random_expand_mask --no_vgg_loss --no_orient_loss --no_lab_loss "
         "--load_size 284 --crop_size 256 --data_dir ./datasets/FFHQ_demo_train --niter 2 --niter_decay 0 --no_html --nThreads 0 "
         "--checkpoints_dir").split()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_config1_inference_py_through_the_dropin(tmp_path):
    cfg = dict(ngf=64, ndf=64, size=512, batch=1, data_seed=13)
    sd = reference_layout_state("G", cfg, 24)
    os.makedirs(tmp_path / "cfg1")
    dump = str(tmp_path / "dump.pt")
    img_path = os.path.join(REF, "inference_samples", "fake_image.jpg")

    def run_script():
        torch.save(sd, tmp_path / "cfg1" / "latest_net_G.pth")        # util.save_network's layout: a plain state dict
        if os.path.exists(img_path):
            os.remove(img_path)
        code = DRIVER.format(root=ROOT, ref=REF, argv=INFER + [str(tmp_path)], dump=dump)
        r = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
        assert r.returncode == 0 and "DROPIN-OK" in r.stdout, r.stdout[-3000:]
        return torch.load(dump)

    # Pass 1 (uncalibrated checkpoint) only to obtain the `data` dict the reference's own loader builds from datasets/FFHQ_single.
    # A checkpoint's running statistics describe its data: fill_state_dict leaves mean 0 / var 1, so they are calibrated with ONE
    # train-mode oracle pass (momentum 1) over this very sample in the padded 576x576 geometry - what a network trained on such
    # images carries.  (Statistics of unrelated inputs make |x_hat| large and leave a 2-4e-3 tail on an otherwise 2e-5 mean error.)
    z = run_script()
    data = z["data"]
    pre = dict(input_ref=orc.one_hot(data["label_ref"].long()), input_tag=orc.one_hot(data["label_tag"].long()),
               image_ref=data["image_ref"].float(), image_tag=data["image_tag"].float(), orient_mask=data["orient"].float(),
               noise=data["noise"].float())
    with torch.no_grad():
        orc.generate_fake(sd, orc.default_opt(isTrain=True, add_feat_zeros=True), pre, True, rng_k=5, momentum=1.0)
    # Pass 2: the calibrated checkpoint through the unmodified script (same seed => the loader draws the same noise image)
    z = run_script()
    assert z["launches"] > 100, "the CUDA library did not run"
    gen, data = z["generated"], z["data"]
    assert tuple(gen.shape) == (1, 3, 576, 576)
    pre = dict(input_ref=orc.one_hot(data["label_ref"].long()), input_tag=orc.one_hot(data["label_tag"].long()),
               image_ref=data["image_ref"].float(), image_tag=data["image_tag"].float(), orient_mask=data["orient"].float(),
               noise=data["noise"].float())
    with torch.no_grad():
        ref = orc.generate_fake(sd, orc.default_opt(isTrain=False, add_feat_zeros=True), pre, False)
    mx, mn = max_mean_abs(gen, ref)
    d = (gen - ref).abs()
    print("config 1 (unmodified inference.py, 576x576) vs oracle: max-abs %.3e mean-abs %.3e (inside the 512 window %.3e, 99.99th pct %.3e)"
          % (mx, mn, float(d[..., 32:544, 32:544].max()), float(d.flatten().kthvalue(int(d.numel() * 0.9999))[0])))
    assert mx <= 1e-3 and mn <= 2e-4, (mx, mn)
    # the file the script wrote: tensor2im + crop of the 32-pixel border + JPEG (inference.py:41-56)
    from PIL import Image
    got = np.asarray(Image.open(img_path)).astype(np.float32)
    exp = ((ref[0].permute(1, 2, 0).numpy() + 1) / 2 * 255.0)[32:32 + 512, 32:32 + 512]
    assert got.shape == (512, 512, 3)
    assert np.abs(got - exp).mean() <= 6.0, np.abs(got - exp).mean()     # JPEG quantisation of a noise-like image


def test_train_py_through_the_dropin(tmp_path):
    code = DRIVER.format(root=ROOT, ref=REF, argv=TRAIN + [str(tmp_path)], dump="")
    r = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0 and "Training was successfully finished." in r.stdout and "DROPIN-OK" in r.stdout, r.stdout[-3000:]
    sdG = torch.load(tmp_path / "tr" / "latest_net_G.pth")
    ref_layout = reference_layout_state("G", dict(ngf=64, ndf=64, size=256), 0)
    assert list(sdG.keys()) == list(ref_layout.keys())
    assert all(torch.isfinite(v).all() for v in sdG.values() if v.dtype.is_floating_point)
    # two epochs x one iteration of Adam moved the weights away from their initialisation
    assert float(sdG["up_3.conv_0.weight_orig"].std()) > 0


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_train_py_two_ranks_through_torchrun(tmp_path):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), "-m", "michigan_b200.launch", REF] + TRAIN + [str(tmp_path)]
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900, env=env, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.count("Training was successfully finished.") == 2, r.stdout[-3000:]
    assert os.path.exists(tmp_path / "tr" / "latest_net_G.pth") and os.path.exists(tmp_path / "tr" / "latest_net_D.pth")
    assert not any(f.endswith(".tmp") or ".tmp." in f for f in os.listdir(tmp_path / "tr"))
