"""The drop-in boundary on the CPU-only build box: the reference's UNMODIFIED train.py and inference.py (staged under
baseline/_ref by tools/make_baseline_ref.py) are executed through michigan_b200.launch with the hot-path classes
installed and the no-op library of tests/dryrun.py (values meaningless): option parsing, the name-based class factory,
construction protocol, data loading, the reference trainer driving our autograd Functions, and checkpoint writing all
run for real.  The numerical version of this test is tests/test_reference_scripts.py (-m gpu)."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models", "networks")),
                                reason="baseline/_ref not staged (python tools/make_baseline_ref.py)")

DRIVER = r"""
import sys
sys.path[:0] = [{root!r}, {root!r} + "/tests", {root!r} + "/oracle"]
from dryrun import dry_run
from michigan_b200 import launch
with dry_run():
    g = launch.main([{ref!r}] + {argv!r}, run_name="__main__")
print("DROPIN-OK")
"""

TRAIN = ("train.py --name dry --batchSize 2 --no_confidence_loss --no_style_loss --no_rgb_loss --no_content_loss --use_encoder "
         "--wide_edge 2 --no_background_loss --noise_background --random_expand_mask --no_vgg_loss --no_orient_loss --no_lab_loss "
         "--load_size 72 --crop_size 64 --data_dir ./datasets/FFHQ_demo_train --niter 1 --niter_decay 0 --no_html --nThreads 0 "
         "--gpu_ids -1 --checkpoints_dir").split()
INFER = ("inference.py --name dry --inference_ref_name 67172 --inference_tag_name 67172 --inference_orient_name 67172 --netG spadeb "
         "--which_epoch latest --use_encoder --noise_background --expand_mask_be --expand_th 5 --load_size 64 --crop_size 64 "
         "--add_feat_zeros --data_dir ./datasets/FFHQ_single --gpu_ids -1 --checkpoints_dir").split()


def _run(argv, tmp_path):
    code = DRIVER.format(root=ROOT, ref=REF, argv=argv + [str(tmp_path)])
    r = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and "DROPIN-OK" in r.stdout, r.stdout[-3000:]
    return r.stdout


def test_unmodified_train_py_then_inference_py(tmp_path):
    out = _run(TRAIN, tmp_path)
    assert "Network [SPADEBGenerator] was created" in out and "Training was successfully finished." in out
    from michigan_b200 import networks
    from michigan_b200.options import make_opt
    opt = make_opt(ngf=64, ndf=64, crop_size=64, gpu_ids=[])
    for label, net in (("G", networks.SPADEBGenerator(opt)), ("D", networks.MultiscaleDiscriminator(opt))):
        sd = torch.load(os.path.join(tmp_path, "dry", "latest_net_%s.pth" % label), map_location="cpu")
        own = net.state_dict()
        assert list(sd.keys()) == list(own.keys()) and all(sd[k].shape == own[k].shape for k in own), label
        assert not any(f.endswith(".tmp") for f in os.listdir(os.path.join(tmp_path, "dry")))
    # inference.py picks the checkpoint up (eval mode, --add_feat_zeros => 64+64 padded input) and writes its image
    img = os.path.join(REF, "inference_samples", "fake_image.jpg")
    if os.path.exists(img):
        os.remove(img)
    out = _run(INFER, tmp_path)
    assert "process image" in out and os.path.exists(img)
