"""Golden fixture for the orientation-inpainting sub-net (the first "next" row), generated from
the LIVE reference in the build container:

    python tests/golden/make_golden_inpaint.py

The unmodified `InpaintGenerator` (generator.py:490-575) and `Pix2PixModel.inpainting_orient`
(pix2pix_model.py:407-429) are run on deterministic weights (michigan_b200.synth.fill_state_dict, seed 21) and
inputs; the script asserts that oracle/michigan_oracle.py reproduces them and stores the reference's outputs in
tests/golden/golden_inpaint.npz (re-checked without the reference by tests/test_oracle_inpaint.py).
"""
import json
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import ref_shims  # noqa: E402
import michigan_oracle as orc  # noqa: E402
from michigan_b200.synth import fill_state_dict, synthetic_batch  # noqa: E402

CFG = dict(seed_IG=21, net_hw=32, net_batch=2, input_seed=5, crop_size=64, data_seed=78)


def main():
    torch.set_num_threads(8)
    ref_shims.import_reference()
    from models.networks.generator import InpaintGenerator
    from models.pix2pix_model import Pix2PixModel
    c = CFG
    net = InpaintGenerator(None).eval()
    fill_state_dict(net.state_dict(), c["seed_IG"])
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    out = {"config": np.frombuffer(json.dumps(c).encode(), dtype=np.uint8)}

    # ---- the network alone (eval mode, as the reference always runs it: pix2pix_model.py:196-198)
    g = torch.Generator().manual_seed(c["input_seed"])
    x = torch.rand(c["net_batch"], 4, c["net_hw"], c["net_hw"], generator=g)
    with torch.no_grad():
        ref = net(x)
        got = orc.inpaint_generator(x, sd)
    err = float((ref - got).abs().max())
    print("InpaintGenerator %s: max|ref - oracle| %.2e (range %.3f..%.3f)" % (tuple(ref.shape), err, float(ref.min()), float(ref.max())))
    assert err <= 1e-6
    out["net_out"] = ref.numpy()

    # ---- inpainting_orient at a crop size != 256 (nearest resize to 256 and back)
    data = synthetic_batch(1, c["crop_size"], c["data_seed"], use_ig=True)
    fake_self = SimpleNamespace(opt=SimpleNamespace(crop_size=c["crop_size"]), netIG=net)
    with torch.no_grad():
        ref_out, ref_orient = Pix2PixModel.inpainting_orient(fake_self, data["hole"], data["orient_rgb"], data["noise"], data["label_tag"])
        got_out, got_orient = orc.inpainting_orient(sd, c["crop_size"], data["hole"], data["orient_rgb"], data["noise"], data["label_tag"])
    e1, e2 = float((ref_out - got_out).abs().max()), float((ref_orient - got_orient).abs().max())
    print("inpainting_orient: max|ref - oracle| output %.2e orient %.2e (hole pixels %d)" % (e1, e2, int(data["hole"].sum())))
    assert e1 <= 1e-6 and e2 <= 1e-6
    out["io_output"] = ref_out.numpy()
    out["io_orient"] = ref_orient.numpy()
    path = os.path.join(HERE, "golden_inpaint.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
