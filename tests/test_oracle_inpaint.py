"""The oracle's restatement of the orientation-inpainting sub-net (InpaintGenerator,
generator.py:490-575, and Pix2PixModel.inpainting_orient, pix2pix_model.py:407-429) against the fixture that
tests/golden/make_golden_inpaint.py recorded from the live reference.  CPU only; the reference is not needed."""
import json
import os

import numpy as np
import torch

import michigan_oracle as orc
from michigan_b200.synth import fill_state_dict, synthetic_batch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_inpaint.npz")


def _reference_layout_state(seed):
    """State dict with the reference module's keys/shapes (generator.py:490-561, skips=False) filled deterministically."""
    sd = {}

    def sn_conv(prefix, cout, cin, k, transposed=False):
        shape = (cin, cout, k, k) if transposed else (cout, cin, k, k)
        sd[prefix + ".bias"] = torch.empty(cout)
        sd[prefix + ".weight_orig"] = torch.empty(shape)
        sd[prefix + ".weight_u"] = torch.empty(cout)
        sd[prefix + ".weight_v"] = torch.empty(cin * k * k)

    sn_conv("encoder.1", 64, 4, 7)
    sn_conv("encoder.4", 128, 64, 4)
    sn_conv("encoder.7", 256, 128, 4)
    for i in range(12):
        sn_conv("middle.%d.conv_block.1" % i, 256, 256, 3)
        sn_conv("middle.%d.conv_block.5" % i, 256, 256, 3)
    for name, co in (("query_conv", 64), ("key_conv", 64), ("value_conv", 256)):
        sd["middle.12.%s.weight" % name] = torch.empty(co, 256, 1, 1)
        sd["middle.12.%s.bias" % name] = torch.empty(co)
    sn_conv("decoder.0", 128, 512, 4, transposed=True)
    sn_conv("decoder.3", 64, 128, 4, transposed=True)
    sd["decoder.7.weight"] = torch.empty(3, 64, 7, 7)
    sd["decoder.7.bias"] = torch.empty(3)
    return fill_state_dict(sd, seed)


def test_inpaint_generator_and_orient_vs_golden():
    z = np.load(GOLDEN)
    cfg = json.loads(bytes(z["config"]).decode())
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    sd = _reference_layout_state(cfg["seed_IG"])
    assert len(sd) == 124 and sum(v.numel() for k, v in sd.items() if not k.endswith(("weight_u", "weight_v"))) == 16118211
    g = torch.Generator().manual_seed(cfg["input_seed"])
    x = torch.rand(cfg["net_batch"], 4, cfg["net_hw"], cfg["net_hw"], generator=g)
    with torch.no_grad():
        out = orc.inpaint_generator(x, sd)
    ref = torch.from_numpy(z["net_out"])
    assert out.shape == ref.shape
    assert float((out - ref).abs().max()) <= 2e-6     # fp32 reassociation only (bit-identical in the build container)
    data = synthetic_batch(1, cfg["crop_size"], cfg["data_seed"], use_ig=True)
    with torch.no_grad():
        o, orient = orc.inpainting_orient(sd, cfg["crop_size"], data["hole"], data["orient_rgb"], data["noise"], data["label_tag"])
    assert float((o - torch.from_numpy(z["io_output"])).abs().max()) <= 2e-6
    assert float((orient - torch.from_numpy(z["io_orient"])).abs().max()) <= 4e-6
    # outside the hole the output is the input map, and the 2-channel orientation vanishes outside the hair mask
    keep = (1 - data["hole"]).bool().expand_as(o)
    assert torch.equal(o[keep], data["orient_rgb"][keep])
    assert float((orient * (1 - data["label_tag"])).abs().max()) == 0.0
