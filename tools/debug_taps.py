"""GPU-box debugging aid: per-block comparison of the CUDA generator against the CPU oracle.
    python tools/debug_taps.py <ngf> <size> <batch> [eval]"""
import os, sys, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import michigan_oracle as orc
from helpers import preprocessed, reference_layout_state
from michigan_b200 import networks
from michigan_b200.options import make_opt

ngf, size, batch = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
training = not (len(sys.argv) > 4 and sys.argv[4] == "eval")
cfg = dict(ngf=ngf, ndf=ngf, size=size, batch=batch, data_seed=1)
sd = reference_layout_state("G", cfg, 3)
G = networks.SPADEBGenerator(make_opt(is_train=training, ngf=ngf, ndf=ngf, crop_size=size))
G.load_state_dict(sd, strict=True)
G = G.cuda()
G.train(training)
G.collect_taps = True
_, pre = preprocessed(cfg)
random.seed(0)
th = int(size * 0.05); th = th if th % 2 == 1 else th + 1
k = random.choice([max(th - 4, 1), max(th - 2, 1), th, th + 2, th + 4])
random.seed(0)
p = {kk: v.cuda() for kk, v in pre.items()}
with torch.no_grad():
    out = G(p["input_ref"], orient_mask=p["orient_mask"], image_ref=p["image_ref"], input_tag=p["input_tag"], noise=p["noise"],
            image_tag=p["image_tag"])
torch.cuda.synchronize()
taps = {}
with torch.no_grad():
    ref = orc.generate_fake(sd, orc.default_opt(ngf=ngf, crop_size=size, isTrain=training), pre, training, rng_k=k, taps=taps)
for name, t in G.last_taps.items():
    r = taps[name]
    g = t.permute(0, 3, 1, 2).cpu()
    d = (g - r).abs()
    print("%-12s shape %-20s max|err| %.3e  ref max %.3e  ref std %.3e  rel %.2e" % (name, tuple(r.shape), d.max(), r.abs().max(), r.std(), d.max() / r.abs().max().clamp(min=1e-9)))
d = (out.cpu() - ref).abs()
print("output       max|err| %.3e mean|err| %.3e  ref std %.3e" % (d.max(), d.mean(), ref.std()))
