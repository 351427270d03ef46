"""Host cost of one train iteration: the same Python / launch sequence on a tiny problem (128x128, batch 1), where the GPU work is
negligible, so wall time per iteration ~ the CPU time needed to ISSUE an iteration (compare with the 512x512 step time)."""
import os, sys, time, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from michigan_b200 import _lib
from michigan_b200.options import make_opt
from michigan_b200.pix2pix_model import Pix2PixModel, train_iteration
from michigan_b200.synth import fill_state_dict, synthetic_batch

opt = make_opt(is_train=True, crop_size=128, batchSize=1)
m = Pix2PixModel(opt).train()
fill_state_dict(m.netG.state_dict(), 0); fill_state_dict(m.netD.state_dict(), 1)
oG, oD = m.create_optimizers(opt)
data = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in synthetic_batch(1, 128, 1).items()}
for _ in range(5):
    train_iteration(m, oG, oD, dict(data))
torch.cuda.synchronize()
n0 = _lib.launch_count(); t0 = time.perf_counter()
for _ in range(20):
    train_iteration(m, oG, oD, dict(data))
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 20
print("host-bound train iteration (128x128, batch 1): %.1f ms per iteration, %d library launches per iteration" % (dt * 1e3, (_lib.launch_count() - n0) // 20))
with torch.no_grad():
    pre = m.preprocess_input(dict(data))
    for _ in range(5):
        m.generate_fake(pre[0], pre[2], pre[4], pre[1], pre[3], pre[5])
    torch.cuda.synchronize(); n0 = _lib.launch_count(); t0 = time.perf_counter()
    for _ in range(50):
        m.generate_fake(pre[0], pre[2], pre[4], pre[1], pre[3], pre[5])
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 50
print("host-bound generator forward (128x128, batch 1): %.2f ms per forward, %d library launches" % (dt * 1e3, (_lib.launch_count() - n0) // 50))
