"""cProfile of the host side of one train iteration (the step is launch-bound: ~190 ms of Python per 205 ms step)."""
import cProfile, pstats, os, sys, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from michigan_b200.options import make_opt
from michigan_b200.synth import fill_state_dict, synthetic_batch
from michigan_b200.trainer import Pix2PixTrainer
torch.cuda.set_device(0)
opt = make_opt(is_train=True, gpu_ids=[0], batchSize=8, niter=50, niter_decay=0)
trainer = Pix2PixTrainer(opt)
m = trainer.pix2pix_model_on_one_gpu
fill_state_dict(m.netG.state_dict(), 0); fill_state_dict(m.netD.state_dict(), 1)
m.train()
data = synthetic_batch(8, 512, 1234)
host = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in data.items()}
def step():
    trainer.run_generator_one_step(dict(host))
    trainer.run_discriminator_one_step(dict(host))
for _ in range(2):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
step()
pr.disable()
torch.cuda.synchronize()
for key in ("cumulative", "tottime"):
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats(key).print_stats(38)
    print(s.getvalue()[:9000])
