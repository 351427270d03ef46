"""GPU-box debugging aid for the backward pass: discriminator alone and single SPADEResnetBlocks
against torch autograd on the CPU oracle (smooth loss = mean of squares)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import michigan_oracle as orc
from helpers import reference_layout_state
from michigan_b200 import networks, ops
from michigan_b200.networks import autograd as ag
from michigan_b200.networks.prep import SpectralNormBatch
from michigan_b200.options import make_opt
from michigan_b200.synth import fill_state_dict

torch.set_num_threads(16)
dev = "cuda"


def rel(a, b, floor=0.0):
    return (a.cpu() - b).norm().item() / max(b.norm().item(), floor, 1e-30)


def test_D(size=64, B=4):
    cfg = dict(ngf=64, ndf=64, size=size, batch=B)
    sdD = reference_layout_state("D", cfg, 5)
    opt = make_opt(ndf=64, crop_size=size)
    D = networks.MultiscaleDiscriminator(opt); D.load_state_dict(sdD); D = D.cuda().train()
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, 7, size, size, generator=g)
    xc = x.clone().cuda().requires_grad_(True)
    outs = D(xc)
    loss = sum((t * t).mean() for o in outs for t in o)
    loss.backward()
    xo = x.clone().requires_grad_(True)
    names = [n for n, _ in D.named_parameters()]
    for n in names:
        sdD[n].requires_grad_(True)
    outs_o = orc.multiscale_discriminator(xo, sdD, orc.default_opt(ndf=64), True)
    loss_o = sum((t * t).mean() for o in outs_o for t in o)
    loss_o.backward()
    print("D loss %.6f vs %.6f" % (float(loss), float(loss_o)))
    for i in range(2):
        for j in range(5):
            print("   out[%d][%d] rel err %.2e" % (i, j, rel(outs[i][j].detach(), outs_o[i][j].detach())))
    named = dict(D.named_parameters())
    for n in names:
        print("   dgrad %-45s rel %.3e" % (n, rel(named[n].grad, sdD[n].grad)))
    print("   d input (image channels 4:7) rel %.3e ; other channels max |g| %.2e (oracle %.2e)" %
          (rel(xc.grad[:, 4:7], xo.grad[:, 4:7]), float(xc.grad[:, :4].abs().max()), float(xo.grad[:, :4].abs().max())))


def test_block(fin, fout, h, xs, N=2, blend=False):
    opt = make_opt(ngf=64, crop_size=128)
    torch.manual_seed(0)
    blk = networks.SPADEResnetBlock(fin, fout, opt)
    sd = {k: v.detach().clone() for k, v in blk.state_dict().items()}
    fill_state_dict(sd, 77)
    blk.load_state_dict(sd)
    blk = blk.cuda().train()
    g = torch.Generator().manual_seed(1)
    hs = h >> xs
    x = torch.randn(N, fin, hs, hs, generator=g)
    seg = torch.randn(N, 4, 128, 128, generator=g)
    dout = torch.randn(N, fout, h, h, generator=g)
    snb = SpectralNormBatch(blk.sn_convs())
    inv = snb.run(True)
    inv_of = {c: inv[i:i + 1].clone() for i, c in enumerate(snb.convs)}
    seg4 = ops.nchw_to_nhwc(seg.cuda())
    xn = ops.nchw_to_nhwc(x.cuda())
    out, S = ag.block_fwd(blk, xn, xs, seg4, inv_of, None)
    G = ag._Grads()
    dx, _ = ag.block_bwd(G, blk, S, ops.nchw_to_nhwc(dout.cuda()), seg4, inv_of)
    torch.cuda.synchronize()
    # oracle
    names = [n for n, _ in blk.named_parameters()]
    sdo = {("b." + k): v.clone() for k, v in sd.items()}
    for n in names:
        sdo["b." + n].requires_grad_(True)
    xo = x.clone().requires_grad_(True)
    xu = torch.nn.functional.interpolate(xo, scale_factor=2 ** xs, mode="nearest") if xs else xo
    out_o = orc.spade_resnet_block(xu, seg, sdo, "b", True)
    out_o.backward(dout)
    print("block %d->%d h%d xs%d: out rel %.2e  dx rel %.3e" % (fin, fout, h, xs, rel(out.permute(0, 3, 1, 2), out_o.detach()),
                                                               rel(dx.permute(0, 3, 1, 2), xo.grad)))
    named = dict(blk.named_parameters())
    gmax = max(sdo["b." + n].grad.norm().item() for n in names)
    for n in names:
        gg = G.get(named[n])
        print("      %-40s rel %.3e  |ref| %.2e" % (n, rel(gg, sdo["b." + n].grad, 1e-3 * gmax), sdo["b." + n].grad.norm().item()))


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "D"):
        test_D()
    if which in ("all", "blk"):
        test_block(128, 64, 32, 1)
        test_block(64, 64, 16, 0)
        test_block(128, 128, 4, 1)
