"""What-if timing of the dominant kernel (fused SPADE gamma|beta GEMM, up_3 shape, fp16 operands): disable one
resource at a time through MG_DBG (results are wrong in those modes - timing only) to see which one bounds it.
    1 no weight (B) loads after a CTA's first tile     2 no activation (A) loads after the first tile
    4 no epilogue work at all                          8 no epilogue global loads/stores (TMEM + math kept)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from michigan_b200 import ops
N, S, C = 8, 512, 128
dev = "cuda"
actv = torch.randn(N, S, S, 128, device=dev)
wg = torch.randn(C, 128, 3, 3, device=dev) / 34
xs = torch.randn(N, S // 2, S // 2, C, device=dev)
v = torch.ones(C, device=dev)
a16 = actv.half()
flush = torch.empty(192 * 1024 * 1024 // 4, device=dev)
variants = {
    "f16": (lambda wp: ops.conv_igemm(a16, wp, C, 3, 3, 1, 1, act=2, a_fmt=ops.F16, spade=(xs, 1, v, v, v, v), out16=(ops.BF16, True), want_f32=False),
            ops.pack_weight_gb16(wg, wg)),
    "tf32": (lambda wp: ops.conv_igemm(actv, wp, C, 3, 3, 1, 1, act=2, round_out=True, spade=(xs, 1, v, v, v, v)), ops.pack_weight_gb(wg, wg)),
}
flops = 2.0 * N * S * S * 1152 * 256
if len(sys.argv) > 1:          # e.g. `whatif_spade.py f16`: one variant only
    variants = {k: v for k, v in variants.items() if k in sys.argv[1:]}
for name, (f, wp) in variants.items():
    for dbg in (0, 1, 2, 3, 4, 8, 64, 12, 7, 11, 15):
        os.environ["MG_DBG"] = str(dbg)
        for _ in range(2):
            f(wp)
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); f(wp); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[2]
        print("%-5s MG_DBG=%2d  %.3f ms  %.0f TFLOP/s" % (name, dbg, ms, flops / ms / 1e9), flush=True)
# cycle profile of CTA 0 (MG_DBG=16): who waits for whom
import ctypes
from michigan_b200 import _lib
names = ["producer total", "producer wait-empty", "mma total", "mma wait-tmem-empty", "mma wait-full",
         "epi0 total", "epi0 wait-tmem-full", "epi0 busy", "epi0 tiles", "epi7 total", "epi7 wait-tmem-full", "epi7 busy", "epi7 tiles"]
for name, (f, wp) in variants.items():
    for dbg in (16, 16 + 4, 16 + 3, 16 + 8, 16 + 32, 16 + 64):
        os.environ["MG_DBG"] = str(dbg)
        f(wp); f(wp)
        buf = (ctypes.c_ulonglong * 16)()
        _lib.check(_lib.load().mg_debug_igemm_prof(buf), "prof")
        print("%s MG_DBG=%d cycles: " % (name, dbg) + ", ".join("%s %d" % (n, buf[i]) for i, n in enumerate(names)), flush=True)
os.environ["MG_DBG"] = "0"
