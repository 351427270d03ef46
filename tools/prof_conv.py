"""Cycle profile (MG_DBG=16, CTA 0) + timing of one implicit-GEMM conv configuration, halo mode on and off.
usage: prof_conv.py <fmt: tf32|f16|bf3> <Cin> <Cout> <size> [batch]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from michigan_b200 import ops, _lib
fmt, Cin, Cout, S = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
N = int(sys.argv[5]) if len(sys.argv) > 5 else 8
dev = "cuda"
x = torch.randn(N, S, S, Cin, device=dev)
w = torch.randn(Cout, Cin, 3, 3, device=dev) / (Cin * 9) ** 0.5
if fmt == "bf3":
    hi = x.bfloat16(); lo = (x - hi.float()).bfloat16()
    wp = ops.pack_weight16(w, None, ops.BF16, split=True)
    f = lambda: ops.conv_igemm(hi, wp, Cout, 3, 3, 1, 1, a_fmt=ops.BF16, x_lo=lo)
    passes = 3
elif fmt == "f16":
    xh = x.half()
    wp = ops.pack_weight16(w, None, ops.F16, split=False)
    f = lambda: ops.conv_igemm(xh, wp, Cout, 3, 3, 1, 1, a_fmt=ops.F16)
    passes = 1
else:
    wp = ops.pack_weight(w, None, round_tf32=True)
    f = lambda: ops.conv_igemm(x, wp, Cout, 3, 3, 1, 1)
    passes = 1
flops = 2.0 * N * S * S * 9 * Cin * Cout
names = ["prod total", "prod wait-empty", "mma total", "mma wait-tmem-empty", "mma wait-full", "epi0 total", "epi0 wait-tmem-full", "epi0 busy",
         "epi0 tiles", "epi7 total", "epi7 wait-tmem-full", "epi7 busy", "epi7 tiles"]
flush = torch.empty(192 * 1024 * 1024 // 4, device=dev)
for halo in ("0", "1"):
    os.environ["MG_HALO"] = halo
    for dbg in (0, 4):
        os.environ["MG_DBG"] = str(dbg)
        f(); f()
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); f(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[2]
        print("%s %d->%d %d^2 halo=%s MG_DBG=%d: %.3f ms  %.0f TFLOP/s algorithmic (x%d passes issued)" % (fmt, Cin, Cout, S, halo, dbg, ms, flops / ms / 1e9, passes), flush=True)
    for dbg in (16, 20):
        os.environ["MG_DBG"] = str(dbg)
        f(); f()
        buf = (ctypes.c_ulonglong * 16)()
        _lib.check(_lib.load().mg_debug_igemm_prof(buf), "prof")
        print("   halo=%s MG_DBG=%d cycles: " % (halo, dbg) + ", ".join("%s %d" % (n, buf[i]) for i, n in enumerate(names[:9])), flush=True)
os.environ["MG_DBG"] = "0"; os.environ["MG_HALO"] = "1"
