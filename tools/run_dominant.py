"""Launch the dominant kernel (fused SPADE gamma|beta GEMM, up_3 shape, N=8) a few times - the target of
`ncu --set full` (tools/ncu_dominant.sh).  argv[1]: tf32 | f16"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from michigan_b200 import ops
mode = sys.argv[1] if len(sys.argv) > 1 else "f16"
N, S, C = 8, 512, 128
dev = "cuda"
actv = torch.randn(N, S, S, 128, device=dev)
wg = torch.randn(C, 128, 3, 3, device=dev) / 34
xs = torch.randn(N, S // 2, S // 2, C, device=dev)
v = torch.ones(C, device=dev)
if mode == "tf32":
    wp = ops.pack_weight_gb(wg, wg)
    f = lambda: ops.conv_igemm(actv, wp, C, 3, 3, 1, 1, act=2, round_out=True, spade=(xs, 1, v, v, v, v))
else:
    a16 = actv.half()
    wp = ops.pack_weight_gb16(wg, wg)
    f = lambda: ops.conv_igemm(a16, wp, C, 3, 3, 1, 1, act=2, a_fmt=ops.F16, spade=(xs, 1, v, v, v, v), out16=(ops.BF16, True), want_f32=False)
for _ in range(4):
    f()
torch.cuda.synchronize()
print("done", mode)
