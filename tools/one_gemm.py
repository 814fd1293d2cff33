"""Launch a few representative cb_gemm shapes (for ncu --set full captures)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from celebbasis_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
rnd = lambda *s: torch.randn(*s, generator=g).half().to(dev)
def conv(n, h, cin, cout):
    x = rnd(n * h * h, cin)
    w = ops.pack_conv_weight(torch.randn(cout, cin, 3, 3, generator=g).to(dev) * 0.02, torch.float16)
    out = torch.empty(n * h * h, cout, dtype=torch.float16, device=dev)
    for _ in range(3):
        ops.conv2d(x, ops.Geo(n, h, h), w, cout, out=out)
def lin(M, N, K):
    x, w = rnd(M, K), rnd(N, K)
    out = torch.empty(M, N, dtype=torch.float16, device=dev)
    for _ in range(3):
        ops.linear(x, w, out=out)
conv(1, 64, 320, 320)
conv(1, 8, 1280, 1280)
lin(77, 768, 768)
conv(1, 128, 512, 512)
torch.cuda.synchronize()
print("done")
