"""Bisect the fused GroupNorm kernel's time (CB_GN_DBG bits: 1 no grid barrier, 2 no fold, 4 no staging load, 8 return after load)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from celebbasis_b200 import ops
dev = torch.device("cuda:0")

def per_node(name, fn, n=100):
    fn(); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(n):
            fn()
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        gr.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) * 1000 / (5 * n):.2f} us per node", flush=True)

xn = torch.randn(4096, 320, device=dev); gm = torch.ones(320, device=dev); bt = torch.zeros(320, device=dev)
y = None
for dbg in (0,):
    os.environ["CB_GN_DBG"] = str(dbg)
    per_node(f"groupnorm 4096x320 dbg={dbg}", lambda: ops.groupnorm(xn, ops.Geo(1, 64, 64), gm, bt))
for (hw, c) in ((64, 320), (64, 640), (32, 640), (32, 1280), (16, 1280), (16, 2560), (8, 1280), (8, 2560), (64, 960)):
    xs = torch.randn(hw * hw, c, device=dev); g2 = torch.ones(c, device=dev); b2 = torch.zeros(c, device=dev)
    per_node(f"groupnorm+silu {hw}x{hw}x{c}", lambda: ops.groupnorm(xs, ops.Geo(1, hw, hw), g2, b2, silu=True))
    y, stt = ops.groupnorm(xs, ops.Geo(1, hw, hw), g2, b2, silu=True)
    dy = torch.randn(hw * hw, c, device=dev).half()
    per_node(f"groupnorm_bwd      {hw}x{hw}x{c}", lambda: ops.groupnorm_bwd(dy, xs, ops.Geo(1, hw, hw), g2, b2, stt, silu=True))

# VAE-size tensors (two-kernel path: statistics + apply)
for (hw, c, dt) in ((512, 128, torch.float32), (256, 256, torch.float32), (256, 128, torch.float32), (128, 512, torch.float32), (128, 256, torch.float32)):
    xs = torch.randn(hw * hw, c, device=dev).to(dt); g2 = torch.ones(c, device=dev); b2 = torch.zeros(c, device=dev)
    per_node(f"groupnorm+silu 2-kernel {hw}x{hw}x{c} {str(dt)[6:]}", lambda: ops.groupnorm(xs, ops.Geo(1, hw, hw), g2, b2, silu=True), n=10)
