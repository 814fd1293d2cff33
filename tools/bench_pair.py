"""Microbench: single-CTA tiles vs the tcgen05 cta_group::2 CTA-pair kernel on the large-M convolutions / linears
(graph replay of 8 launches, CUDA events).  Prints one JSON line per (shape, variant)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from celebbasis_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
rnd = lambda *s: torch.randn(*s, generator=g).half().to(dev)

def timeit(fn, n=8):
    fn(); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(n):
            fn()
    gr.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / n)
    return best

def variants(run, flops, name):
    orig = ops._gemm
    for label, cfg in (("1cta bn128", (128, 0)), ("1cta bn160", (160, 0)), ("1cta bn256", (256, 0)), ("pair bn128", (128, 1)),
                       ("pair bn256", (256, 1))):
        def forced(d, what, cfg=cfg):
            d.tile_n, d.cta_pair, d.splits = cfg[0], cfg[1], 1
            return orig(d, what)
        ops._gemm = forced
        try:
            us = timeit(run)
            print(json.dumps({"shape": name, "variant": label, "us": round(us, 2), "tflops": round(flops / us / 1e6, 1)}), flush=True)
        except Exception as e:
            print(json.dumps({"shape": name, "variant": label, "error": str(e)[:100]}), flush=True)
        finally:
            ops._gemm = orig

def conv(n, h, cin, cout):
    x = rnd(n * h * h, cin)
    w = ops.pack_conv_weight(torch.randn(cout, cin, 3, 3, generator=g).to(dev) * 0.02, torch.float16)
    out = torch.empty(n * h * h, cout, dtype=torch.float16, device=dev)
    geo = ops.Geo(n, h, h)
    variants(lambda: ops.conv2d(x, geo, w, cout, out=out), 2.0 * n * h * h * cin * cout * 9, f"conv3x3 {cin}->{cout} @{h}^2 x{n}")

def lin(M, N, K):
    x, w = rnd(M, K), rnd(N, K)
    out = torch.empty(M, N, dtype=torch.float16, device=dev)
    variants(lambda: ops.linear(x, w, out=out), 2.0 * M * N * K, f"linear {M}x{N}x{K}")

conv(1, 128, 512, 512); conv(1, 64, 640, 640); conv(16, 64, 320, 320); conv(16, 32, 640, 640); conv(1, 256, 256, 256)
conv(1, 512, 128, 128); lin(65536, 2560, 320); lin(65536, 320, 1280); lin(8192, 8192, 8192)
