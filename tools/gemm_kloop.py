"""Where the time of ONE cb_gemm launch goes, warm and cold L2: per-CTA clock stamps (desc.debug_timeline) of the step's
deep-K / small-M shapes -> setup, first tile latency, k-loop time per 64-deep iteration, accumulator drain, epilogue."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from celebbasis_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
rnd = lambda *s: torch.randn(*s, generator=g).half().to(dev)
flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
GHZ = 1.965


def report(name, fn, kiters_total, cold, ncta_max=8192):
    fn(); fn()
    torch.cuda.synchronize()
    buf = torch.zeros(ncta_max * 8, dtype=torch.int64, device=dev)
    if cold:
        flush_buf.fill_(1)
        torch.cuda.synchronize()
    ops.GEMM_DEBUG_TIMELINE = buf
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record()
    torch.cuda.synchronize()
    ops.GEMM_DEBUG_TIMELINE = None
    t = buf.view(-1, 8).cpu()
    t = t[t[:, 0] > 0]
    d = lambda a, b: ((t[:, b] - t[:, a]).double() / GHZ / 1000.0)
    win = [v for k, v in ops._TUNE.items()][-1] if ops._TUNE else None
    n = t.shape[0]
    span_us = float(t[:, 7].max() - t[:, 7].min()) / 1000.0
    print(json.dumps({"case": name, "l2": "cold" if cold else "warm", "ctas": n, "tuned": win, "event_us": round(e0.elapsed_time(e1) * 1e3, 2),
                      "cta_start_spread_us": round(span_us, 2), "setup": round(d(0, 1).mean().item(), 2),
                      "first_full": round(d(1, 2).mean().item(), 2), "kloop": round(d(2, 3).mean().item(), 2),
                      "kloop_max": round(d(2, 3).max().item(), 2), "drain": round(d(3, 4).mean().item(), 2),
                      "epilogue": round(d(4, 5).mean().item(), 2), "epilogue_max": round(d(4, 5).max().item(), 2),
                      "cta_total": round(d(0, 6).mean().item(), 2), "cta_total_max": round(d(0, 6).max().item(), 2),
                      "kiters_total": kiters_total}), flush=True)


def conv(n, h, cin, cout, dgrad=False):
    x = rnd(n * h * h, cin)
    w = ops.pack_conv_weight(torch.randn(cout, cin, 3, 3, generator=g).to(dev) * 0.02, torch.float16)
    out = torch.empty(n * h * h, cout, dtype=torch.float16, device=dev)
    for cold in (False, True):
        report(f"conv3x3 {n}x{h}x{h} {cin}->{cout}", lambda: ops.conv2d(x, ops.Geo(n, h, h), w, cout, out=out), 9 * cin // 64, cold)


def lin(M, N, K):
    x, w = rnd(M, K), rnd(N, K)
    out = torch.empty(M, N, dtype=torch.float16, device=dev)
    for cold in (False, True):
        report(f"linear {M}x{N}x{K}", lambda: ops.linear(x, w, out=out), K // 64, cold)


conv(1, 16, 1280, 1280); conv(1, 8, 1280, 1280); conv(1, 32, 640, 640); conv(1, 64, 320, 320); conv(1, 16, 2560, 1280)
lin(256, 1280, 1280); lin(77, 768, 768); lin(1024, 640, 640); lin(4096, 320, 320); lin(256, 10240, 1280); lin(256, 1280, 5120)
