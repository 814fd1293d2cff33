"""Device time of cb_attention_fwd vs the materialised GEMM+softmax+GEMM path (CUDA-graph replay)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from celebbasis_b200 import ops
from celebbasis_b200.unet_engine import _Attn
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
rnd = lambda *s: torch.randn(*s, generator=g).half().to(dev)

def timeit(fn, n=10):
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
    return e0.elapsed_time(e1) * 1000 / (5 * n)

for (nq, nk, dh, H) in [(4096, 4096, 40, 8), (1024, 1024, 80, 8), (4096, 77, 40, 8), (1024, 77, 80, 8), (256, 256, 160, 8)]:
    C = H * dh
    q, k, v = rnd(nq, C), rnd(nk, C), rnd(nk, C)
    o = torch.empty_like(q)
    flops = 4.0 * nq * nk * C
    res = {"case": f"attn_nq{nq}_nk{nk}_d{dh}"}
    for mode in ("flash_1pass", "flash_2pass_P", "materialised"):
        if mode.startswith("flash") and dh > 128:
            continue
        _Attn.FLASH = mode.startswith("flash")
        fn = lambda: _Attn.fwd(q, k, v, images=1, heads=H, dh=dh, nq=nq, nk=nk, scale=dh ** -0.5, out=o,
                               need_p=(mode != "flash_1pass"))
        us = timeit(fn)
        res[mode + "_us"] = round(us, 1)
        res[mode + "_tflops"] = round(flops / us / 1e6, 1)
    print(json.dumps(res), flush=True)

# ---- backward: flash (cb_attention_bwd, 2 launches) vs materialised (dP GEMM + softmax_bwd + 3 GEMMs) ----
_Attn.FLASH = True
for (nq, nk, dh, H) in [(4096, 4096, 40, 8), (1024, 1024, 80, 8), (4096, 77, 40, 8), (1024, 77, 80, 8)]:
    C = H * dh
    q, k, v, dO = rnd(nq, C), rnd(nk, C), rnd(nk, C), rnd(nq, C)
    o = torch.empty_like(q)
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    res = {"case": f"attn_bwd_nq{nq}_nk{nk}_d{dh}"}
    for mode in ("flash", "materialised"):
        _Attn.FLASH_BWD = mode == "flash"
        _Attn.FLASH_BWD_MIN_NK = 0
        saved = _Attn.fwd(q, k, v, images=1, heads=H, dh=dh, nq=nq, nk=nk, scale=dh ** -0.5, out=o)
        fn = lambda: _Attn.bwd(dO, q, k, v, saved, images=1, heads=H, dh=dh, nq=nq, nk=nk, scale=dh ** -0.5, dq=dq, dk=dk, dv=dv)
        us = timeit(fn)
        res[mode + "_us"] = round(us, 1)
        res[mode + "_tflops"] = round(10.0 * nq * nk * C / us / 1e6, 1)
    print(json.dumps(res), flush=True)
