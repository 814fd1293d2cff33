"""Per-phase cycles of the flash-attention forward softmax thread (thread 0 of each CTA), averaged over CTAs and blocks."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from celebbasis_b200 import lib, ops
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
rnd = lambda *s: torch.randn(*s, generator=g).half().to(dev)
L = lib.load()
for (nq, dh, H) in ((4096, 40, 8), (1024, 80, 8)):
    C = H * dh
    q, k, v = rnd(nq, C), rnd(nq, C), rnd(nq, C)
    o = torch.empty_like(q)
    nct = (nq + 127) // 128 * H
    buf = torch.zeros(nct * 8, dtype=torch.int64, device=dev)
    for _ in range(3):
        ops.attention_fwd(q, k, v, o, images=1, heads=H, dh=dh, nq=nq, nk=nq, scale=dh ** -0.5)
    L.cb_debug_attention_timeline(ctypes.c_void_p(buf.data_ptr()))
    ops.attention_fwd(q, k, v, o, images=1, heads=H, dh=dh, nq=nq, nk=nq, scale=dh ** -0.5)
    torch.cuda.synchronize()
    L.cb_debug_attention_timeline(None)
    t = buf.view(nct, 8).double()
    nb = t[:, 6].mean().item()
    names = ["wait s_full", "tmem ld + s_empty", "row max", "pv_done wait + rescale", "exp + P store", "fence + p_full"]
    per = t[:, :6].mean(0) / nb
    print(f"nq={nq} d={dh}: blocks/CTA {nb:.0f}; cycles per block: " + ", ".join(f"{n} {c:.0f}" for n, c in zip(names, per.tolist())) + f"; total {per.sum().item():.0f}")
