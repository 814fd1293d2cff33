"""One UNet forward at the txt2img batch (16 x 4 x 64 x 64, 77-token context) between cudaProfilerStart/Stop."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from celebbasis_b200 import synth, workload
from celebbasis_b200.unet_engine import UNetEngine
from oracle import torch_ref
dev = torch.device("cuda:0")
cfg = workload.model_params("full")["unet_config"]["params"]
ref = torch_ref.UNetModel(**cfg)
sd = synth.synth_state_dict(ref, seed=0, prefix="model.diffusion_model.")
del ref
eng = UNetEngine(cfg, sd, dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
g = torch.Generator().manual_seed(0)
x = torch.randn(B, 4, 64, 64, generator=g).to(dev)
t = torch.randint(0, 1000, (B,), generator=g).to(dev)
ctx = torch.randn(B, 77, 768, generator=g).to(dev)
for _ in range(2):
    eng.forward(x, t, ctx, need_grad=False)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
eng.forward(x, t, ctx, need_grad=False)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
