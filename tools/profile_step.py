"""Run warm-up steps, then ONE eager training step between cudaProfilerStart/Stop (for ncu --profile-from-start off)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from celebbasis_b200 import synth, workload
from celebbasis_b200.tokenizer import SyntheticCLIPTokenizer
from celebbasis_b200.train_step import CelebBasisStep
from oracle import torch_ref

kind = sys.argv[1] if len(sys.argv) > 1 else "full"
dev = torch.device("cuda:0")
params = workload.model_params(kind)
om = torch_ref.OracleModel(params, clip_layers=workload.clip_layers(kind))
sd = synth.synth_state_dict(om, seed=0)
del om
eng = CelebBasisStep(params, sd, synth.synth_celeb_basis(seed=0), dev, tokenizer=SyntheticCLIPTokenizer())
batch, draws = workload.synth_batch(kind, B=1, seed=1234)
b = {"image": batch["image"].to(dev), "caption": batch["caption"],
     "image_ori": {"faces": batch["image_ori"]["faces"].to(dev), "ids": batch["image_ori"]["ids"],
                   "num_ids": batch["image_ori"]["num_ids"]}}
d = {k: v.to(dev) for k, v in draws.items()}
for _ in range(2):
    eng.forward_backward(b, d)
    eng.optimizer_step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
eng.forward_backward(b, d)
eng.optimizer_step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled one step")
