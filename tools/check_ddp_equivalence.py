"""BASELINE config 3 / VERDICT N4: under torchrun with N ranks (NCCL), each rank runs the engine on ITS sample; the
all-reduced mean gradient must equal the gradient of the single-process B=N step on the same N samples, and rank 0 must be
able to save every identity's EMA state (gather).  Rank 0 prints one JSON line.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_ddp_equivalence.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

from celebbasis_b200 import dist as cbd
from celebbasis_b200 import synth, workload
from celebbasis_b200.tokenizer import SyntheticCLIPTokenizer
from celebbasis_b200.train_step import CelebBasisStep
from oracle import torch_ref

kind = sys.argv[1] if len(sys.argv) > 1 else "full"
world, rank, local = cbd.init()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
params = workload.model_params(kind)
om = torch_ref.OracleModel(params, clip_layers=workload.clip_layers(kind))
sd = synth.synth_state_dict(om, seed=0)
del om
g = torch.Generator().manual_seed(5)
init_c = [torch.randn(2, 1, 512, generator=g)] * 10
eng = CelebBasisStep(params, sd, synth.synth_celeb_basis(seed=0), dev, tokenizer=SyntheticCLIPTokenizer(), id_coefficients=init_c)
batch, draws = workload.synth_batch(kind, B=world, seed=77)


def sl(i, j):
    b = {"image": batch["image"][i:j].to(dev).contiguous(), "caption": batch["caption"][i:j],
         "image_ori": {"faces": batch["image_ori"]["faces"][i:j].to(dev).contiguous(), "ids": batch["image_ori"]["ids"][i:j],
                       "num_ids": batch["image_ori"]["num_ids"][i:j]}}
    d = {k: v[i:j].to(dev).contiguous() for k, v in draws.items()}
    return b, d


b, d = sl(rank, rank + 1)
loss = eng.forward_backward(b, d)
local_grad = eng.grad.clone()
cbd.allreduce_mean_(eng.grad)                      # the step's single collective (NCCL over NVLink)
avg = eng.grad.clone()
losses = [torch.zeros_like(loss) for _ in range(world)]
dist.all_gather(losses, loss)
grads = [torch.zeros_like(local_grad) for _ in range(world)]
dist.all_gather(grads, local_grad)
owned = [int(batch["image_ori"]["ids"][rank, 0])]
coef_all, emb_all = eng.gathered_identity_state(owned_ids=owned)
out = None
if rank == 0:
    exact = torch.stack(grads).mean(0)
    eng2 = CelebBasisStep(params, sd, synth.synth_celeb_basis(seed=0), dev, tokenizer=SyntheticCLIPTokenizer(), id_coefficients=init_c)
    bb, dd = sl(0, world)
    loss_b = eng2.forward_backward(bb, dd)
    gb = eng2.grad
    rel = lambda a, c: float((a - c).norm() / c.norm())
    cos = lambda a, c: float(torch.dot(a.flatten(), c.flatten()) / (a.norm() * c.norm()))
    moved = [i for i in range(10) if not torch.equal(coef_all[i].cpu(), init_c[0].reshape(coef_all[i].shape))]
    out = {"world": world, "kind": kind, "backend": dist.get_backend(),
           "allreduce_mean_vs_exact_mean_rel": rel(avg, exact),
           "loss_mean_of_ranks": float(torch.stack(losses).mean()), "loss_single_process_batch": float(loss_b),
           "loss_rel": abs(float(torch.stack(losses).mean()) - float(loss_b)) / abs(float(loss_b)),
           "grad_allreduced_vs_batched_rel": rel(avg, gb), "grad_cos": cos(avg, gb),
           "identities_with_ema_after_gather": moved, "expected_identities": sorted({int(batch["image_ori"]["ids"][r, 0]) for r in range(world)})}
    print(json.dumps(out), flush=True)
dist.barrier()
dist.destroy_process_group()
