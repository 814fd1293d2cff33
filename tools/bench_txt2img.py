"""BASELINE config 4: scripts/stable_txt2img.py semantics through the ldm mirror -- 50 DDIM steps, eta 0, CFG scale 10,
512x512, n_samples images per GPU (UNet batch 2*n_samples), fp16 operands, + VAE decode.  Prints one JSON line
(images/s on this GPU; synthetic weights, synthetic coefficients)."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F

ap = argparse.ArgumentParser()
ap.add_argument("--n-samples", type=int, default=8)
ap.add_argument("--ddim-steps", type=int, default=50)
ap.add_argument("--scale", type=float, default=10.0)
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--kind", default="full")
args = ap.parse_args()

from celebbasis_b200 import lib, synth, workload
from ldm.models.diffusion.ddim import DDIMSampler
from ldm.models.diffusion.ddpm import LatentDiffusion

dev = torch.device("cuda:0")
params = workload.model_params(args.kind)
params["cond_stage_config"]["params"].update(device="cuda")
model = LatentDiffusion(**params)
sd = synth.synth_state_dict(model, seed=0)
model.load_state_dict(sd, strict=False)
del sd
model = model.to(dev).eval()
model.cond_stage_model.celeb_embeddings = synth.synth_celeb_basis(seed=0).to(dev)
g = torch.Generator().manual_seed(3)
model.embedding_manager.id_coefficients = [F.normalize(torch.randn(2, 1, 512, generator=g), dim=-1) for _ in range(10)]
B = args.n_samples
prompts = ["a photo of sks person"] * B
image_ori = {"faces": None, "ids": [[i % 10, i % 10] for i in range(B)], "num_ids": torch.ones(B, dtype=torch.long)}
hw = 64 if args.kind == "full" else 8

def run(steps):
    with torch.no_grad():
        uc = model.get_learned_conditioning([""] * B)
        c = model.get_learned_conditioning(prompts, image_ori=image_ori)
        x_T = torch.randn(B, 4, hw, hw, generator=g).to(dev)
        sampler = DDIMSampler(model)
        samples, _ = sampler.sample(S=steps, conditioning=c, batch_size=B, shape=[4, hw, hw], verbose=False,
                                    unconditional_guidance_scale=args.scale, unconditional_conditioning=uc, eta=0.0, x_T=x_T)
        return model.decode_first_stage(samples)

run(2)                                    # warm-up: builds engines, autotunes the batch-2B GEMM shapes
torch.cuda.synchronize()
n0 = lib.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.iters):
    img = run(args.ddim_steps)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.iters
assert torch.isfinite(img.float()).all()
print(json.dumps({"metric": "DDIM txt2img images/sec (50 steps, 512x512, CFG)", "value": B / (ms / 1e3), "unit": "images/s",
                  "n_gpus": 1, "ms_per_batch": ms, "config": {"n_samples": B, "ddim_steps": args.ddim_steps, "scale": args.scale,
                                                             "unet_batch": 2 * B, "kind": args.kind},
                  "gpu_launches_per_batch": (lib.launch_count() - n0) // args.iters, "data": "synthetic",
                  "algorithmic_tflop_per_image": 82.9 if args.kind == "full" else None,
                  "achieved_tflops": (82.9 * B / (ms / 1e3)) if args.kind == "full" else None}))
