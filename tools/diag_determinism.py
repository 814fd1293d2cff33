"""Diagnostic: which launcher is run-to-run nondeterministic?  Every ops.conv2d / linear / groupnorm call of a VAE
encode + CosFace forward is issued twice; bitwise mismatches are reported with the shape and the tuned tile/split."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from celebbasis_b200 import ops, synth, workload
from celebbasis_b200.tokenizer import SyntheticCLIPTokenizer
from celebbasis_b200.train_step import CelebBasisStep
from oracle import torch_ref
dev = torch.device("cuda:0")
kind = sys.argv[1] if len(sys.argv) > 1 else "tiny"
params = workload.model_params(kind)
om = torch_ref.OracleModel(params, clip_layers=workload.clip_layers(kind))
sd = synth.synth_state_dict(om, seed=0)
eng = CelebBasisStep(params, sd, synth.synth_celeb_basis(seed=0), dev, tokenizer=SyntheticCLIPTokenizer())
batch, draws = workload.synth_batch(kind, B=1, seed=1234)
img, faces, peps = batch["image"].to(dev), batch["image_ori"]["faces"].to(dev), draws["posterior_eps"].to(dev)
eng.encode_first_stage(img, peps); eng.face_features(faces, 2); torch.cuda.synchronize()   # autotune
stats = collections.OrderedDict()
last_desc = {}
orig_gemm = ops._gemm
def spy_gemm(d, what):
    orig_gemm(d, what)
    last_desc["d"] = (d.M, d.N, d.K, d.conv, d.img_h, d.tile_n, d.splits, d.stages, d.d_dtype)
ops._gemm = spy_gemm
def wrap(name):
    fn = getattr(ops, name)
    def w(*a, **k):
        if k.get("out") is not None:
            return fn(*a, **k)
        r1 = fn(*a, **k)
        d1 = last_desc.get("d")
        r2 = fn(*a, **k)
        t1 = r1[0] if isinstance(r1, tuple) else r1
        t2 = r2[0] if isinstance(r2, tuple) else r2
        same = torch.equal(t1, t2)
        key = (name, tuple(t1.shape), str(t1.dtype), d1 if name in ("conv2d", "linear") else None)
        s = stats.setdefault(key, [0, 0, 0.0])
        s[0] += 1
        if not same:
            s[1] += 1
            s[2] = max(s[2], ((t1.float() - t2.float()).norm() / t2.float().norm()).item())
        return r1
    setattr(ops, name, w)
for n in ("conv2d", "linear", "groupnorm", "channel_affine_act", "cast", "l2norm_rows"):
    wrap(n)
for rep in range(3):
    eng.encode_first_stage(img, peps); eng.face_features(faces, 2)
torch.cuda.synchronize()
bad = {k: v for k, v in stats.items() if v[1]}
print(f"{len(stats)} distinct calls, {len(bad)} nondeterministic")
for k, v in bad.items():
    print(k, "calls", v[0], "mismatch", v[1], "max rel", f"{v[2]:.2e}")
