"""Device time of the step's phases, each replayed alone as a CUDA graph: VAE encode | face net + MLP + CLIP text fwd |
UNet fwd | UNet bwd | CLIP bwd + celeb-basis bwd + AdamW."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from celebbasis_b200 import ops, synth, workload
from celebbasis_b200.tokenizer import SyntheticCLIPTokenizer
from celebbasis_b200.train_step import CelebBasisStep
from oracle import torch_ref

dev = torch.device("cuda:0")
params = workload.model_params("full")
om = torch_ref.OracleModel(params, clip_layers=12)
sd = synth.synth_state_dict(om, seed=0)
del om
eng = CelebBasisStep(params, sd, synth.synth_celeb_basis(seed=0), dev, tokenizer=SyntheticCLIPTokenizer())
del sd
batch, draws = workload.synth_batch("full", B=1, seed=1234)
image, faces = batch["image"].to(dev), batch["image_ori"]["faces"].to(dev)
t, noise, peps = draws["t"].to(dev), draws["noise"].to(dev), draws["posterior_eps"].to(dev)
ids, map_np, _ = eng.prepare(batch["caption"])
ids_dev, map_dev = ids.to(dev), torch.from_numpy(map_np).to(dev)
ids_person = batch["image_ori"]["ids"]
B, T = 1, ids_dev.shape[1]
for _ in range(2):
    eng.run(image, faces, ids_person, ids_dev, map_dev, t, noise, peps)
    eng.optimizer_step()
torch.cuda.synchronize()
st = {}

def timeit(name, fn, n=3):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(json.dumps({"phase": name, "ms": round(ms, 3)}), flush=True)
    return g

def vae():
    st["z"], _ = eng.encode_first_stage(image, peps)
    st["xn"] = eng.q_sample(st["z"], t, noise.contiguous())

def face_clip():
    v = eng.face_features(faces, ids_person.shape[1])
    pre, coef, nrm = ops.celeb_mlp_fwd(v, eng.W, eng.b, eng.es)
    zc = ops.celeb_basis_fwd(coef, eng.basis)
    tok = ops.embedding_gather(ids_dev.view(-1), eng.clip.tok_table)
    emb = ops.embed_inject_fwd(tok, zc.view(-1, zc.shape[-1]), map_dev.view(-1), eng.clip.pos_table, B, T)
    st.update(v=v, pre=pre, coef=coef, nrm=nrm, zc=zc)
    st["ctx"] = eng.clip.forward(emb, B, need_grad=True)

def unet_fwd():
    st["eps"] = eng.unet.forward(st["xn"], t, st["ctx"].view(B, T, -1), need_grad=True)
    st["loss"], st["d_eps"] = ops.mse_fwd_bwd(st["eps"], noise.contiguous(), 1.0, want_grad=True)

keep = []
keep.append(timeit("vae_encode+q_sample", vae))
keep.append(timeit("face_net+mlp+clip_fwd", face_clip))
# forward/backward pairs: the tape is consumed by backward, so capture fwd+bwd together and subtract
keep.append(timeit("unet_fwd+loss", unet_fwd))
def unet_fwd_bwd():
    unet_fwd()
    st["dctx"] = eng.unet.backward(st["d_eps"])
keep.append(timeit("unet_fwd+loss+unet_bwd", unet_fwd_bwd))
def clip_fwd_bwd():
    face_clip()
    demb = eng.clip.backward(st["dctx"].view(B * T, -1))
    dz = ops.embed_inject_bwd(demb, map_dev.view(-1), st["zc"].shape[0] * eng.es, B, T)
    dcoef = ops.celeb_basis_bwd(dz.view(st["zc"].shape), eng.basis)
    ops.celeb_mlp_bwd(dcoef, st["coef"], st["nrm"], st["pre"], st["v"], eng.gW, eng.gb)
    eng.optimizer_step()
keep.append(timeit("face_net+mlp+clip_fwd + clip_bwd+celeb_bwd+adamw", clip_fwd_bwd))
