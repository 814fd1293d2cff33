"""run the UNMODIFIED reference drivers end to end on this package.

    python tools/run_reference_drivers.py --ref /path/to/reference_checkout --work /tmp/cb_drivers [--steps 6]

Prepares a miniature but structurally identical experiment (a yaml with the aigc_id.yaml schema and the tiny model sizes of
celebbasis_b200.workload, a synthetic SD checkpoint, 4 synthetic face images + ffhq.pickle, a synthetic celebrity-name list)
and then runs, through `python -m celebbasis_b200.compat.run`,

  1. <ref>/main_id_embed.py  -t  (Trainer.fit with SetupCallback / ImageLogger / ModelCheckpoint cadence / CUDACallback)
  2. <ref>/scripts/stable_txt2img.py  with the embeddings_gs-*.pt the training run wrote

and checks the artefacts.  The reference checkout is NOT part of this repository; the scripts are executed from where it lies.
"""
import argparse
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

YAML = """
model:
  base_learning_rate: 5.0e-03
  target: ldm.models.diffusion.ddpm.LatentDiffusion
  params:
    linear_start: 0.00085
    linear_end: 0.0120
    num_timesteps_cond: 1
    log_every_t: 200
    timesteps: 1000
    first_stage_key: image
    cond_stage_key: caption
    image_size: 8
    channels: 4
    cond_stage_trainable: true
    conditioning_key: crossattn
    monitor: val/loss_simple_ema
    scale_factor: 0.18215
    use_ema: False
    embedding_reg_weight: 0.0
    unfreeze_model: False
    model_lr: 0.0
    personalization_config:
      target: ldm.modules.embedding_manager.EmbeddingManagerId
      params:
        placeholder_strings: ['sks', 'ks', 'ata', 'tre', 'ry', 'bop', 'rn', '&', '*', '`']
        initializer_words: ["face", "face", "face", "face", "face", "face", "face", "face", "face", "face"]
        max_ids: 10
        num_embeds_per_token: 2
        meta_mlp_depth: 1
        loss_type: 'none'
        meta_inner_dim: 512
        test_mode: 'coefficient'
        momentum: 0.99
        save_fp16: False
    unet_config:
      target: ldm.modules.diffusionmodules.openaimodel.UNetModel
      params:
        image_size: 32
        in_channels: 4
        out_channels: 4
        model_channels: 64
        attention_resolutions: [4, 2, 1]
        num_res_blocks: 1
        channel_mult: [1, 2, 4, 4]
        num_heads: 8
        use_spatial_transformer: True
        transformer_depth: 1
        context_dim: 768
        use_checkpoint: True
        legacy: False
    first_stage_config:
      target: ldm.models.autoencoder.AutoencoderKL
      params:
        embed_dim: 4
        monitor: val/rec_loss
        ddconfig:
          double_z: true
          z_channels: 4
          resolution: 64
          in_channels: 3
          out_ch: 3
          ch: 64
          ch_mult: [1, 2, 4, 4]
          num_res_blocks: 1
          attn_resolutions: []
          dropout: 0.0
        lossconfig:
          target: torch.nn.Identity
    cond_stage_config:
      target: ldm.modules.encoders.modules.FrozenCLIPEmbedder
      params:
        use_celeb: True
        use_svd: True
        rm_repeats: True
        celeb_txt: "{celeb_txt}"
        n_components: 512
        use_sample_reduce: False
        n_samples: 513
        use_flatten: False
        num_embeds_per_token: 2
        num_hidden_layers: 2

data:
  target: main.DataModuleFromConfig
  params:
    batch_size: 1
    num_workers: 0
    wrap: false
    train:
      target: ldm.data.face_id.FaceIdDatasetOneShot
      params:
        pickle_path: "{pickle}"
        split: train
        num_ids: 2
        specific_ids: [1, 2]
        image_size: 64
        repeats: 100
        diff_cnt: 0
    validation:
      target: ldm.data.face_id.FaceIdDatasetOneShot
      params:
        pickle_path: "{pickle}"
        split: val
        num_ids: 1
        image_size: 64
        repeats: 1
        diff_cnt: 0

lightning:
  modelcheckpoint:
    params:
      every_n_train_steps: {ckpt_every}
  callbacks:
    image_logger:
      target: main.ImageLogger
      params:
        batch_frequency: {log_every}
        max_images: 1
        increase_log_steps: False
  trainer:
    benchmark: True
    max_steps: {steps}
"""


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", required=True)
    ap.add_argument("--work", default="/tmp/cb_drivers")
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--gpus", default="0,")
    args = ap.parse_args()
    args.ref, args.work = os.path.abspath(args.ref), os.path.abspath(args.work)
    import numpy as np
    import torch
    from celebbasis_b200 import synth, workload
    os.makedirs(args.work, exist_ok=True)
    pk, _ = workload.synth_face_files(os.path.join(args.work, "faces"), n=4, hw=64, seed=0)
    rng = np.random.RandomState(0)
    syll = ["ka", "lo", "mi", "ren", "su", "ta", "vo", "xi", "ya", "zen", "bro", "cla", "dil", "fen", "gor", "hal"]
    names = sorted({" ".join("".join(rng.choice(syll, size=rng.randint(1, 4))) for _ in range(rng.randint(1, 4)))
                    for _ in range(900)})
    celeb_txt = os.path.join(args.work, "names.txt")
    open(celeb_txt, "w").write("\n".join(names) + "\n")
    yaml_path = os.path.join(args.work, "tiny_id.yaml")
    open(yaml_path, "w").write(YAML.format(celeb_txt=celeb_txt, pickle=pk, ckpt_every=max(2, args.steps // 2),
                                           log_every=max(2, args.steps // 2), steps=args.steps))
    # synthetic SD checkpoint with the reference's key names (what --actual_resume loads by name, strict=False)
    ckpt = os.path.join(args.work, "tiny_sd.ckpt")
    if not os.path.exists(ckpt):
        from ldm.models.diffusion.ddpm import LatentDiffusion
        params = workload.model_params("tiny")
        params["cond_stage_config"]["params"].update(num_hidden_layers=2)
        m = LatentDiffusion(**params)
        torch.save({"state_dict": synth.synth_state_dict(m, seed=0)}, ckpt)
        del m
    env = dict(os.environ, PYTHONPATH=ROOT, PYTHONDONTWRITEBYTECODE="1")
    logdir = os.path.join(args.work, "logs")
    cmd = [sys.executable, "-m", "celebbasis_b200.compat.run", os.path.join(args.ref, "main_id_embed.py"),
           "--base", yaml_path, "-t", "True", "--actual_resume", ckpt, "-n", "e2e", "--gpus", args.gpus,
           "--logdir", logdir, "--no-test", "True", "--max_steps", str(args.steps), "--datadir_in_name", "False"]
    print("[drivers] $", " ".join(cmd), flush=True)
    r = subprocess.run(cmd, cwd=args.ref, env=env, capture_output=True, text=True)
    sys.stdout.write(r.stdout[-6000:])
    sys.stderr.write(r.stderr[-6000:])
    result = {"train_rc": r.returncode}
    runs = sorted(glob.glob(os.path.join(logdir, "*e2e*")))
    if runs:
        ck = sorted(glob.glob(os.path.join(runs[-1], "checkpoints", "*")))
        imgs = sorted(glob.glob(os.path.join(runs[-1], "images", "**", "*.jpg"), recursive=True))
        result.update(checkpoints=[os.path.basename(c) for c in ck], n_logged_images=len(imgs))
    ok = r.returncode == 0
    emb = sorted(glob.glob(os.path.join(runs[-1], "checkpoints", "embeddings_gs-*.pt"))) if runs else []
    if ok and emb:
        # 2. scripts/stable_txt2img.py with the embedding checkpoint the training run wrote (02_start_test.sh)
        out = os.path.join(args.work, "txt2img")
        cmd = [sys.executable, "-m", "celebbasis_b200.compat.run", os.path.join(args.ref, "scripts", "stable_txt2img.py"),
               "--config", yaml_path, "--ckpt", ckpt, "--embedding_path", emb[-1], "--prompt", "a photo of sks person",
               "--outdir", out, "--n_samples", "2", "--n_iter", "1", "--ddim_steps", "4", "--H", "64", "--W", "64",
               "--scale", "10.0", "--ddim_eta", "0.0", "--eval_id1", "1", "--eval_id2", "1"]
        print("[drivers] $", " ".join(cmd), flush=True)
        r2 = subprocess.run(cmd, cwd=args.ref, env=env, capture_output=True, text=True)
        sys.stdout.write(r2.stdout[-3000:])
        sys.stderr.write(r2.stderr[-3000:])
        samples = sorted(glob.glob(os.path.join(out, "samples", "*.jpg")))
        grids = sorted(glob.glob(os.path.join(out, "*.jpg")))
        result.update(txt2img_rc=r2.returncode, n_samples_written=len(samples), n_grids=len(grids))
        ok = ok and r2.returncode == 0 and len(samples) == 2
        # the embedding checkpoint must also load in the reference's format: a list of max_ids (es, 1, K) tensors
        import torch as _t
        ck = _t.load(emb[-1], map_location="cpu")
        result.update(embedding_keys=sorted(ck.keys()), n_coefficients=len(ck.get("id_coefficients", [])),
                      coefficient_shape=list(ck["id_coefficients"][0].shape) if ck.get("id_coefficients") else None)
    print("[drivers] RESULT " + json.dumps(result), flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
