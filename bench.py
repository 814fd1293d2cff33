"""bench.py -- CelebBasis training steps/sec (SD-v1 UNet 512^2, bs=1/GPU) on N B200s  (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]                 # our arm (N>1: launched by torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W      # the reference's CPU path (oracle port)

One step = SURVEY.md §8 rows a1-a31 for one sample per GPU: VAE encode, CosFace R100 on the 2 face crops, celeb-basis
MLP + inject, CLIP text fwd, UNet fwd, eps-MSE, full backward to the 1024x512 MLP, gradient all-reduce, AdamW.
Synthetic 512x512 inputs / 77-token prompt, deterministic synthetic SD-v1 / CLIP / CosFace weights (no network).

Output: ONE JSON line (rank 0).  `value` = whole-job steps/s with inputs resident in HBM (CUDA events, max over ranks);
`e2e` = the same step driven from pinned HOST buffers (H2D of image/faces/draws + D2H of the loss every step);
`roofline` = tensor-pipe fraction of the dominant kernel (cb_gemm_kernel: every conv / linear / attention GEMM),
measured live by replaying exactly the step's GEMM launches as their own CUDA graph under CUDA events;
`cpu_baseline` = the oracle port of the reference step on this box's host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "celeb-basis training steps/sec (SD-v1 UNet 512^2, bs=1/GPU)"
WORKLOAD = "configs[1]: single identity per GPU, 512x512, bs=1, SD-v1 UNet + CLIP text fwd/bwd + VAE encode + CosFace R100, AdamW on the 525,312 MLP weights"


def _ncu_traffic():
    """DRAM bytes per cb_gemm launch (dram__bytes_read.sum + dram__bytes_write.sum, averaged over the step's launches)
    from the newest committed ncu pass under profiles/ (written by the command in its "source" field), or None."""
    import glob
    best = None
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_gemm_traffic_*.json"))):
        try:
            best = json.load(open(f))
        except Exception:
            pass
    return None if best is None else best.get("traffic_bytes_per_launch")


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p, "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            parts = [p.strip() for p in r.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ reference / CPU arm
def host_cpu_limits():
    """Cores this process may actually use: scheduler affinity, capped by the cgroup CPU quota (v2 cpu.max, v1
    cfs_quota_us).  os.cpu_count() reports the machine (128 on the GPU boxes) even when the container gets a slice."""
    aff = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    quota = None
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            quota = float(q) / float(per)
    except Exception:
        try:
            q = float(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            per = float(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                quota = q / per
        except Exception:
            pass
    limit = aff if quota is None else max(1, min(aff, int(quota + 0.5)))
    return {"cpu_count": os.cpu_count(), "affinity": aff, "cgroup_quota": quota, "limit": limit}


def pick_cpu_threads():
    """Thread count for the CPU arm: the quota/affinity limit, checked by a ~3 s calibration (a 3x3 fp32 convolution of
    the VAE's 256^2 level, the dominant CPU op of the step) over {limit, limit/2, limit/4, 32, 16, 8} -- a pool wider
    than the cores the container really gets runs many times slower (round 1: 327 s/step with 128 threads)."""
    import torch
    info = host_cpu_limits()
    lim = info["limit"]
    cands = sorted({c for c in (lim, lim // 2, lim // 4, 32, 16, 8) if 1 <= c <= lim} | {min(lim, 8)})
    if len(cands) == 1:
        info["threads"], info["calibration_ms"] = cands[0], {}
        torch.set_num_threads(cands[0])
        return info
    x = torch.randn(1, 128, 256, 256)
    w = torch.randn(128, 128, 3, 3)
    res = {}
    for c in cands:
        torch.set_num_threads(c)
        with torch.no_grad():
            torch.nn.functional.conv2d(x, w, padding=1)
            t0 = time.perf_counter()
            for _ in range(3):
                torch.nn.functional.conv2d(x, w, padding=1)
            res[c] = (time.perf_counter() - t0) / 3 * 1e3
    best = min(res, key=res.get)
    info["threads"], info["calibration_ms"] = best, {str(k): round(v, 1) for k, v in res.items()}
    torch.set_num_threads(best)
    return info


def cpu_reference_steps(n_timed, budget_s, min_steps=3):
    """The reference step (oracle port of LatentDiffusion.shared_step + backward + AdamW, fp32) on the host cores.
    Always runs >= min_steps steps (the first is cold: oneDNN primitive creation, page faults on 4.4 GB of weights) unless
    a single step alone exceeds the budget; returns per-step seconds, losses and the host description."""
    import torch
    from celebbasis_b200 import synth, workload
    from celebbasis_b200.tokenizer import SyntheticCLIPTokenizer
    from oracle import torch_ref
    info = pick_cpu_threads()
    params = workload.model_params("full")
    om = torch_ref.OracleModel(params, clip_layers=12)
    om.load_state_dict(synth.synth_state_dict(om, seed=0))
    om.eval()
    W, b = om.trainable()
    W.requires_grad_(True)
    b.requires_grad_(True)
    opt = torch.optim.AdamW([W, b], lr=5e-3)
    basis = synth.synth_celeb_basis(seed=0)
    tok = SyntheticCLIPTokenizer()
    times, losses = [], []
    t_start = time.perf_counter()
    want = max(min_steps, n_timed + 1)
    for i in range(want):
        batch, draws = workload.synth_batch("full", B=1, seed=1234, step=i)
        t0 = time.perf_counter()
        out = om.step(batch, draws, tok(batch["caption"])["input_ids"], basis, tok.word_id("sks"))
        opt.zero_grad(set_to_none=True)
        out["loss"].backward()
        opt.step()
        times.append(time.perf_counter() - t0)
        losses.append(float(out["loss"]))
        elapsed = time.perf_counter() - t_start
        if len(times) >= min_steps and elapsed + times[-1] > budget_s:
            break
        if elapsed + times[-1] > 2.0 * budget_s:      # hard stop: never more than twice the budget
            break
    return times, losses, info


def summarize_cpu(times):
    """Drop the cold first step whenever more than one was run; median of the rest."""
    timed = times[1:] if len(times) > 1 else times
    return statistics.median(timed), len(timed)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    times, losses, info = cpu_reference_steps(args.steps, budget_s=200.0)
    sec, n_timed = summarize_cpu(times)
    val = 1.0 / sec
    sample = (f"{len(times)} full bs=1 steps (fwd+bwd+AdamW, fp32, oracle/torch_ref.py) on {info['threads']} threads "
              f"(affinity {info['affinity']}, cgroup quota {info['cgroup_quota']}, os.cpu_count {info['cpu_count']}); "
              f"first step dropped as cold, median of the other {n_timed}; per-step s: {[round(t, 1) for t in times]}")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD},
        "cpu_baseline": {"value": val, "unit": "steps/s", "cores": info["threads"], "kind": "port", "sample": sample,
                         "thread_calibration_ms": info["calibration_ms"], "loss_first": losses[0]},
        "e2e": {"value": val, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from celebbasis_b200 import dist as cbd
    from celebbasis_b200 import lib, ops, synth, workload
    from celebbasis_b200.tokenizer import SyntheticCLIPTokenizer
    from celebbasis_b200.train_step import CelebBasisStep
    from oracle import torch_ref  # only to enumerate checkpoint keys/shapes and for the cpu_baseline leg

    world, rank, local = cbd.init()
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback in the product path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    assert lib.load().cb_device_ok() == 1, "not an sm_100 device"
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    params = workload.model_params("full")
    om = torch_ref.OracleModel(params, clip_layers=12)
    sd = synth.synth_state_dict(om, seed=0)
    del om
    eng = CelebBasisStep(params, sd, synth.synth_celeb_basis(seed=0), dev, tokenizer=SyntheticCLIPTokenizer(),
                         lr=cbd.scaled_lr(5e-3, 1) if world > 1 else 5e-3)
    del sd
    B = 1
    # ---- host-side inputs (pinned) and static device buffers -------------------------------------------------
    n_host = 4
    host = []
    for i in range(n_host):
        batch, draws = workload.synth_batch("full", B=B, seed=1234 + 101 * rank, step=i)
        batch["image_ori"]["ids"] = torch.full((B, 2), rank % 10, dtype=torch.long)
        host.append({"image": batch["image"].pin_memory(), "faces": batch["image_ori"]["faces"].pin_memory(),
                     "t": draws["t"].pin_memory(), "noise": draws["noise"].pin_memory(),
                     "eps": draws["posterior_eps"].pin_memory(), "caption": batch["caption"],
                     "ids": batch["image_ori"]["ids"]})
    st = {k: host[0][k].to(dev) for k in ("image", "faces", "t", "noise", "eps")}
    ids, map_np, _ = eng.prepare(host[0]["caption"])
    ids_dev, map_dev = ids.to(dev), torch.from_numpy(map_np).to(dev)
    ids_person = host[0]["ids"]
    h2d_bytes = sum(host[0][k].numel() * host[0][k].element_size() for k in ("image", "faces", "t", "noise", "eps"))
    loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()

    def step_device():
        return eng.run(st["image"], st["faces"], ids_person, ids_dev, map_dev, st["t"], st["noise"], st["eps"])

    # ---- eager warm-up (also builds every lazily created buffer), then capture the step as ONE CUDA graph -----
    for _ in range(2):
        loss_dev = step_device()
    torch.cuda.synchronize()
    graph, use_graph = None, not args.no_graph
    n_before = lib.launch_count()
    if use_graph:
        try:
            graph = torch.cuda.CUDAGraph()
            ops.GEMM_RECORD = []
            with torch.cuda.graph(graph):
                loss_dev = step_device()
            gemm_record, ops.GEMM_RECORD = ops.GEMM_RECORD, None
        except Exception as e:  # pragma: no cover
            sys.stderr.write(f"[bench] CUDA graph capture failed ({e!r}); running eagerly\n")
            graph, use_graph, ops.GEMM_RECORD = None, False, None
            torch.cuda.synchronize()
    if not use_graph:
        ops.GEMM_RECORD = []
        loss_dev = step_device()
        gemm_record, ops.GEMM_RECORD = ops.GEMM_RECORD, None
    launches_per_step = lib.launch_count() - n_before + 2      # + AdamW + step bump

    def one_step():
        if graph is not None:
            graph.replay()
        else:
            step_device()
        cbd.allreduce_mean_(eng.grad)     # the single per-step collective (no-op at N=1)
        eng.optimizer_step()

    def timed(fn, n):
        cbd.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        cbd.barrier()
        return float(ms.item())

    for _ in range(max(args.warmup, 3)):
        one_step()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms_total = timed(one_step, args.steps)
    ms_per_step = ms_total / args.steps
    value = world * B * 1000.0 / ms_per_step

    # ---- e2e: host buffers -> H2D -> step -> D2H loss, every step ------------------------------------------------
    counter = {"i": 0}

    def e2e_step():
        h = host[counter["i"] % n_host]
        counter["i"] += 1
        for k in ("image", "faces", "t", "noise", "eps"):
            st[k].copy_(h[k], non_blocking=True)
        one_step()
        loss_host.copy_(loss_dev, non_blocking=True)
        torch.cuda.current_stream().synchronize()      # the caller reads the loss every step

    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, args.steps) / args.steps
    clk = clocks.stop() if rank == 0 else None
    final_loss = float(loss_host.item())

    # ---- roofline of the dominant kernel: replay exactly this step's GEMM launches as their own graph ----------
    roof = None
    if rank == 0:
        import ctypes
        from celebbasis_b200.lib import GemmDesc
        L = lib.load()
        descs = [GemmDesc.from_buffer_copy(b) for b, _ in gemm_record]
        flops = sum(f for _, f in gemm_record)
        sp = ctypes.c_void_p

        def launch_all():
            s = sp(torch.cuda.current_stream().cuda_stream)
            for d in descs:
                L.cb_gemm(ctypes.byref(d), s)
        gg = torch.cuda.CUDAGraph()
        launch_all()
        torch.cuda.synchronize()
        with torch.cuda.graph(gg):
            launch_all()
        for _ in range(3):
            gg.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            gg.replay()
        e1.record()
        torch.cuda.synchronize()
        gemm_ms = e0.elapsed_time(e1) / reps
        peaks, src = _peaks()
        peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1400.0)))
        ach = flops / (gemm_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": "cb_gemm_kernel (tcgen05 GEMM / implicit-GEMM conv, all instantiations)",
                "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": _ncu_traffic(),
                "peak_source": src + " bf16_tflops_sustained (kernel timed inside a long step)",
                "launches_per_step": len(descs), "avg_launch_us": gemm_ms * 1e3 / len(descs),
                "algorithmic_gflop_per_step": flops / 1e9, "gemm_ms_per_step": gemm_ms,
                "gemm_share_of_step": gemm_ms / ms_per_step}

    # ---- CPU baseline (rank 0, N=1 only): the oracle port on this box's host cores --------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu_baseline:
        del eng, graph
        torch.cuda.empty_cache()
        times, _, threads = cpu_reference_steps(2, budget_s=60.0)
        sec = min(times)
        cpu = {"value": 1.0 / sec, "unit": "steps/s", "cores": threads, "kind": "port",
               "sample": f"{len(times)} full bs=1 step(s) (fwd+bwd+AdamW, fp32, oracle/torch_ref.py on host cores); best"}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16 operands / f32 accumulate+residual (reference: f32 with TF32 convs)",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "per_gpu_batch": B, "parallelism": f"dp{world}",
                       "cuda_graph": bool(use_graph),
                       "l2": "inputs larger than L2: 2.1 GB fp16 weights + ~4 GB activations are streamed every step (L2 = 126 MB)",
                       "v100_published_it_s": 2.75, "final_loss": final_loss},
            "e2e": {"value": world * B * 1000.0 / ms_e2e, "unit": "steps/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches_per_step * args.steps),
            "gpu_launches_per_step": int(launches_per_step),
            "clocks": clk, "roofline": roof, "cpu_baseline": cpu,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
