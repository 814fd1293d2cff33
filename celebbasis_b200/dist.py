"""Data-parallel plumbing: one process per GPU, ONE collective per step.

The reference shards through Lightning DDP (main_id_embed.py:596-606) and all-reduces every requires_grad
parameter (65.7 M elements incl. the never-updated iresnet weights).  Independent identities /
timesteps shard as pure data parallel here too, but the exchange is exactly the flat 525,312-element fp32 gradient
of the two trainable tensors (2.1 MB, latency-bound on NVLink 5 / NVSwitch): `all_reduce(SUM)` then * 1/world, which
is DDP's gradient averaging.  At save time the per-identity EMA coefficients, which the reference keeps rank-local and
therefore loses for ranks > 0 (ddpm.py:1519-1528), are all-gathered.
"""
import os
import sys

import torch
import torch.distributed as dist


def env_world():
    return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend=None):
    world, rank, local = env_world()
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
        # NCCL prints its version banner to stdout when the communicator is created (first collective); callers such as
        # bench.py own stdout (one JSON line), so create the communicator here with fd 1 pointed at stderr.
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
            if backend == "nccl":
                t = torch.zeros(1, device="cuda")
                dist.all_reduce(t)
                torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    return world, rank, local


def world_size():
    return dist.get_world_size() if dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_initialized() else 0


def allreduce_mean_(flat_grad):
    """In-place mean over ranks of the flat trainable-gradient buffer (the single per-step collective)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
        flat_grad.mul_(1.0 / dist.get_world_size())   # the averaging DDP applies; 0.5 M elements
    return flat_grad


def scaled_lr(base_lr, batch_size, accumulate=1):
    """main_id_embed.py:778-779: lr = accumulate_grad_batches * ngpu * bs * base_lr."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    return accumulate * world * batch_size * base_lr


def identity_shard(num_ids, rank=None, world=None):
    """Identities owned by this rank (round-robin), the data-parallel partition of BASELINE.json config 3."""
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    return [i for i in range(num_ids) if i % world == rank]


def gather_identity_state(local_coeffs, owned_ids, num_ids):
    """All-gather per-identity EMA coefficients so rank 0 can save every identity.
    local_coeffs: (num_ids, ...) tensor where only rows `owned_ids` are meaningful on this rank."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return local_coeffs
    mask = torch.zeros(num_ids, dtype=local_coeffs.dtype, device=local_coeffs.device)
    mask[owned_ids] = 1
    shape = (-1, *([1] * (local_coeffs.dim() - 1)))
    contrib = local_coeffs * mask.view(shape)
    owners = mask.clone()
    dist.all_reduce(contrib, op=dist.ReduceOp.SUM)    # a row owned by exactly one rank: sum == that rank's row
    dist.all_reduce(owners, op=dist.ReduceOp.SUM)
    # identities nobody trained keep their (identical on every rank) initial value; shared ones are averaged
    return torch.where(owners.view(shape) > 0, contrib / owners.clamp_min(1).view(shape), local_coeffs)


def barrier():
    if dist.is_initialized():
        dist.barrier()
