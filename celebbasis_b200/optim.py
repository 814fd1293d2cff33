"""torch.optim front end of cb_adamw_step (AdamW with torch defaults, ddpm.py:1442-1454)."""
import torch

from . import ops


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for g in self.param_groups:
            for p in g["params"]:
                if p.grad is None:      # torch skips parameters without a gradient (the frozen iresnet weights)
                    continue
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["m"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["v"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] += 1
                ops.adamw_step(p.data, p.grad.contiguous(), st["m"], st["v"], lr=g["lr"], beta1=g["betas"][0],
                               beta2=g["betas"][1], eps=g["eps"], weight_decay=g["weight_decay"], step=st["step"])
        return loss
