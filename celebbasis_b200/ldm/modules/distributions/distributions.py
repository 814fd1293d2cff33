"""Host mirror of ldm/modules/distributions/distributions.py:24-62 (DiagonalGaussianDistribution)."""
import torch

from celebbasis_b200 import ops


class DiagonalGaussianDistribution(object):
    def __init__(self, parameters, deterministic=False):
        self.parameters = parameters                       # (B, 2*C, H, W) fp32 NCHW
        self.deterministic = deterministic

    @property
    def mean(self):
        return torch.chunk(self.parameters, 2, dim=1)[0]

    def sample(self, eps=None, scale=1.0):
        """mean + exp(0.5*clamp(logvar,-30,20)) * eps.  Like the reference (:36) the normal draw comes from the CPU
        generator unless the caller supplies `eps` (the parity harness replays it)."""
        shape = list(self.parameters.shape)
        shape[1] //= 2
        if eps is None:
            eps = torch.randn(shape).to(device=self.parameters.device)
        if self.deterministic:
            eps = torch.zeros_like(eps)
        return ops.posterior_sample(self.parameters.contiguous(), eps.contiguous().float(), float(scale))

    def mode(self):
        return self.mean
