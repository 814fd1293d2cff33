"""Host mirror of ldm/models/diffusion/ddim.py (DDIMSampler: make_schedule :25-54, sample :57-111,
ddim_sampling :114-163, p_sample_ddim :166-204).  The schedule is host float64 arithmetic as in the reference; the
per-step update with classifier-free guidance is one kernel (cb_ddim_step); the UNet runs through the mirror."""
import numpy as np
import torch

from celebbasis_b200 import ops
from ldm.modules.diffusionmodules.util import make_ddim_sampling_parameters, make_ddim_timesteps, noise_like


class DDIMSampler(object):
    def __init__(self, model, schedule="linear", **kwargs):
        self.model = model
        self.ddpm_num_timesteps = model.num_timesteps
        self.schedule = schedule

    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0., verbose=True):
        self.ddim_timesteps = make_ddim_timesteps(ddim_discr_method=ddim_discretize, num_ddim_timesteps=ddim_num_steps,
                                                  num_ddpm_timesteps=self.ddpm_num_timesteps, verbose=verbose)
        ac = self.model.alphas_cumprod.detach().cpu().double().numpy()
        assert ac.shape[0] == self.ddpm_num_timesteps
        sig, a, a_prev = make_ddim_sampling_parameters(alphacums=ac, ddim_timesteps=self.ddim_timesteps, eta=ddim_eta,
                                                       verbose=verbose)
        self.ddim_sigmas, self.ddim_alphas, self.ddim_alphas_prev = sig, a, a_prev
        self.ddim_sqrt_one_minus_alphas = np.sqrt(1. - a)

    @torch.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, callback=None, normals_sequence=None,
               img_callback=None, quantize_x0=False, eta=0., mask=None, x0=None, temperature=1.,
               noise_dropout=0., score_corrector=None, corrector_kwargs=None, verbose=True, x_T=None,
               log_every_t=100, unconditional_guidance_scale=1., unconditional_conditioning=None, **kwargs):
        self.make_schedule(ddim_num_steps=S, ddim_eta=eta, verbose=verbose)
        C, H, W = shape
        return self.ddim_sampling(conditioning, (batch_size, C, H, W), callback=callback, img_callback=img_callback,
                                  x_T=x_T, log_every_t=log_every_t, temperature=temperature,
                                  unconditional_guidance_scale=unconditional_guidance_scale,
                                  unconditional_conditioning=unconditional_conditioning)

    @torch.no_grad()
    def ddim_sampling(self, cond, shape, x_T=None, callback=None, img_callback=None, log_every_t=100, temperature=1.,
                      unconditional_guidance_scale=1., unconditional_conditioning=None, **kwargs):
        device = self.model.betas.device
        b = shape[0]
        img = torch.randn(shape, device=device) if x_T is None else x_T
        intermediates = {'x_inter': [img], 'pred_x0': [img]}
        time_range = np.flip(self.ddim_timesteps)
        total_steps = self.ddim_timesteps.shape[0]
        for i, step in enumerate(time_range):
            index = total_steps - i - 1
            ts = torch.full((b,), int(step), device=device, dtype=torch.long)
            img, pred_x0 = self.p_sample_ddim(img, cond, ts, index=index, temperature=temperature,
                                              unconditional_guidance_scale=unconditional_guidance_scale,
                                              unconditional_conditioning=unconditional_conditioning)
            if callback:
                callback(i)
            if img_callback:
                img_callback(pred_x0, i)
            if index % log_every_t == 0 or index == total_steps - 1:
                intermediates['x_inter'].append(img)
                intermediates['pred_x0'].append(pred_x0)
        return img, intermediates

    @torch.no_grad()
    def p_sample_ddim(self, x, c, t, index, repeat_noise=False, use_original_steps=False, quantize_denoised=False,
                      temperature=1., noise_dropout=0., score_corrector=None, corrector_kwargs=None,
                      unconditional_guidance_scale=1., unconditional_conditioning=None):
        b, device = x.shape[0], x.device
        if unconditional_conditioning is None or unconditional_guidance_scale == 1.:
            e_u, e_c = self.model.apply_model(x, t, c), None
        else:
            x_in = torch.cat([x] * 2)            # CFG batch doubling (ddim.py:176-180): layout glue
            t_in = torch.cat([t] * 2)
            c_in = torch.cat([unconditional_conditioning, c])
            e_u, e_c = self.model.apply_model(x_in, t_in, c_in).chunk(2)
            e_u, e_c = e_u.contiguous(), e_c.contiguous()
        sigma = float(self.ddim_sigmas[index])
        noise = None
        if sigma > 0:
            noise = (noise_like(x.shape, device, repeat_noise) * temperature).contiguous()
        return ops.ddim_step(x.contiguous(), e_u, e_c, noise, scale=float(unconditional_guidance_scale),
                             a_t=float(self.ddim_alphas[index]), a_prev=float(self.ddim_alphas_prev[index]),
                             sigma_t=sigma, sqrt_one_minus_at=float(self.ddim_sqrt_one_minus_alphas[index]))
