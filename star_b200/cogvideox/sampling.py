"""Sampling loop of STAR's CogVideoX path (cogvideox-based/sat): what ``SATVideoDiffusionEngine.sample_sr``
(diffusion_video.py:245-292) runs between the LQ latent and the VAE decode, for the shipped configuration
(configs/cogvideox_5b/cogvideox_5b_infer_sr.yaml:20-33,:157-174):

    sampler      VPSDEDPMPP2MSampler, 50 steps                    sgm/modules/diffusionmodules/sampling.py:574-687
    schedule     ZeroSNRDDPMDiscretization(shift_scale=1)         discretizer.py:74-126
    guidance     DynamicCFG(scale 6, exp 5, 50 steps)             guiders.py:62-80, input layout :44-58
    denoiser     DiscreteDenoiser(num_idx 1000) + VideoScaling    denoiser.py:9-77, denoiser_scaling.py:51-60
    network      DiffusionTransformer on cat(noisy, LQ) latents   wrappers.py:25-41

The reference spreads one solver step over six classes and ~25 small device launches that recompute log-SNRs from 1-element
tensors every step.  Here every per-step scalar (sigma quantised to the 1000-entry table, c_skip / c_out, the three solver
multipliers, the noise multiplier, the guidance scale) is a host float computed ONCE per schedule (``StepPlan``); a step is the
DiT forward of the CFG pair plus two fused elementwise expressions on the 1.1 M-element latent.  The random stream is consumed
exactly like the reference does (one draw in the first step, two per later step of which the first is discarded, none in the
last), so a seeded run reproduces the reference's noise.
"""
import math
from dataclasses import dataclass

import numpy as np
import torch


def _zero_snr_table(alphas_cumprod):
    """sqrt(alpha_bar) of the selected timesteps, shifted/scaled to reach exactly 0 at the last one (discretizer.py:107-113),
    evaluated in float32 like the reference; returned ascending in t (index 0 = least noisy)."""
    a = torch.tensor(alphas_cumprod, dtype=torch.float32).sqrt()
    a0, aT = a[0].clone(), a[-1].clone()
    return (a - aT) * (a0 / (a0 - aT))


def alphas_cumprod_linear(num_timesteps=1000, linear_start=0.00085, linear_end=0.0120, shift_scale=1.0):
    """`make_beta_schedule("linear")` (sgm/modules/diffusionmodules/util.py) + the SNR shift of discretizer.py:91-95"""
    betas = (torch.linspace(linear_start ** 0.5, linear_end ** 0.5, num_timesteps, dtype=torch.float64) ** 2).numpy()
    ac = np.cumprod(1.0 - betas, axis=0)
    return ac / (shift_scale + (1 - shift_scale) * ac)


@dataclass
class Step:
    timestep: int          # DDPM index fed to the network (c_noise of VideoScaling)
    c_skip: float          # sqrt(alpha_bar) quantised to the 1000-entry table (DiscreteDenoiser.possibly_quantize_sigma)
    c_out: float
    cfg_scale: float       # DynamicCFG
    last: bool             # idx == 1: x <- denoised
    mult1: float = 0.0
    mult2: float = 0.0
    mult3: float = 1.0
    mult4: float = 0.0
    mult_noise: float = 0.0


class StepPlan:
    """all scalars of a VPSDEDPMPP2MSampler run"""

    def __init__(self, num_steps=50, num_idx=1000, shift_scale=1.0, cfg_scale=6.0, cfg_exp=5, cfg_steps=50):
        ac = alphas_cumprod_linear(num_idx, shift_scale=shift_scale)
        ts = np.linspace(num_idx - 1, 0, num_steps, endpoint=False).astype(int)[::-1]      # ascending (discretizer.py:11-12)
        sub = torch.flip(_zero_snr_table(ac[ts]), (0,))                                   # 0 (t = 999) ... ~0.99
        acs = torch.cat([sub, sub.new_ones(1)])                                           # sampling.py:492
        timesteps = [-1] + [int(t) for t in ts]                                           # sampling.py:493
        full = _zero_snr_table(ac)                                                        # DiscreteDenoiser.sigmas[t], t ascending
        self.alphas_cumprod_sqrt, self.timesteps, self.steps = acs, timesteps, []
        f32 = torch.float32

        def lam(a):
            return ((a ** 2 / (1 - a ** 2)) ** 0.5).log()

        for i in range(num_steps):
            a, a_next = acs[i], acs[i + 1]
            t = timesteps[-(i + 1)]
            q = full[(a - full).abs().argmin()]                                           # nearest table entry
            step_index = num_steps - t                                                    # sampling.py:521 (as written there)
            g = 1 + cfg_scale * (1 - math.cos(math.pi * (step_index / cfg_steps) ** cfg_exp)) / 2
            st = Step(timestep=t, c_skip=float(q), c_out=float(-((1 - q ** 2) ** 0.5)), cfg_scale=g, last=(num_steps - i == 1))
            if not st.last:
                h = lam(a_next) - lam(a)
                st.mult1 = float(((1 - a_next ** 2) / (1 - a ** 2)) ** 0.5 * (-h).exp())
                st.mult2 = float((-2 * h).expm1() * a_next)
                st.mult_noise = float((1 - a_next ** 2) ** 0.5 * (1 - (-2 * h).exp()) ** 0.5)
                if i > 0:
                    r = (lam(a) - lam(acs[i - 1])) / h
                    st.mult3, st.mult4 = float(1 + 1 / (2 * r)), float(1 / (2 * r))
            self.steps.append(st)
        assert f32 == acs.dtype


class VPSDEDPMPP2MSampler:
    """``sampler(network, x, cond, uc, lq)``: x (B, T, 16, h, w) fp32 start noise, cond / uc dicts with 'crossattn'
    (B, 226, 4096), lq (2B, T, 16, h, w) LQ latents already doubled for CFG (diffusion_video.py:283) -> denoised latent (B, T, 16, h, w).
    ``network(x_in (2B, T, 32, h, w), timesteps=(2B,), context=(2B, 226, 4096))`` is the DiffusionTransformer."""

    def __init__(self, num_steps=50, shift_scale=1.0, cfg_scale=6.0, cfg_exp=5, cfg_steps=50, dtype=torch.bfloat16):
        self.num_steps, self.dtype = num_steps, dtype
        self.plan = StepPlan(num_steps, 1000, shift_scale, cfg_scale, cfg_exp, cfg_steps)

    @torch.no_grad()
    def step(self, network, x, old, st, context, lq=None):
        """one solver step: DiT forward of the CFG pair, guidance, DPM-Solver++(2M) SDE update -> (x_next, denoised)"""
        B = x.shape[0]
        xin = torch.cat([x, x], 0)
        if lq is not None:
            xin = torch.cat((xin, lq.to(xin.dtype)), dim=2)
        ts = torch.full((2 * B,), float(st.timestep), device=x.device, dtype=x.dtype)
        out = network(xin, timesteps=ts, context=context)
        x_u, x_c = (out.float() * st.c_out + xin[:, :, :x.shape[2]] * st.c_skip).chunk(2)  # Denoiser.forward: net * c_out + x * c_skip
        denoised = x_u + st.cfg_scale * (x_c - x_u)                                    # NoDynamicThresholding (sampling_utils.py:8-11)
        if st.last:
            return denoised, denoised
        noise = torch.randn_like(x)
        if old is not None:
            noise = torch.randn_like(x)                                                 # the reference draws twice, uses the second
            d = st.mult3 * denoised - st.mult4 * old
        else:
            d = denoised
        return st.mult1 * x - st.mult2 * d + st.mult_noise * noise, denoised

    @torch.no_grad()
    def __call__(self, network, x, cond, uc=None, lq=None, callback=None):
        uc = cond if uc is None else uc
        context = torch.cat((uc["crossattn"], cond["crossattn"]), 0).to(self.dtype)        # unconditional branch first (guiders.py:47-49)
        x = x.float()
        old = None
        for i, st in enumerate(self.plan.steps):
            x, old = self.step(network, x, old, st, context, lq)
            if callback is not None:
                callback(i, x)
        return x


def split_cfg_pair(network, group=None):
    """Multi-GPU axis of one CogVideoX clip: the two branches of the CFG pair on two ranks.  Wraps ``network`` so that
    rank 0 / 1 of ``group`` evaluates row 0 (unconditional) / row 1 (conditional) of the batch-2 input the sampler builds, and ONE
    all-gather of the (1, T, 16, h, w) prediction per solver step reassembles the pair -- every rank then applies the identical
    guidance and solver update (same seed => same solver noise), so the ranks stay bit-identical without further traffic."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    assert world == 2, "the CFG pair splits over exactly two ranks (more GPUs = replicas of other clips)"

    def pair(xin, timesteps=None, context=None, **kw):
        assert xin.shape[0] == 2, "split_cfg_pair serves one clip (batch 1) per pair of ranks"
        out = network(xin[rank:rank + 1], timesteps=timesteps[rank:rank + 1], context=context[rank:rank + 1], **kw).contiguous()
        both = [torch.empty_like(out) for _ in range(world)]
        dist.all_gather(both, out, group=group)
        return torch.cat(both, 0)
    return pair


@torch.no_grad()
def sample_sr_latent(network, sampler, cond, uc, lq_latent, generator_seed=None, randn=None):
    """the part of ``sample_sr`` (diffusion_video.py:245-292) after the LQ clip has been encoded: start noise of the LQ latent's
    shape (drawn on the host like the reference, :256, unless given), the LQ latent doubled for the CFG pair, sampler -> latent in
    the model dtype"""
    if generator_seed is not None:
        torch.manual_seed(generator_seed)
    if randn is None:
        randn = torch.randn(lq_latent.shape, dtype=torch.float32).to(lq_latent.device)
    lq = torch.cat((lq_latent, lq_latent), dim=0)
    return sampler(network, randn, cond, uc=uc, lq=lq).to(sampler.dtype)
