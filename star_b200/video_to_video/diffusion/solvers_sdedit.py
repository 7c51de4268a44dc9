"""k-diffusion style solvers used by STAR's sampler (host side, fp32 state).

Mirrors video_to_video/diffusion/solvers_sdedit.py: ``sample_dpmpp_2m_sde``
(:144-203, the solver the pipeline uses, video_to_video_model.py:109) and
``sample_heun`` (:34-74), with the same argument meaning.  The latent update is
a handful of elementwise ops on a (1,4,F,h,w) fp32 tensor per step -- ~1e-6 of
the step's work -- so it stays in torch on the sampler's CUDA stream.

Noise: the reference draws the SDE noise from ``torchsde.BrownianTree``
(:110-140), an un-vendored dependency (torchsde==0.2.6) seeded from the global
RNG.  The solver only ever asks for increments over consecutive, disjoint
sigma intervals, which are independent N(0, dt) variables, and divides them by
sqrt(dt): each call therefore returns an independent N(0, I) tensor.
``IntervalNoiseSampler`` produces exactly that from a private generator; pass
``noise_sampler=`` to inject a specific stream (the parity tests inject the
same stream into the reference).
"""
import torch

from ..utils.logger import get_logger

logger = get_logger()

__all__ = ["sample_dpmpp_2m_sde", "sample_heun", "IntervalNoiseSampler",
           "BrownianTreeNoiseSampler", "get_scalings", "get_ancestral_step"]


def get_ancestral_step(sigma_from, sigma_to, eta=1.0):
    if not eta:
        return sigma_to, 0.0
    up = min(sigma_to, eta * (sigma_to ** 2 * (sigma_from ** 2 - sigma_to ** 2) / sigma_from ** 2) ** 0.5)
    return (sigma_to ** 2 - up ** 2) ** 0.5, up


def get_scalings(sigma):
    """(c_out, c_in) of the VP<->k-diffusion change of variables (ref :27-30)."""
    return -sigma, 1.0 / (sigma ** 2 + 1.0) ** 0.5


class IntervalNoiseSampler:
    """Independent N(0, I) per (sigma, sigma_next) interval; see module doc."""

    def __init__(self, x, sigma_min=None, sigma_max=None, seed=None, transform=lambda s: s):
        if seed is None:
            seed = torch.randint(0, 2 ** 63 - 1, []).item()     # like ref :87-88
        self.gen = torch.Generator(device=x.device).manual_seed(int(seed))
        self.shape, self.dtype, self.device = x.shape, x.dtype, x.device

    def __call__(self, sigma, sigma_next):
        return torch.randn(self.shape, generator=self.gen, device=self.device, dtype=self.dtype)


BrownianTreeNoiseSampler = IntervalNoiseSampler      # reference name (:110)


@torch.no_grad()
def sample_heun(noise, model, sigmas, s_churn=0.0, s_tmin=0.0, s_tmax=float("inf"), s_noise=1.0,
                show_progress=True, variant_info=None):
    """Karras et al. (2022) Alg. 2 (ref :34-74)."""
    x = noise * sigmas[0]
    n = len(sigmas) - 1
    for i in range(n):
        gamma = 0.0
        if s_tmin <= sigmas[i] <= s_tmax and sigmas[i] < float("inf"):
            gamma = min(s_churn / n, 2 ** 0.5 - 1)
        eps = torch.randn_like(x) * s_noise
        sig_hat = sigmas[i] * (gamma + 1)
        if gamma > 0:
            x = x + eps * (sig_hat ** 2 - sigmas[i] ** 2) ** 0.5
        if sigmas[i] == float("inf"):
            den = model(noise, sig_hat)
            x = den + sigmas[i + 1] * (gamma + 1) * noise
            continue
        den = model(x * get_scalings(sig_hat)[1], sig_hat)
        d = (x - den) / sig_hat
        dt = sigmas[i + 1] - sig_hat
        if sigmas[i + 1] == 0:
            x = x + d * dt
        else:
            x2 = x + d * dt
            den2 = model(x2 * get_scalings(sigmas[i + 1])[1], sigmas[i + 1])
            d2 = (x2 - den2) / sigmas[i + 1]
            x = x + 0.5 * (d + d2) * dt
    return x


@torch.no_grad()
def sample_dpmpp_2m_sde(noise, model, sigmas, eta=1.0, s_noise=1.0, solver_type="midpoint",
                        show_progress=True, variant_info=None, noise_sampler=None):
    """DPM-Solver++(2M) SDE (ref :144-203).

    ``model(x_scaled, sigma, variant_info=...)`` returns the x0 prediction;
    the solver keeps the k-diffusion state x (fp32) and the previous x0.
    """
    if solver_type not in ("heun", "midpoint"):
        raise AssertionError(solver_type)
    x = noise * sigmas[0]
    if noise_sampler is None:
        pos = sigmas[sigmas > 0]
        noise_sampler = IntervalNoiseSampler(x, pos.min(), sigmas[sigmas < float("inf")].max())
    prev_x0, prev_h = None, None
    for i in range(len(sigmas) - 1):
        logger.info(f"step: {i}")
        s_cur, s_next = sigmas[i], sigmas[i + 1]
        if s_cur == float("inf"):
            # Euler step from pure noise; does not update the 2M history (ref :166-169)
            x0 = model(noise, s_cur, variant_info=variant_info)
            x = x0 + s_next * noise
            continue
        else:
            x0 = model(x * get_scalings(s_cur)[1], s_cur, variant_info=variant_info)
            if s_next == 0:
                x = x0
                h = None
            else:
                lam_cur, lam_next = -s_cur.log(), -s_next.log()
                h = lam_next - lam_cur
                eh = eta * h
                c_x0 = -torch.expm1(-h - eh)
                x = (s_next / s_cur) * torch.exp(-eh) * x + c_x0 * x0
                if prev_x0 is not None:
                    r = prev_h / h
                    if solver_type == "heun":
                        x = x + (c_x0 / (-h - eh) + 1.0) * (1.0 / r) * (x0 - prev_x0)
                    else:
                        x = x + 0.5 * c_x0 * (1.0 / r) * (x0 - prev_x0)
                x = x + noise_sampler(s_cur, s_next) * s_next * (-torch.expm1(-2.0 * eh)).sqrt() * s_noise
        prev_x0 = x0
        prev_h = h if h is not None else prev_h
    if variant_info is not None and variant_info.get("type") == "variant1":
        x_long, x_short = x.chunk(2, dim=0)
        x = x_long * (1 - variant_info["alpha"]) + x_short * variant_info["alpha"]
    return x
