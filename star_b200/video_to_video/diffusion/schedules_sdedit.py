"""Noise schedules of the STAR sampler (host side, fp32).

Mirrors the public functions of the reference's
video_to_video/diffusion/schedules_sdedit.py (names and argument meaning):
``noise_schedule`` (:72-85) is the only one the hot path calls
(video_to_video_model.py:46-52); the converters and ``karras_schedule`` (:54)
are kept for API compatibility.
"""
import math

import torch

__all__ = [
    "betas_to_sigmas", "sigmas_to_betas", "logsnrs_to_sigmas", "sigmas_to_logsnrs",
    "karras_schedule", "logsnr_cosine_interp_schedule", "noise_schedule",
]


def betas_to_sigmas(betas):
    return (1.0 - torch.cumprod(1.0 - betas, dim=0)).sqrt()


def sigmas_to_betas(sigmas):
    a2 = 1.0 - sigmas ** 2
    return 1.0 - torch.cat([a2[:1], a2[1:] / a2[:-1]])


def logsnrs_to_sigmas(logsnrs):
    return torch.sigmoid(-logsnrs).sqrt()


def sigmas_to_logsnrs(sigmas):
    s2 = sigmas ** 2
    return torch.log(s2 / (1.0 - s2))


def _cosine_logsnr(n, logsnr_min, logsnr_max, scale):
    """log-SNR of the cosine schedule shifted by 2*log(1/scale) (ref :28-39)."""
    lo = math.atan(math.exp(-0.5 * logsnr_min))
    hi = math.atan(math.exp(-0.5 * logsnr_max))
    u = torch.linspace(1, 0, n)
    out = -2.0 * torch.log(torch.tan(lo + u * (hi - lo)))
    out += 2.0 * math.log(1.0 / scale)
    return out


def logsnr_cosine_interp_schedule(n, logsnr_min=-15, logsnr_max=15, scale_min=2, scale_max=4):
    """Interpolates between the scale_min- and scale_max-shifted cosine
    schedules along t (ref :42-51, :61-69)."""
    u = torch.linspace(1, 0, n)
    a = _cosine_logsnr(n, logsnr_min, logsnr_max, scale_min)
    b = _cosine_logsnr(n, logsnr_min, logsnr_max, scale_max)
    return logsnrs_to_sigmas(u * a + (1 - u) * b)


def karras_schedule(n, sigma_min=0.002, sigma_max=80.0, rho=7.0):
    ramp = torch.linspace(1, 0, n)
    lo, hi = sigma_min ** (1 / rho), sigma_max ** (1 / rho)
    s = (hi + ramp * (lo - hi)) ** rho
    return torch.sqrt(s ** 2 / (1 + s ** 2))


_SCHEDULES = {"logsnr_cosine_interp": logsnr_cosine_interp_schedule}


def noise_schedule(schedule="logsnr_cosine_interp", n=1000, zero_terminal_snr=False, **kwargs):
    """sigma table of length n; with zero_terminal_snr the table is affinely
    stretched so that max(sigma) == 1 (ref :72-85)."""
    sigmas = _SCHEDULES[schedule](n, **kwargs)
    if zero_terminal_snr and sigmas.max() != 1.0:
        lo = sigmas.min()
        sigmas = lo + (1.0 - lo) / (sigmas.max() - lo) * (sigmas - lo)
    return sigmas
