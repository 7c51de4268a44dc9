"""Noise schedules of the STAR sampler (host side, fp32).

Mirrors the public functions of the reference's
video_to_video/diffusion/schedules_sdedit.py (names and argument meaning):
``noise_schedule`` (:72-85) with the ``logsnr_cosine_interp`` schedule is the only one the hot path
calls (video_to_video_model.py:46-52).  The reference's other converters and ``karras_schedule`` (:54) are
unreachable from STAR's sampler configuration and are not carried over.
"""
import math

import torch

__all__ = ["logsnrs_to_sigmas", "logsnr_cosine_interp_schedule", "noise_schedule"]


def logsnrs_to_sigmas(logsnrs):
    return torch.sigmoid(-logsnrs).sqrt()


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


_SCHEDULES = {"logsnr_cosine_interp": logsnr_cosine_interp_schedule}


def noise_schedule(schedule="logsnr_cosine_interp", n=1000, zero_terminal_snr=False, **kwargs):
    """sigma table of length n; with zero_terminal_snr the table is affinely
    stretched so that max(sigma) == 1 (ref :72-85)."""
    sigmas = _SCHEDULES[schedule](n, **kwargs)
    if zero_terminal_snr and sigmas.max() != 1.0:
        lo = sigmas.min()
        sigmas = lo + (1.0 - lo) / (sigmas.max() - lo) * (sigmas - lo)
    return sigmas
