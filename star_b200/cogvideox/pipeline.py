"""Latents-to-frames glue of STAR's CogVideoX path (cogvideox-based/sat/sample_sr.py:186-230): sampler -> (b t c h w -> b c t h w)
-> 1 / scale_factor -> chunked 3-D VAE decode -> clamp((x + 1) / 2).  Inputs are what the out-of-scope stages produce: the T5 text
embeddings (cond / uc) and the VAE-encoded LQ clip (``lq_latent``, already multiplied by scale_factor as encode_first_stage
does, diffusion_video.py:188-213)."""
import torch

from .sampling import VPSDEDPMPP2MSampler, sample_sr_latent


@torch.no_grad()
def sample_sr(network, decoder, cond, uc, lq_latent, scale_factor=0.7, num_steps=50, seed=None, sampler=None):
    """network: DiffusionTransformer; decoder: ContextParallelDecoder3D; lq_latent (1, T, 16, h, w).
    Returns (frames (1, F, 3, 8h, 8w) fp32 in [0, 1] -- the `samples` tensor of sample_sr.py:231 before the colour fix --, latent)."""
    sampler = sampler or VPSDEDPMPP2MSampler(num_steps=num_steps, dtype=getattr(network, "dtype", torch.bfloat16))
    samples_z = sample_sr_latent(network, sampler, cond, uc, lq_latent, generator_seed=seed)     # (1, T, 16, h, w)
    latent = (1.0 / scale_factor) * samples_z.permute(0, 2, 1, 3, 4).contiguous()                # (1, 16, T, h, w)
    recon = decoder.decode_latent(latent).to(torch.float32)                                      # (1, 3, F, H, W)
    samples_x = recon.permute(0, 2, 1, 3, 4).contiguous()
    return torch.clamp((samples_x + 1.0) / 2.0, min=0.0, max=1.0), samples_z
