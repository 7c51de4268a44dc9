"""Frames-to-frames glue of STAR's CogVideoX path (cogvideox-based/sat/sample_sr.py:186-230 around
``SATVideoDiffusionEngine.sample_sr``, diffusion_video.py:245-292): start noise -> 3-D VAE encode of the (pre-upsampled) LQ clip
-> sampler -> (b t c h w -> b c t h w) -> 1 / scale_factor -> chunked 3-D VAE decode -> clamp((x + 1) / 2).  What stays outside
is what the hot-path contract leaves outside: the T5 text embeddings (``cond`` / ``uc``) and video file I/O."""
import torch

from .sampling import VPSDEDPMPP2MSampler, sample_sr_latent


@torch.no_grad()
def sample_sr(network, decoder, cond, uc, lq_latent=None, lq=None, encoder=None, scale_factor=0.7, num_steps=50, seed=None,
              sampler=None):
    """network: DiffusionTransformer; decoder / encoder: ContextParallel{Decoder,Encoder}3D.  Give either ``lq`` (1, F, 3, H, W) in
    [-1, 1] with ``encoder`` (the reference's path: posterior SAMPLE of the encoded clip times scale_factor,
    diffusion_video.py:279-282) or a ready ``lq_latent`` (1, T, 16, H/8, W/8).
    Returns (frames (1, F, 3, H, W) fp32 in [0, 1] -- the `samples` tensor of sample_sr.py:231 before the colour fix --, latent).
    Random draws happen in the reference's order: start noise on the host, posterior sample, solver noise."""
    dtype = getattr(network, "dtype", torch.bfloat16)
    sampler = sampler or VPSDEDPMPP2MSampler(num_steps=num_steps, dtype=dtype)
    if seed is not None:
        torch.manual_seed(seed)
    if lq_latent is None:
        if lq is None or encoder is None:
            raise ValueError("sample_sr needs either lq_latent or (lq, encoder)")
        _, F, _, H, W = lq.shape
        shape, dev = (1, (F - 1) // 4 + 1, 16, H // 8, W // 8), lq.device
    else:
        shape, dev = tuple(lq_latent.shape), lq_latent.device
    randn = torch.randn(shape, dtype=torch.float32).to(dev)
    if lq_latent is None:
        z = encoder.encode(lq.to(dtype).permute(0, 2, 1, 3, 4).contiguous())                    # (1, 16, T, h, w)
        lq_latent = (scale_factor * z).permute(0, 2, 1, 3, 4).contiguous()
    samples_z = sample_sr_latent(network, sampler, cond, uc, lq_latent, randn=randn)             # (1, T, 16, h, w)
    latent = (1.0 / scale_factor) * samples_z.permute(0, 2, 1, 3, 4).contiguous()                # (1, 16, T, h, w)
    recon = decoder.decode_latent(latent).to(torch.float32)                                      # (1, 3, F, H, W)
    samples_x = recon.permute(0, 2, 1, 3, 4).contiguous()
    return torch.clamp((samples_x + 1.0) / 2.0, min=0.0, max=1.0), samples_z
