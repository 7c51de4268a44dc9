"""TEST INFRASTRUCTURE -- CPU/fp32 restatement of the temporal VAE the reference calls (rows a3 / a16).

PARITY UNPINNED.  STAR's own code on this path is only the call sites
    video_to_video/video_to_video_model.py:57-63    AutoencoderKLTemporalDecoder.from_pretrained(...)
    video_to_video/video_to_video_model.py:141-151  temporal_vae_decode / vae_decode_chunk  (3-frame windows, z / 0.18215)
    video_to_video/video_to_video_model.py:153-161  vae_encode  (per frame, latent_dist.sample() * 0.18215)
The network itself lives in diffusers==0.30.0 (requirements.txt:14), which is neither vendored in
/root/reference nor installed in this image, and the reference ships no test or golden vector for it.
What follows restates the PUBLISHED architecture of diffusers 0.30.0
    models/autoencoders/autoencoder_kl_temporal_decoder.py   AutoencoderKLTemporalDecoder, TemporalDecoder
    models/autoencoders/vae.py                               Encoder, DiagonalGaussianDistribution
    models/unets/unet_2d_blocks.py                           DownEncoderBlock2D, UNetMidBlock2D
    models/unets/unet_3d_blocks.py                           MidBlockTemporalDecoder, UpBlockTemporalDecoder
    models/resnet.py                                         ResnetBlock2D, TemporalResnetBlock, SpatioTemporalResBlock,
                                                             AlphaBlender, Downsample2D (pad (0,1,0,1)), Upsample2D (nearest)
    models/attention_processor.py                            Attention (1 head, GroupNorm, residual)
with the stabilityai/stable-video-diffusion-img2vid `vae/config.json` hyper-parameters as the defaults
(block_out_channels 128/256/512/512, layers_per_block 2, latent_channels 4, scaling_factor 0.18215), in the state-dict
key layout of that checkpoint, so that real weights load into both this oracle and the CUDA path.
Once diffusers is importable the pin is one call: compare `decode`/`encode_moments` below with the real module on the
synthetic state dict (tests/test_vae.py::test_oracle_matches_diffusers is skipped until then).
"""
from dataclasses import dataclass

import torch
import torch.nn.functional as F


@dataclass
class VaeCfg:
    block_out_channels: tuple = (128, 256, 512, 512)
    layers_per_block: int = 2
    latent_channels: int = 4
    in_channels: int = 3
    out_channels: int = 3
    groups: int = 32
    scaling_factor: float = 0.18215


def _resnet_keys(m, p, cin, cout, temporal=False):
    k = (3, 1, 1) if temporal else (3, 3)
    m[p + ".norm1.weight"] = (cin,); m[p + ".norm1.bias"] = (cin,)
    m[p + ".conv1.weight"] = (cout, cin) + k; m[p + ".conv1.bias"] = (cout,)
    m[p + ".norm2.weight"] = (cout,); m[p + ".norm2.bias"] = (cout,)
    m[p + ".conv2.weight"] = (cout, cout) + k; m[p + ".conv2.bias"] = (cout,)
    if cin != cout:
        m[p + ".conv_shortcut.weight"] = (cout, cin) + ((1, 1, 1) if temporal else (1, 1))
        m[p + ".conv_shortcut.bias"] = (cout,)


def _attn_keys(m, p, c):
    m[p + ".group_norm.weight"] = (c,); m[p + ".group_norm.bias"] = (c,)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        m[f"{p}.{n}.weight"] = (c, c); m[f"{p}.{n}.bias"] = (c,)


def _st_resnet_keys(m, p, cin, cout):
    _resnet_keys(m, p + ".spatial_res_block", cin, cout)
    _resnet_keys(m, p + ".temporal_res_block", cout, cout, temporal=True)
    m[p + ".time_mixer.mix_factor"] = (1,)


def vae_manifest(cfg=VaeCfg()):
    """{key: shape} in the diffusers checkpoint layout (insertion order = module order)."""
    ch, L = cfg.block_out_channels, cfg.layers_per_block
    m = {}
    m["encoder.conv_in.weight"] = (ch[0], cfg.in_channels, 3, 3); m["encoder.conv_in.bias"] = (ch[0],)
    cin = ch[0]
    for i, cout in enumerate(ch):
        for j in range(L):
            _resnet_keys(m, f"encoder.down_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout)
        if i < len(ch) - 1:
            m[f"encoder.down_blocks.{i}.downsamplers.0.conv.weight"] = (cout, cout, 3, 3)
            m[f"encoder.down_blocks.{i}.downsamplers.0.conv.bias"] = (cout,)
        cin = cout
    _attn_keys(m, "encoder.mid_block.attentions.0", ch[-1])
    for j in range(2):
        _resnet_keys(m, f"encoder.mid_block.resnets.{j}", ch[-1], ch[-1])
    m["encoder.conv_norm_out.weight"] = (ch[-1],); m["encoder.conv_norm_out.bias"] = (ch[-1],)
    m["encoder.conv_out.weight"] = (2 * cfg.latent_channels, ch[-1], 3, 3); m["encoder.conv_out.bias"] = (2 * cfg.latent_channels,)
    m["quant_conv.weight"] = (2 * cfg.latent_channels,) * 2 + (1, 1); m["quant_conv.bias"] = (2 * cfg.latent_channels,)
    # temporal decoder
    m["decoder.conv_in.weight"] = (ch[-1], cfg.latent_channels, 3, 3); m["decoder.conv_in.bias"] = (ch[-1],)
    _attn_keys(m, "decoder.mid_block.attentions.0", ch[-1])
    for j in range(L):
        _st_resnet_keys(m, f"decoder.mid_block.resnets.{j}", ch[-1], ch[-1])
    rev = tuple(reversed(ch))
    cin = rev[0]
    for i, cout in enumerate(rev):
        for j in range(L + 1):
            _st_resnet_keys(m, f"decoder.up_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout)
        if i < len(rev) - 1:
            m[f"decoder.up_blocks.{i}.upsamplers.0.conv.weight"] = (cout, cout, 3, 3)
            m[f"decoder.up_blocks.{i}.upsamplers.0.conv.bias"] = (cout,)
        cin = cout
    m["decoder.conv_norm_out.weight"] = (ch[0],); m["decoder.conv_norm_out.bias"] = (ch[0],)
    m["decoder.conv_out.weight"] = (cfg.out_channels, ch[0], 3, 3); m["decoder.conv_out.bias"] = (cfg.out_channels,)
    m["decoder.time_conv_out.weight"] = (cfg.out_channels, cfg.out_channels, 3, 1, 1)
    m["decoder.time_conv_out.bias"] = (cfg.out_channels,)
    return m


# ------------------------------------------------------------------------------------------ building blocks
def _gn(x, sd, p, cfg, eps):
    return F.group_norm(x, cfg.groups, sd[p + ".weight"], sd[p + ".bias"], eps)


def _resnet2d(sd, p, x, cfg, eps=1e-6):
    """ResnetBlock2D, no time embedding, output_scale_factor 1."""
    h = F.conv2d(F.silu(_gn(x, sd, p + ".norm1", cfg, eps)), sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], padding=1)
    h = F.conv2d(F.silu(_gn(h, sd, p + ".norm2", cfg, eps)), sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], padding=1)
    if p + ".conv_shortcut.weight" in sd:
        x = F.conv2d(x, sd[p + ".conv_shortcut.weight"], sd[p + ".conv_shortcut.bias"])
    return x + h


def _temporal_resnet(sd, p, x, cfg, eps=1e-5):
    """TemporalResnetBlock on (b, c, t, h, w): GroupNorm statistics span all frames; Conv3d kernel (3,1,1)."""
    h = F.conv3d(F.silu(_gn(x, sd, p + ".norm1", cfg, eps)), sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], padding=(1, 0, 0))
    h = F.conv3d(F.silu(_gn(h, sd, p + ".norm2", cfg, eps)), sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], padding=(1, 0, 0))
    return x + h


def _st_resnet(sd, p, x, num_frames, cfg):
    """SpatioTemporalResBlock with AlphaBlender('learned', switch_spatial_to_temporal_mix=True)."""
    x = _resnet2d(sd, p + ".spatial_res_block", x, cfg, 1e-6)
    bf, c, h, w = x.shape
    x5 = x.reshape(bf // num_frames, num_frames, c, h, w).permute(0, 2, 1, 3, 4)
    xt = _temporal_resnet(sd, p + ".temporal_res_block", x5, cfg, 1e-5)
    alpha = 1.0 - torch.sigmoid(sd[p + ".time_mixer.mix_factor"].float())
    out = alpha * x5 + (1.0 - alpha) * xt
    return out.permute(0, 2, 1, 3, 4).reshape(bf, c, h, w)


def _attention(sd, p, x, cfg):
    """Attention(heads=1, dim_head=C, norm_num_groups=32, residual_connection=True) on (b, c, h, w)."""
    b, c, h, w = x.shape
    t = _gn(x.reshape(b, c, h * w), sd, p + ".group_norm", cfg, 1e-6).transpose(1, 2)          # b, hw, c
    q = F.linear(t, sd[p + ".to_q.weight"], sd[p + ".to_q.bias"])
    k = F.linear(t, sd[p + ".to_k.weight"], sd[p + ".to_k.bias"])
    v = F.linear(t, sd[p + ".to_v.weight"], sd[p + ".to_v.bias"])
    a = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]
    a = F.linear(a, sd[p + ".to_out.0.weight"], sd[p + ".to_out.0.bias"])
    return x + a.transpose(1, 2).reshape(b, c, h, w)


# ------------------------------------------------------------------------------------------ encoder
@torch.no_grad()
def encode_moments(sd, x, cfg=VaeCfg()):
    """x (n, 3, H, W) in [-1, 1] -> moments (n, 2*latent, H/8, W/8) = quant_conv(Encoder(x)); mean | logvar."""
    ch, L = cfg.block_out_channels, cfg.layers_per_block
    h = F.conv2d(x, sd["encoder.conv_in.weight"], sd["encoder.conv_in.bias"], padding=1)
    for i in range(len(ch)):
        for j in range(L):
            h = _resnet2d(sd, f"encoder.down_blocks.{i}.resnets.{j}", h, cfg)
        if i < len(ch) - 1:                                                     # Downsample2D(padding=0): pad right/bottom by 1
            p = f"encoder.down_blocks.{i}.downsamplers.0.conv"
            h = F.conv2d(F.pad(h, (0, 1, 0, 1)), sd[p + ".weight"], sd[p + ".bias"], stride=2)
    h = _resnet2d(sd, "encoder.mid_block.resnets.0", h, cfg)
    h = _attention(sd, "encoder.mid_block.attentions.0", h, cfg)
    h = _resnet2d(sd, "encoder.mid_block.resnets.1", h, cfg)
    h = F.silu(_gn(h, sd, "encoder.conv_norm_out", cfg, 1e-6))
    h = F.conv2d(h, sd["encoder.conv_out.weight"], sd["encoder.conv_out.bias"], padding=1)
    return F.conv2d(h, sd["quant_conv.weight"], sd["quant_conv.bias"])


def sample_posterior(moments, noise):
    """DiagonalGaussianDistribution.sample(): mean + exp(0.5 * clamp(logvar, -30, 20)) * noise."""
    mean, logvar = moments.chunk(2, dim=1)
    return mean + torch.exp(0.5 * logvar.clamp(-30.0, 20.0)) * noise


# ------------------------------------------------------------------------------------------ temporal decoder
@torch.no_grad()
def decode(sd, z, num_frames, cfg=VaeCfg()):
    """z (b*num_frames, 4, h, w) (already divided by scaling_factor) -> (b*num_frames, 3, 8h, 8w)."""
    ch, L = cfg.block_out_channels, cfg.layers_per_block
    h = F.conv2d(z, sd["decoder.conv_in.weight"], sd["decoder.conv_in.bias"], padding=1)
    h = _st_resnet(sd, "decoder.mid_block.resnets.0", h, num_frames, cfg)
    for j in range(1, L):
        h = _attention(sd, "decoder.mid_block.attentions.0", h, cfg)
        h = _st_resnet(sd, f"decoder.mid_block.resnets.{j}", h, num_frames, cfg)
    for i in range(len(ch)):
        for j in range(L + 1):
            h = _st_resnet(sd, f"decoder.up_blocks.{i}.resnets.{j}", h, num_frames, cfg)
        if i < len(ch) - 1:
            p = f"decoder.up_blocks.{i}.upsamplers.0.conv"
            h = F.conv2d(F.interpolate(h, scale_factor=2.0, mode="nearest"), sd[p + ".weight"], sd[p + ".bias"], padding=1)
    h = F.silu(_gn(h, sd, "decoder.conv_norm_out", cfg, 1e-6))
    h = F.conv2d(h, sd["decoder.conv_out.weight"], sd["decoder.conv_out.bias"], padding=1)
    bf, c, hh, ww = h.shape
    h5 = h.reshape(bf // num_frames, num_frames, c, hh, ww).permute(0, 2, 1, 3, 4)
    h5 = F.conv3d(h5, sd["decoder.time_conv_out.weight"], sd["decoder.time_conv_out.bias"], padding=(1, 0, 0))
    return h5.permute(0, 2, 1, 3, 4).reshape(bf, c, hh, ww)
