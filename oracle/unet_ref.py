"""TEST INFRASTRUCTURE -- CPU restatement of the STAR denoiser forward.

Plain-torch (NCHW, functional) restatement of
    ControlledV2VUNet.forward      video_to_video/modules/unet_v2v.py:1717-1809
    VideoControlNet.forward        video_to_video/modules/unet_v2v.py:2134-2206
driven directly by a reference-layout ``state_dict`` (no nn.Module tree, no
code shared with ``star_b200``).  Precision is whatever dtype the state_dict /
inputs carry (fp32 for the oracle).  Pinned against the real reference modules
in tests/test_oracle_pinning.py (runs where /root/reference exists) and
against tests/golden/*.pt everywhere.

Each helper cites the reference lines it restates.
"""
import math
from dataclasses import dataclass, field
from typing import List

import torch
import torch.nn.functional as F


@dataclass
class UNetCfg:
    """Constructor defaults of Vid2VidSDUNet / VideoControlNet
    (unet_v2v.py:1283-1303, :1898-1918)."""
    in_dim: int = 4
    dim: int = 320
    context_dim: int = 1024
    out_dim: int = 4
    dim_mult: List[int] = field(default_factory=lambda: [1, 2, 4, 4])
    num_heads: int = 8          # heads of the *initial* temporal transformer only (:1358-1368)
    head_dim: int = 64
    num_res_blocks: int = 2
    attn_scales: List[float] = field(default_factory=lambda: [1.0, 0.5, 0.25])


# ----------------------------------------------------------------------------
# leaf ops
# ----------------------------------------------------------------------------
def sinusoidal_embedding(t, dim):
    """unet_v2v.py:96-108 (cos || sin)."""
    half = dim // 2
    t = t.float()
    freqs = torch.pow(10000, -torch.arange(half).to(t).div(half))
    s = torch.outer(t, freqs)
    return torch.cat([torch.cos(s), torch.sin(s)], dim=1)


def _lin(sd, p, x, bias=True):
    return F.linear(x, sd[p + ".weight"], sd[p + ".bias"] if bias else None)


def _gn(sd, p, x, eps):
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], eps)


def _ln(sd, p, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def _attention(sd, p, x, context, heads, dim_head=64):
    """MemoryEfficientCrossAttention.forward, unet_v2v.py:158-195.
    xformers MEA == softmax(q k^T / sqrt(d)) v; the max_bs chunking (:173-182)
    only splits the batch and does not change values."""
    q = _lin(sd, p + ".to_q", x, bias=False)
    ctx = x if context is None else context
    k = _lin(sd, p + ".to_k", ctx, bias=False)
    v = _lin(sd, p + ".to_v", ctx, bias=False)
    b, n, _ = q.shape

    def split(t):
        return t.reshape(b, t.shape[1], heads, dim_head).permute(0, 2, 1, 3)

    o = F.scaled_dot_product_attention(split(q), split(k), split(v))
    o = o.permute(0, 2, 1, 3).reshape(b, n, heads * dim_head)
    return _lin(sd, p + ".to_out.0", o)


def _ff(sd, p, x):
    """FeedForward with GEGLU, unet_v2v.py:496-529 (exact erf GELU)."""
    h = _lin(sd, p + ".net.0.proj", x)
    a, gate = h.chunk(2, dim=-1)
    return _lin(sd, p + ".net.2", a * F.gelu(gate))


def _liem_spatial(sd, p, x):
    """SpatialAttention (spatial LIEM), unet_v2v.py:380-394; x is (b,c,h,w)."""
    mx = x.max(dim=1, keepdim=True)[0]
    av = x.mean(dim=1, keepdim=True)
    w = F.conv2d(torch.cat([mx, av], 1), sd[p + ".conv1.weight"], padding=3)
    return torch.sigmoid(w) * x


def _liem_temporal(sd, p, x):
    """TemporalLocalAttention, unet_v2v.py:396-411; x is (..., c)."""
    mx = x.max(dim=-1, keepdim=True)[0]
    av = x.mean(dim=-1, keepdim=True)
    w = F.linear(torch.cat([mx, av], -1), sd[p + ".conv1.weight"])
    return torch.sigmoid(w) * x


# ----------------------------------------------------------------------------
# blocks
# ----------------------------------------------------------------------------
def _block_space(sd, p, x, context, heads, h, w):
    """BasicTransformerBlock.forward, local_type='space', unet_v2v.py:466-477."""
    b, n, c = x.shape
    xl = _liem_spatial(sd, p + ".local1", x.transpose(1, 2).reshape(b, c, h, w))
    xl = xl.reshape(b, c, n).transpose(1, 2)
    x = _attention(sd, p + ".attn1", _ln(sd, p + ".norm1", xl), None, heads) + x
    x = _attention(sd, p + ".attn2", _ln(sd, p + ".norm2", x), context, heads) + x
    x = _ff(sd, p + ".ff", _ln(sd, p + ".norm3", x)) + x
    return x


def _block_temp(sd, p, x, heads):
    """BasicTransformerBlock.forward, local_type='temp', unet_v2v.py:479-490
    (attn2 gets context=None => second self-attention)."""
    xl = _liem_temporal(sd, p + ".local1", x)
    x = _attention(sd, p + ".attn1", _ln(sd, p + ".norm1", xl), None, heads) + x
    xl = _liem_temporal(sd, p + ".local2", x)
    x = _attention(sd, p + ".attn2", _ln(sd, p + ".norm2", xl), None, heads) + x
    x = _ff(sd, p + ".ff", _ln(sd, p + ".norm3", x)) + x
    return x


def _spatial_transformer(sd, p, x, context, heads):
    """SpatialTransformer.forward (use_linear=True), unet_v2v.py:297-317."""
    bf, c, h, w = x.shape
    x_in = x
    x = _gn(sd, p + ".norm", x, 1e-6)
    x = x.reshape(bf, c, h * w).transpose(1, 2)
    x = _lin(sd, p + ".proj_in", x)
    x = _block_space(sd, p + ".transformer_blocks.0", x, context, heads, h, w)
    x = _lin(sd, p + ".proj_out", x)
    x = x.transpose(1, 2).reshape(bf, c, h, w)
    return x + x_in


def _temporal_transformer(sd, p, x, batch, heads):
    """TemporalTransformer.forward (use_linear=False, only_self_att=True),
    unet_v2v.py:1034-1092; x is (b f) c h w."""
    bf, c, h, w = x.shape
    f = bf // batch
    x5 = x.reshape(batch, f, c, h, w).permute(0, 2, 1, 3, 4)          # b c f h w (:1839)
    x_in = x5
    xn = F.group_norm(x5, 32, sd[p + ".norm.weight"], sd[p + ".norm.bias"], 1e-6)   # 5-D GN (:1042)
    t = xn.permute(0, 3, 4, 1, 2).reshape(batch * h * w, c, f)         # (b h w) c f (:1045)
    t = F.conv1d(t, sd[p + ".proj_in.weight"], sd[p + ".proj_in.bias"])
    t = t.transpose(1, 2)                                              # (bhw) f c (:1056)
    t = _block_temp(sd, p + ".transformer_blocks.0", t, heads)
    t = t.transpose(1, 2)                                              # (bhw) c f (:1083)
    t = F.conv1d(t, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
    t = t.reshape(batch, h, w, c, f).permute(0, 3, 4, 1, 2)            # b c f h w (:1085)
    out = t + x_in
    return out.permute(0, 2, 1, 3, 4).reshape(bf, c, h, w)


def _temporal_conv(sd, p, x5):
    """TemporalConvBlock_v2.forward, default branch, unet_v2v.py:1266-1277."""
    idn = x5
    for i, conv_idx in ((1, 2), (2, 3), (3, 3), (4, 3)):
        q = "%s.conv%d" % (p, i)
        x5 = F.group_norm(x5, 32, sd[q + ".0.weight"], sd[q + ".0.bias"], 1e-5)
        x5 = F.silu(x5)
        x5 = F.conv3d(x5, sd["%s.%d.weight" % (q, conv_idx)], sd["%s.%d.bias" % (q, conv_idx)], padding=(1, 0, 0))
    return idn + x5


def _resblock(sd, p, x, e, batch):
    """ResBlock._forward (no up/down, use_scale_shift_norm=False), unet_v2v.py:666-692."""
    h = F.conv2d(F.silu(_gn(sd, p + ".in_layers.0", x, 1e-5)),
                 sd[p + ".in_layers.2.weight"], sd[p + ".in_layers.2.bias"], padding=1)
    emb = _lin(sd, p + ".emb_layers.1", F.silu(e))
    h = h + emb[:, :, None, None]
    h = F.conv2d(F.silu(_gn(sd, p + ".out_layers.0", h, 1e-5)),
                 sd[p + ".out_layers.3.weight"], sd[p + ".out_layers.3.bias"], padding=1)
    if (p + ".skip_connection.weight") in sd:
        x = F.conv2d(x, sd[p + ".skip_connection.weight"], sd[p + ".skip_connection.bias"])
    h = x + h
    bf, c, hh, ww = h.shape
    h5 = h.reshape(batch, bf // batch, c, hh, ww).permute(0, 2, 1, 3, 4)
    h5 = _temporal_conv(sd, p + ".temopral_conv", h5)
    return h5.permute(0, 2, 1, 3, 4).reshape(bf, c, hh, ww)


def _downsample(sd, p, x):
    """Downsample.forward: conv3x3 stride 2 padding (2,1), unet_v2v.py:709-729."""
    return F.conv2d(x, sd[p + ".op.weight"], sd[p + ".op.bias"], stride=2, padding=(2, 1))


def _upsample(sd, p, x):
    """Upsample.forward (dims=2.0): nearest x2, crop one row top and bottom,
    conv3x3; unet_v2v.py:556-567."""
    x = F.interpolate(x, scale_factor=2, mode="nearest")
    x = x[..., 1:-1, :]
    return F.conv2d(x, sd[p + ".conv.weight"], sd[p + ".conv.bias"], padding=1)


# ----------------------------------------------------------------------------
# layout of input/middle/output blocks (unet_v2v.py:1351-1547, :1983-2118)
# ----------------------------------------------------------------------------
def encoder_layout(cfg):
    """list of (kind, c_in, c_out, has_attn) for input_blocks[1:]; kind in
    {'res','down'}; also returns shortcut dims and final (channels, scale)."""
    dims = [cfg.dim * u for u in [1] + list(cfg.dim_mult)]
    out, shortcuts, scale = [], [cfg.dim], 1.0
    for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
        for j in range(cfg.num_res_blocks):
            out.append(("res", cin, cout, scale in cfg.attn_scales))
            shortcuts.append(cout)
            cin = cout
            if i != len(cfg.dim_mult) - 1 and j == cfg.num_res_blocks - 1:
                out.append(("down", cout, cout, False))
                shortcuts.append(cout)
                scale /= 2.0
    return out, shortcuts, dims[-1], scale


def decoder_layout(cfg):
    enc, shortcuts, _, scale = encoder_layout(cfg)
    shortcuts = list(shortcuts)
    dims = [cfg.dim * u for u in [cfg.dim_mult[-1]] + list(cfg.dim_mult[::-1])]
    out = []
    for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
        for j in range(cfg.num_res_blocks + 1):
            attn = scale in cfg.attn_scales
            up = (i != len(cfg.dim_mult) - 1 and j == cfg.num_res_blocks)
            out.append((cin + shortcuts.pop(), cout, attn, up))
            cin = cout
            if up:
                scale *= 2.0
    return out


def _time_embed(sd, p, t, dim):
    e = sinusoidal_embedding(t, dim).to(sd[p + ".0.weight"].dtype)
    return _lin(sd, p + ".2", F.silu(_lin(sd, p + ".0", e)))


def _run_encoder(sd, pre, x, e, context, batch, cfg, hint=None, zero_convs=False):
    """Shared by the UNet encoder (unet_v2v.py:1775-1778) and the ControlNet
    encoder (:2186-2198).  Returns (x, list of per-block outputs)."""
    xs = []
    # input_blocks.0 = [conv3x3, TemporalTransformer(dim, num_heads=8 -> inner 512)]
    x = F.conv2d(x, sd[pre + "input_blocks.0.0.weight"], sd[pre + "input_blocks.0.0.bias"], padding=1)
    if hint is not None:
        x = x + hint                                                    # :2191-2194
    x = _temporal_transformer(sd, pre + "input_blocks.0.1", x, batch, cfg.num_heads)

    def emit(k, x):
        if zero_convs:
            xs.append(F.conv2d(x, sd["%szero_convs.%d.0.weight" % (pre, k)], sd["%szero_convs.%d.0.bias" % (pre, k)]))
        else:
            xs.append(x)

    emit(0, x)
    layout, _, _, _ = encoder_layout(cfg)
    for k, (kind, cin, cout, attn) in enumerate(layout, start=1):
        p = "%sinput_blocks.%d" % (pre, k)
        if kind == "down":
            x = _downsample(sd, p, x)
        else:
            x = _resblock(sd, p + ".0", x, e, batch)
            if attn:
                heads = cout // cfg.head_dim
                x = _spatial_transformer(sd, p + ".1", x, context, heads)
                x = _temporal_transformer(sd, p + ".2", x, batch, heads)
        emit(k, x)
    return x, xs


def _run_middle(sd, pre, x, e, context, batch, cfg):
    c = cfg.dim * cfg.dim_mult[-1]
    heads = c // cfg.head_dim
    x = _resblock(sd, pre + "middle_block.0", x, e, batch)
    x = _spatial_transformer(sd, pre + "middle_block.1", x, context, heads)
    x = _temporal_transformer(sd, pre + "middle_block.2", x, batch, heads)
    x = _resblock(sd, pre + "middle_block.3", x, e, batch)
    return x


def controlnet_forward(sd, x, t, y, hint, cfg, pre="VideoControlNet."):
    """VideoControlNet.forward, unet_v2v.py:2134-2206.  Returns the list xs."""
    batch, _, f, h, w = x.shape
    hint4 = hint.permute(0, 2, 1, 3, 4).reshape(batch * f, -1, h, w)
    hint4 = F.conv2d(hint4, sd[pre + "input_hint_block.weight"], sd[pre + "input_hint_block.bias"], padding=1)
    e = _time_embed(sd, pre + "time_embed", t, cfg.dim).repeat_interleave(f, dim=0)
    context = y.repeat_interleave(f, dim=0)
    x4 = x.permute(0, 2, 1, 3, 4).reshape(batch * f, -1, h, w)
    x4, xs = _run_encoder(sd, pre, x4, e, context, batch, cfg, hint=hint4, zero_convs=True)
    x4 = _run_middle(sd, pre, x4, e, context, batch, cfg)
    xs.append(F.conv2d(x4, sd[pre + "middle_block_out.0.weight"], sd[pre + "middle_block_out.0.bias"]))
    return xs


@torch.no_grad()
def controlled_unet_forward(sd, x, t, y, hint, cfg=None):
    """ControlledV2VUNet.forward, unet_v2v.py:1717-1809.
    x, hint: (b,4,f,h,w); t: (b,) long; y: (b,77,1024).  Returns (b,4,f,h,w)."""
    cfg = cfg or UNetCfg()
    batch, _, f, h, w = x.shape
    control = controlnet_forward(sd, x, t, y, hint, cfg)                # :1746
    e = _time_embed(sd, "time_embed", t, cfg.dim).repeat_interleave(f, dim=0)   # :1765-1766
    context = y.repeat_interleave(f, dim=0)                             # :1769
    x4 = x.permute(0, 2, 1, 3, 4).reshape(batch * f, -1, h, w)          # :1772
    x4, xs = _run_encoder(sd, "", x4, e, context, batch, cfg)
    x4 = _run_middle(sd, "", x4, e, context, batch, cfg)
    x4 = control.pop() + x4                                             # :1784-1785
    for k, (cin, cout, attn, up) in enumerate(decoder_layout(cfg)):
        p = "output_blocks.%d" % k
        x4 = torch.cat([x4, xs.pop() + control.pop()], dim=1)           # :1792
        x4 = _resblock(sd, p + ".0", x4, e, batch)
        idx = 1
        if attn:
            heads = cout // cfg.head_dim
            x4 = _spatial_transformer(sd, p + ".1", x4, context, heads)
            x4 = _temporal_transformer(sd, p + ".2", x4, batch, heads)
            idx = 3
        if up:
            x4 = _upsample(sd, "%s.%d" % (p, idx), x4)
    x4 = F.conv2d(F.silu(_gn(sd, "out.0", x4, 1e-5)), sd["out.2.weight"], sd["out.2.bias"], padding=1)   # :1805
    return x4.reshape(batch, f, -1, h, w).permute(0, 2, 1, 3, 4)        # :1808
