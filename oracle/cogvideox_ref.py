"""TEST INFRASTRUCTURE -- CPU restatement of ONE CogVideoX-5B DiT layer with STAR's LIEM gates.

PARITY UNPINNED.  The layer's control flow is STAR's own code and is restated line by line from
    cogvideox-based/sat/dit_video_concat.py:482-563   AdaLNMixin.layer_forward
    cogvideox-based/sat/dit_video_concat.py:570-598   AdaLNMixin.attention_fn      (qk-LayerNorm)
    cogvideox-based/sat/dit_video_concat.py:254-346   Rotary3DPositionEmbeddingMixin (tables, rotary, attention_fn)
    cogvideox-based/transformer.py:316-348            SpatialAttention / TemporalLocalAttention (LIEM)
but the leaf modules it calls (fused QKV ColumnParallelLinear, attention_fn_default, RowParallelLinear
dense, MLP with gelu, LayerNorm) live in SwissArmyTransformer==0.4.12 (cogvideox-based/sat/requirements.txt:1),
which is neither vendored in /root/reference nor installed here, and the reference ships no test or golden
vector for them.  Their semantics below follow sat's published defaults: q|k|v = three contiguous thirds of
one Linear with bias (stride 3, transformer.py:62-76), heads split as (b, heads, s, 64), softmax(QK^T/8)V,
dense with bias, MLP = Linear -> gelu (tanh approximation) -> Linear, LayerNorm with affine.
"""
from dataclasses import dataclass

import torch
import torch.nn.functional as F


@dataclass
class DiTCfg:
    hidden: int = 3072
    heads: int = 48
    head_dim: int = 64
    mlp_ratio: int = 4
    text_length: int = 226
    frames: int = 13            # compressed latent frames (49 video frames)
    height: int = 30            # latent 60x90 patchified by 2
    width: int = 45
    ln_eps: float = 1e-5
    qk_ln_eps: float = 1e-6     # dit_video_concat.py:471


def layer_manifest(cfg):
    """{key: shape} of one layer (names follow the reference's module attributes)."""
    h, hd = cfg.hidden, cfg.head_dim
    m = {
        "adaLN_modulation.1.weight": (12 * h, 512), "adaLN_modulation.1.bias": (12 * h,),        # time_embed_dim 512
        "input_layernorm.weight": (h,), "input_layernorm.bias": (h,),
        "post_attention_layernorm.weight": (h,), "post_attention_layernorm.bias": (h,),
        "spa_local.conv1.weight": (1, 2, 7, 7), "temp_local.conv1.weight": (1, 2),
        "attention.query_key_value.weight": (3 * h, h), "attention.query_key_value.bias": (3 * h,),
        "attention.dense.weight": (h, h), "attention.dense.bias": (h,),
        "query_layernorm.weight": (hd,), "query_layernorm.bias": (hd,),
        "key_layernorm.weight": (hd,), "key_layernorm.bias": (hd,),
        "mlp.dense_h_to_4h.weight": (cfg.mlp_ratio * h, h), "mlp.dense_h_to_4h.bias": (cfg.mlp_ratio * h,),
        "mlp.dense_4h_to_h.weight": (h, cfg.mlp_ratio * h), "mlp.dense_4h_to_h.bias": (h,),
    }
    return m


def rope_tables(cfg, theta=10000.0):
    """freqs_cos / freqs_sin of shape (t*h*w, head_dim) (dit_video_concat.py:269-297)."""
    hd = cfg.head_dim
    dim_t, dim_h, dim_w = hd // 4, hd // 8 * 3, hd // 8 * 3

    def axis(n, dim):
        fr = 1.0 / (theta ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))
        f = torch.arange(n, dtype=torch.float32)[:, None] * fr[None, :]
        return f.repeat_interleave(2, dim=-1)                     # "... n -> ... (n r)", r=2

    ft, fh, fw = axis(cfg.frames, dim_t), axis(cfg.height, dim_h), axis(cfg.width, dim_w)
    T, H, W = cfg.frames, cfg.height, cfg.width
    freqs = torch.cat([ft[:, None, None, :].expand(T, H, W, -1), fh[None, :, None, :].expand(T, H, W, -1),
                       fw[None, None, :, :].expand(T, H, W, -1)], dim=-1).reshape(T * H * W, hd)
    return freqs.cos(), freqs.sin()


def _rotate_half(x):
    """interleaved pairs (x1, x2) -> (-x2, x1) (sat's rotate_half on '... (d r)', r=2)."""
    x = x.reshape(*x.shape[:-1], -1, 2)
    x1, x2 = x.unbind(-1)
    return torch.stack((-x2, x1), dim=-1).reshape(*x.shape[:-2], -1)


def _modulate(x, shift, scale):
    return x * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1)


@torch.no_grad()
def dit_layer_forward(sd, hidden, emb, cfg, cos=None, sin=None):
    """hidden: (b, text_length + t*h*w, hidden); emb: (b, 512) timestep embedding.  Returns the same shape."""
    if cos is None:
        cos, sin = rope_tables(cfg)
    b, _, d = hidden.shape
    tl, T, H, W = cfg.text_length, cfg.frames, cfg.height, cfg.width
    txt, img = hidden[:, :tl], hidden[:, tl:]
    mod = F.linear(F.silu(emb), sd["adaLN_modulation.1.weight"], sd["adaLN_modulation.1.bias"])      # SiLU -> Linear
    (sh_msa, sc_msa, g_msa, sh_mlp, sc_mlp, g_mlp, tsh_msa, tsc_msa, tg_msa, tsh_mlp, tsc_mlp, tg_mlp) = mod.chunk(12, dim=1)

    def ln(x, p):
        return F.layer_norm(x, (d,), sd[p + ".weight"], sd[p + ".bias"], cfg.ln_eps)

    img_in = _modulate(ln(img, "input_layernorm"), sh_msa, sc_msa)                                   # :518-521
    txt_in = _modulate(ln(txt, "input_layernorm"), tsh_msa, tsc_msa)
    # spatial LIEM on (b t) c h w                                                                     # :523-527
    spa = img_in.reshape(b * T, H, W, d).permute(0, 3, 1, 2)
    wgt = torch.cat([spa.max(dim=1, keepdim=True)[0], spa.mean(dim=1, keepdim=True)], dim=1)
    spa = torch.sigmoid(F.conv2d(wgt, sd["spa_local.conv1.weight"], padding=3)) * spa
    # temporal LIEM on (b h w) t c                                                                    # :529-531
    tmp = spa.permute(0, 2, 3, 1).reshape(b, T, H * W, d)
    g = torch.sigmoid(F.linear(torch.cat([tmp.max(dim=-1, keepdim=True)[0], tmp.mean(dim=-1, keepdim=True)], -1),
                               sd["temp_local.conv1.weight"]))
    img_in = (g * tmp).reshape(b, T * H * W, d)
    x = torch.cat([txt_in, img_in], dim=1)                                                           # :535
    # attention (sat default attention_forward + the two attention_fn mixins)
    qkv = F.linear(x, sd["attention.query_key_value.weight"], sd["attention.query_key_value.bias"])
    q, k, v = qkv.chunk(3, dim=-1)

    def heads(t):
        return t.reshape(b, -1, cfg.heads, cfg.head_dim).permute(0, 2, 1, 3)

    q, k, v = heads(q), heads(k), heads(v)
    q = F.layer_norm(q, (cfg.head_dim,), sd["query_layernorm.weight"], sd["query_layernorm.bias"], cfg.qk_ln_eps)   # :583-587
    k = F.layer_norm(k, (cfg.head_dim,), sd["key_layernorm.weight"], sd["key_layernorm.bias"], cfg.qk_ln_eps)
    c, s = cos[None, None].to(q), sin[None, None].to(q)
    q = torch.cat([q[:, :, :tl], q[:, :, tl:] * c + _rotate_half(q[:, :, tl:]) * s], dim=2)          # :332-333
    k = torch.cat([k[:, :, :tl], k[:, :, tl:] * c + _rotate_half(k[:, :, tl:]) * s], dim=2)
    a = F.scaled_dot_product_attention(q, k, v)
    a = a.permute(0, 2, 1, 3).reshape(b, -1, d)
    a = F.linear(a, sd["attention.dense.weight"], sd["attention.dense.bias"])
    img = img + g_msa.unsqueeze(1) * a[:, tl:]                                                       # :543-544
    txt = txt + tg_msa.unsqueeze(1) * a[:, :tl]
    # MLP                                                                                            # :546-561
    mi = _modulate(ln(img, "post_attention_layernorm"), sh_mlp, sc_mlp)
    mt = _modulate(ln(txt, "post_attention_layernorm"), tsh_mlp, tsc_mlp)
    m = torch.cat([mt, mi], dim=1)
    m = F.linear(F.gelu(F.linear(m, sd["mlp.dense_h_to_4h.weight"], sd["mlp.dense_h_to_4h.bias"]), approximate="tanh"),
                 sd["mlp.dense_4h_to_h.weight"], sd["mlp.dense_4h_to_h.bias"])
    img = img + g_mlp.unsqueeze(1) * m[:, tl:]
    txt = txt + tg_mlp.unsqueeze(1) * m[:, :tl]
    return torch.cat([txt, img], dim=1)
