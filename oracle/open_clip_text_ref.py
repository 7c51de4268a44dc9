"""TEST INFRASTRUCTURE -- fp32 restatement of the OpenCLIP ViT-H-14 text transformer as FrozenOpenCLIPEmbedder runs it
(video_to_video/modules/embedder.py:52-71).

PARITY UNPINNED: the arithmetic lives in the un-vendored open-clip-torch==2.20.0 (requirements.txt:9), absent from
/root/reference and from this image.  Restated from its published model code (open_clip/transformer.py
``ResidualAttentionBlock`` / ``Transformer`` / ``TextTransformer``, open_clip/model.py ``CLIP.encode_text``):
    x = token_embedding(text) + positional_embedding
    for each block but the last `skip_last`:   x = x + MHA(ln_1(x), attn_mask = -inf above the diagonal)
                                               x = x + c_proj(GELU(c_fc(ln_2(x))))
    x = ln_final(x)
The leaves are torch's own nn.MultiheadAttention / nn.LayerNorm / nn.GELU(erf) -- exactly the modules open_clip
instantiates -- so only the block wiring above is restated.  Never imported by the product.
"""
import torch
import torch.nn as nn


def text_manifest(width=1024, layers=24, vocab=49408, context=77):
    m = {"token_embedding.weight": (vocab, width), "positional_embedding": (context, width),
         "ln_final.weight": (width,), "ln_final.bias": (width,)}
    for i in range(layers):
        p = f"transformer.resblocks.{i}."
        m.update({p + "ln_1.weight": (width,), p + "ln_1.bias": (width,), p + "ln_2.weight": (width,), p + "ln_2.bias": (width,),
                  p + "attn.in_proj_weight": (3 * width, width), p + "attn.in_proj_bias": (3 * width,),
                  p + "attn.out_proj.weight": (width, width), p + "attn.out_proj.bias": (width,),
                  p + "mlp.c_fc.weight": (4 * width, width), p + "mlp.c_fc.bias": (4 * width,),
                  p + "mlp.c_proj.weight": (width, 4 * width), p + "mlp.c_proj.bias": (width,)})
    return m


@torch.no_grad()
def encode_with_transformer(sd, tokens, heads, skip_last=1):
    """sd: open_clip key names (fp32); tokens (B, N) int64 -> (B, N, width) fp32"""
    width = sd["positional_embedding"].shape[1]
    layers = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("transformer.resblocks."))
    dev = tokens.device
    N = tokens.shape[1]
    x = sd["token_embedding.weight"].float()[tokens] + sd["positional_embedding"].float()[:N]
    x = x.permute(1, 0, 2)                                            # (N, B, C) as open_clip runs it
    mask = torch.full((N, N), float("-inf"), device=dev).triu_(1)
    for i in range(layers - skip_last):
        p = f"transformer.resblocks.{i}."
        ln1, ln2 = nn.LayerNorm(width).to(dev), nn.LayerNorm(width).to(dev)
        mha = nn.MultiheadAttention(width, heads).to(dev)
        ln1.load_state_dict({"weight": sd[p + "ln_1.weight"].float(), "bias": sd[p + "ln_1.bias"].float()})
        ln2.load_state_dict({"weight": sd[p + "ln_2.weight"].float(), "bias": sd[p + "ln_2.bias"].float()})
        mha.load_state_dict({"in_proj_weight": sd[p + "attn.in_proj_weight"].float(), "in_proj_bias": sd[p + "attn.in_proj_bias"].float(),
                             "out_proj.weight": sd[p + "attn.out_proj.weight"].float(), "out_proj.bias": sd[p + "attn.out_proj.bias"].float()})
        h = ln1(x)
        x = x + mha(h, h, h, need_weights=False, attn_mask=mask)[0]
        h = ln2(x)
        h = nn.functional.gelu(nn.functional.linear(h, sd[p + "mlp.c_fc.weight"].float(), sd[p + "mlp.c_fc.bias"].float()))
        x = x + nn.functional.linear(h, sd[p + "mlp.c_proj.weight"].float(), sd[p + "mlp.c_proj.bias"].float())
    x = x.permute(1, 0, 2)
    return nn.functional.layer_norm(x, (width,), sd["ln_final.weight"].float(), sd["ln_final.bias"].float())
