"""Text encoder of the STAR pipeline (reference: video_to_video/modules/embedder.py:12-74) -- SURVEY section 8 row f3.

``FrozenOpenCLIPEmbedder`` keeps the reference's constructor / ``forward`` / ``encode`` / ``encode_with_transformer`` contract.
The OpenCLIP ViT-H-14 *text transformer* itself (24 pre-LN residual blocks, width 1024, 16 heads of 64, causal mask; the
reference stops one block early, ``layer='penultimate'``, and applies ``ln_final``) runs on the star_b200 kernels through
``OpenCLIPTextTower``: LayerNorm -> fused in_proj GEMM -> causal attention (``star_attention_causal``) -> out_proj GEMM with
the residual in its epilogue -> LayerNorm -> c_fc GEMM with the erf-GELU in its epilogue -> c_proj GEMM + residual.

What still comes from the un-vendored ``open_clip`` package (open-clip-torch==2.20.0, requirements.txt:9) is data, not
arithmetic: the pretrained weights (``create_model_and_transforms``) and the BPE tokenizer with its vocabulary file
(``open_clip.tokenize``).  Without the package, pass ``state_dict=`` (the CLIP text-side tensors, open_clip key names) and
``tokenizer=`` (str -> LongTensor (B, 77)) instead -- or ``weights_path=`` (a local open_clip_pytorch_model.bin) and ``bpe_path=``
(the package's bpe_simple_vocab_16e6.txt.gz, read by ``clip_tokenizer.SimpleTokenizer``): then nothing of open_clip is imported.
The bench and the parity tests feed text embeddings directly.
"""
import torch
import torch.nn as nn

from ... import ops

__all__ = ["FrozenOpenCLIPEmbedder", "OpenCLIPTextTower"]


class _Block(nn.Module):
    def __init__(self, width, heads):
        super().__init__()
        self.ln_1 = nn.LayerNorm(width)
        self.attn = nn.MultiheadAttention(width, heads)            # parameter holder: in_proj_weight / in_proj_bias / out_proj.*
        self.ln_2 = nn.LayerNorm(width)
        self.mlp = nn.Sequential()
        self.mlp.add_module("c_fc", nn.Linear(width, 4 * width))
        self.mlp.add_module("c_proj", nn.Linear(4 * width, width))


class OpenCLIPTextTower(nn.Module):
    """open_clip ``CLIP`` text side with its key layout (token_embedding.weight, positional_embedding,
    transformer.resblocks.i.{ln_1, attn.in_proj_*, attn.out_proj, ln_2, mlp.c_fc, mlp.c_proj}.*, ln_final.*).
    ``nn`` only stores the parameters; tokens are a [B*77, width] matrix."""

    def __init__(self, width=1024, layers=24, heads=16, vocab_size=49408, context_length=77):
        super().__init__()
        assert width == heads * 64, "the attention kernels are head_dim 64"
        self.width, self.heads, self.context_length = width, heads, context_length
        self.token_embedding = nn.Embedding(vocab_size, width)
        self.positional_embedding = nn.Parameter(torch.zeros(context_length, width))
        self.transformer = nn.Module()
        self.transformer.resblocks = nn.ModuleList(_Block(width, heads) for _ in range(layers))
        self.ln_final = nn.LayerNorm(width)
        self._packed = None

    @classmethod
    def from_state_dict(cls, sd):
        """builds the tower from open_clip's CLIP state dict (visual.* / text_projection / logit_scale are ignored)"""
        keys = [k for k in sd if k.startswith("transformer.resblocks.")]
        layers = 1 + max(int(k.split(".")[2]) for k in keys)
        vocab, width = sd["token_embedding.weight"].shape
        tower = cls(width=width, layers=layers, heads=width // 64, vocab_size=vocab,
                    context_length=sd["positional_embedding"].shape[0])
        own = tower.state_dict()
        tower.load_state_dict({k: sd[k] for k in own})
        return tower

    def load_state_dict(self, *a, **k):
        self._packed = None
        return super().load_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    def _pack(self):
        dev = self.positional_embedding.device

        def h(t):
            return t.detach().to(device=dev, dtype=torch.float16).contiguous()

        blocks = []
        for b in self.transformer.resblocks:
            blocks.append({"ln1": (h(b.ln_1.weight), h(b.ln_1.bias)), "qkv": (h(b.attn.in_proj_weight), h(b.attn.in_proj_bias)),
                           "out": (h(b.attn.out_proj.weight), h(b.attn.out_proj.bias)), "ln2": (h(b.ln_2.weight), h(b.ln_2.bias)),
                           "fc": (h(b.mlp.c_fc.weight), h(b.mlp.c_fc.bias)), "proj": (h(b.mlp.c_proj.weight), h(b.mlp.c_proj.bias))})
        self._packed = {"blocks": blocks, "lnf": (h(self.ln_final.weight), h(self.ln_final.bias))}
        return self._packed

    @torch.no_grad()
    def forward(self, tokens, skip_last=1):
        """tokens (B, 77) int64 -> (B, 77, width) fp32: ln_final of the residual stream after all but the last ``skip_last``
        blocks (embedder.py:52-71)."""
        pk = self._packed or self._pack()
        B, N = tokens.shape
        C, hd = self.width, self.heads
        x = (self.token_embedding.weight[tokens].float() + self.positional_embedding[:N].float()).reshape(B * N, C).half()
        blocks = pk["blocks"][:len(pk["blocks"]) - skip_last]
        for p in blocks:
            qkv = ops.linear(ops.layernorm(x, *p["ln1"]), *p["qkv"])                       # [B*N, 3C]: q | k | v, head h at 64h
            a = ops.attention_causal(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], B, hd, N)
            x = ops.linear(a, *p["out"], residual=x)
            f = ops.linear(ops.layernorm(x, *p["ln2"]), *p["fc"], flags=ops.FLAG_GELU_ERF)
            x = ops.linear(f, *p["proj"], residual=x)
        return ops.layernorm(x, *pk["lnf"]).float().reshape(B, N, C)


class FrozenOpenCLIPEmbedder(nn.Module):
    """
    Uses the OpenCLIP transformer encoder for text (reference contract), evaluated by ``OpenCLIPTextTower``
    """
    LAYERS = ["last", "penultimate"]

    def __init__(self, pretrained="laion2b_s32b_b79k", arch="ViT-H-14", device="cuda", max_length=77,
                 freeze=True, layer="penultimate", state_dict=None, tokenizer=None, weights_path=None, bpe_path=None):
        super().__init__()
        assert layer in self.LAYERS
        if weights_path is not None and state_dict is None:           # a local open_clip_pytorch_model.bin / .safetensors
            if str(weights_path).endswith(".safetensors"):
                from safetensors.torch import load_file
                state_dict = load_file(weights_path)
            else:
                state_dict = torch.load(weights_path, map_location="cpu")
            state_dict = state_dict.get("state_dict", state_dict)
        if bpe_path is not None and tokenizer is None:                # open_clip's bpe_simple_vocab_16e6.txt.gz
            from .clip_tokenizer import SimpleTokenizer
            tokenizer = SimpleTokenizer(bpe_path, context_length=max_length)
        if state_dict is None or tokenizer is None:
            try:
                import open_clip
            except ImportError as e:                                   # pragma: no cover
                raise ImportError("FrozenOpenCLIPEmbedder needs the `open_clip` package for the pretrained weights and the BPE "
                                  "tokenizer (or pass state_dict= and tokenizer=); precomputed (B,77,1024) text embeddings "
                                  "can be given to VideoToVideo_sr instead") from e
            if state_dict is None:
                clip, _, _ = open_clip.create_model_and_transforms(arch, device=torch.device("cpu"), pretrained=pretrained)
                del clip.visual
                state_dict = clip.state_dict()
            tokenizer = tokenizer or open_clip.tokenize
        self._tokenize = tokenizer
        self.model = OpenCLIPTextTower.from_state_dict(state_dict)     # same attribute name and key layout as the reference's
        self.device, self.max_length = device, max_length
        self.layer = layer
        self.layer_idx = 0 if layer == "last" else 1
        if freeze:
            self.freeze()

    def freeze(self):
        self.model = self.model.eval()
        for p in self.parameters():
            p.requires_grad = False

    def forward(self, text):
        tokens = self._tokenize(text)
        return self.encode_with_transformer(tokens.to(self.device))

    def encode_with_transformer(self, text):
        if self.model.positional_embedding.device != text.device:
            self.model.to(text.device)
        return self.model(text, skip_last=self.layer_idx)

    def encode(self, text):
        return self(text)
