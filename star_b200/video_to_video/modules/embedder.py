"""Text-encoder boundary (reference: video_to_video/modules/embedder.py:12-74).

The OpenCLIP ViT-H-14 text tower is outside the accelerated hot path: the
north-star fixes the (B, 77, 1024) text embedding as an input of the denoiser.
``FrozenOpenCLIPEmbedder`` therefore stays a thin wrapper over the un-vendored
``open_clip`` package (open-clip-torch==2.20.0 in the reference's
requirements.txt:9) with the reference's constructor/``forward``/``encode``
contract; it is imported lazily so that the rest of the package works without
open_clip (the bench and the parity tests feed text embeddings directly)."""
import torch
import torch.nn as nn

__all__ = ["FrozenOpenCLIPEmbedder"]


class FrozenOpenCLIPEmbedder(nn.Module):
    LAYERS = ["last", "penultimate"]

    def __init__(self, pretrained="laion2b_s32b_b79k", arch="ViT-H-14", device="cuda", max_length=77,
                 freeze=True, layer="penultimate"):
        super().__init__()
        try:
            import open_clip
        except ImportError as e:                                   # pragma: no cover
            raise ImportError("FrozenOpenCLIPEmbedder needs the `open_clip` package; pass precomputed "
                              "(B,77,1024) text embeddings to VideoToVideo_sr instead") from e
        assert layer in self.LAYERS
        model, _, _ = open_clip.create_model_and_transforms(arch, device=torch.device("cpu"), pretrained=pretrained)
        del model.visual
        self._tokenize = open_clip.tokenize
        self.model, self.device, self.max_length = model, device, max_length
        self.layer_idx = 0 if layer == "last" else 1
        if freeze:
            self.model = self.model.eval()
            for p in self.parameters():
                p.requires_grad = False

    def forward(self, text):
        tokens = self._tokenize(text)
        return self.encode_with_transformer(tokens.to(self.device))

    def encode_with_transformer(self, text):
        m = self.model
        x = m.token_embedding(text) + m.positional_embedding
        x = x.permute(1, 0, 2)
        blocks = m.transformer.resblocks
        for i, blk in enumerate(blocks):
            if i == len(blocks) - self.layer_idx:
                break
            x = blk(x, attn_mask=m.attn_mask)
        return m.ln_final(x.permute(1, 0, 2))

    def encode(self, text):
        return self(text)
