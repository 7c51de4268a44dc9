"""OpenCLIP text tower on the kernels.  PARITY UNPINNED: oracle/open_clip_text_ref.py restates
open_clip 2.20.0's block wiring on torch's own MultiheadAttention / LayerNorm / GELU; open_clip itself is absent."""
import pytest
import torch

from tests.util import assert_close, rel_l2


def _patch(monkeypatch):
    from oracle import kernel_ref as KR
    from star_b200 import ops
    for name in dir(KR):
        if not name.startswith("_") and callable(getattr(KR, name)) and hasattr(ops, name):
            monkeypatch.setattr(ops, name, getattr(KR, name))


def _setup(width, layers, seed=2, device="cpu"):
    from oracle.open_clip_text_ref import text_manifest
    from star_b200.utils.synth import synth_state_dict
    from star_b200.video_to_video.modules.embedder import FrozenOpenCLIPEmbedder
    sd = synth_state_dict(text_manifest(width, layers, vocab=512), seed=seed)
    sd["token_embedding.weight"] = sd["token_embedding.weight"] * width ** 0.5 * 0.02      # CLIP-like embedding scale
    sd["positional_embedding"] = sd["positional_embedding"] * width ** 0.5 * 0.01
    sd["text_projection"], sd["logit_scale"] = torch.zeros(width, width), torch.zeros(())   # present in CLIP, unused here

    def tokenizer(text):                                           # stand-in for open_clip.tokenize: (B, 77) int64
        g = torch.Generator().manual_seed(len(text[0]) if isinstance(text, (list, tuple)) else len(text))
        n = 1 if isinstance(text, str) else len(text)
        return torch.randint(0, 512, (n, 77), generator=g)
    emb = FrozenOpenCLIPEmbedder(device=device, state_dict=sd, tokenizer=tokenizer)
    return sd, emb.to(device), tokenizer


def test_text_tower_host_graph_on_emulated_kernels(monkeypatch):
    from oracle.open_clip_text_ref import encode_with_transformer
    _patch(monkeypatch)
    sd, emb, tok = _setup(128, 4)
    assert set(emb.model.state_dict()) == {k for k in sd if k not in ("text_projection", "logit_scale")}
    text = ["a cat", "two dogs"]
    got = emb.encode(text)
    want = encode_with_transformer(sd, tok(text), heads=2, skip_last=1)
    assert got.shape == (2, 77, 128) and got.dtype == torch.float32
    assert rel_l2(got, want) < 3e-3
    emb.layer_idx = 0                                              # layer='last'
    assert rel_l2(emb.encode(text), encode_with_transformer(sd, tok(text), heads=2, skip_last=0)) < 3e-3
    assert rel_l2(emb.encode(text), want) > 1e-2


@pytest.mark.gpu
@pytest.mark.parametrize("batch,heads,N", [(2, 16, 77), (3, 4, 128), (1, 8, 200), (2, 2, 33)])
def test_attention_causal(batch, heads, N):
    from oracle import kernel_ref as KR
    from star_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(N)
    qkv = torch.randn(batch * N, 3 * heads * 64, device="cuda", generator=g).half()
    C = heads * 64
    got = ops.attention_causal(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], batch, heads, N)
    assert_close(got, KR.attention_causal(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], batch, heads, N), what="attention_causal")


@pytest.mark.gpu
def test_linear_gelu_erf():
    from oracle import kernel_ref as KR
    from star_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(4)
    a = torch.randn(154, 1024, device="cuda", generator=g).half()
    w = (torch.randn(4096, 1024, device="cuda", generator=g) / 32).half()
    b = (torch.randn(4096, device="cuda", generator=g) * 0.1).half()
    assert_close(ops.linear(a, w, b, flags=ops.FLAG_GELU_ERF), KR.linear(a, w, b, flags=KR.FLAG_GELU_ERF), what="linear+gelu_erf")


@pytest.mark.gpu
def test_text_tower_gpu_full_width():
    """ViT-H-14 text tower at full size (24 blocks x 1024, 16 heads, 23 evaluated), both prompts of the pipeline"""
    from oracle.open_clip_text_ref import encode_with_transformer
    sd, emb, tok = _setup(1024, 24, device="cuda")
    text = ["positive prompt", "negative"]
    got = emb.encode(text)
    want = encode_with_transformer({k: v.cuda() for k, v in sd.items()}, tok(text).cuda(), heads=16, skip_last=1)
    err = rel_l2(got, want)
    print(f"[text tower] rel-L2 vs fp32 restatement (unpinned) {err:.2e}")
    assert got.shape == (2, 77, 1024) and torch.isfinite(got).all() and err < 4e-3


def test_bpe_tokenizer_mechanics(tmp_path):
    """CLIP's byte-level BPE on a synthetic merge list: merge order by rank, end-of-word marker, lower-casing / whitespace cleaning,
    <start_of_text> / <end_of_text>, zero padding and truncation to the context length"""
    import gzip
    from star_b200.video_to_video.modules.clip_tokenizer import SimpleTokenizer, bytes_to_unicode
    merges = ["#version: synthetic", "c a", "ca t</w>", "d o", "do g</w>", "t h", "th e</w>"]
    path = tmp_path / "bpe.txt.gz"
    with gzip.open(path, "wb") as f:
        f.write("\n".join(merges).encode())
    tok = SimpleTokenizer(str(path), context_length=8)
    assert len(bytes_to_unicode()) == 256 and len(set(bytes_to_unicode().values())) == 256
    base = 512                                                     # 256 byte symbols + 256 end-of-word variants, then the merges
    assert tok.encoder["ca"] == base and tok.encoder["cat</w>"] == base + 1 and tok.encoder["the</w>"] == base + 5
    assert tok.sot == base + 6 and tok.eot == base + 7
    ids = tok("  The   CAT &amp; dog ")                            # cleaned: "the cat & dog"
    amp = tok.encoder["&</w>"]
    assert ids.shape == (1, 8) and ids.dtype == torch.long
    assert ids[0].tolist() == [tok.sot, base + 5, base + 1, amp, base + 3, tok.eot, 0, 0]
    assert tok.encode("cats") == [base, tok.encoder["t"], tok.encoder["s</w>"]]      # 'ca' merges, 't s</w>' has no rule
    long = tok(["cat " * 20, "dog"])
    assert long.shape == (2, 8) and long[0, 0] == tok.sot and long[0, -1] == tok.eot and (long[0, 1:-1] == base + 1).all()
    assert long[1].tolist() == [tok.sot, base + 3, tok.eot, 0, 0, 0, 0, 0]


def test_embedder_without_open_clip(tmp_path, monkeypatch):
    """FrozenOpenCLIPEmbedder from a local weight file + merge list: nothing of open_clip is imported"""
    import builtins
    import gzip
    from oracle.open_clip_text_ref import encode_with_transformer, text_manifest
    from star_b200.utils.synth import synth_state_dict
    from star_b200.video_to_video.modules.embedder import FrozenOpenCLIPEmbedder
    _patch(monkeypatch)
    real_import = builtins.__import__
    monkeypatch.setattr(builtins, "__import__", lambda name, *a, **k: (_ for _ in ()).throw(ImportError(name))
                        if name == "open_clip" else real_import(name, *a, **k))
    sd = synth_state_dict(text_manifest(128, 3, vocab=520), seed=8)
    torch.save({"state_dict": sd}, tmp_path / "w.bin")
    with gzip.open(tmp_path / "bpe.txt.gz", "wb") as f:
        f.write(b"#v\nc a\nca t</w>\nd o\ndo g</w>\nt h\nth e</w>")
    emb = FrozenOpenCLIPEmbedder(device="cpu", weights_path=str(tmp_path / "w.bin"), bpe_path=str(tmp_path / "bpe.txt.gz"))
    got = emb.encode(["the cat", "dog"])
    want = encode_with_transformer(sd, emb._tokenize(["the cat", "dog"]), heads=2, skip_last=1)
    assert got.shape == (2, 77, 128) and rel_l2(got, want) < 3e-3
