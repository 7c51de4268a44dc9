"""CogVideoX-5B DiT layer (STAR's patched block).  PARITY UNPINNED: the oracle (oracle/cogvideox_ref.py) restates
STAR's layer_forward on top of sat's default leaf modules, which are not in the reference tree."""
import pytest
import torch

from tests.util import rel_l2


def _setup(cfg_kw, seed=3):
    from oracle.cogvideox_ref import DiTCfg, layer_manifest, rope_tables
    from star_b200.utils.synth import synth_state_dict
    cfg = DiTCfg(**cfg_kw)
    sd = synth_state_dict(layer_manifest(cfg), seed=seed)
    cos, sin = rope_tables(cfg)
    g = torch.Generator().manual_seed(0)
    S = cfg.text_length + cfg.frames * cfg.height * cfg.width
    hidden = torch.randn(2, S, cfg.hidden, generator=g)
    emb = torch.randn(2, 512, generator=g)
    return cfg, sd, cos, sin, hidden, emb


def test_dit_layer_host_graph_on_emulated_kernels(monkeypatch):
    from oracle import kernel_ref as KR
    from oracle.cogvideox_ref import dit_layer_forward
    from star_b200 import ops
    from star_b200.cogvideox import DiTLayer
    for name in dir(KR):
        if not name.startswith("_") and callable(getattr(KR, name)) and hasattr(ops, name):
            monkeypatch.setattr(ops, name, getattr(KR, name))
    cfg, sd, cos, sin, hidden, emb = _setup(dict(hidden=256, heads=4, text_length=10, frames=3, height=6, width=5))
    ref = dit_layer_forward(sd, hidden, emb, cfg, cos, sin)
    layer = DiTLayer(sd, cfg.hidden, cfg.heads, cfg.text_length, cfg.frames, cfg.height, cfg.width, cfg.ln_eps,
                     cfg.qk_ln_eps, cos, sin, device="cpu")
    assert rel_l2(layer.forward(hidden, emb), ref) < 2e-3


def test_rope_tables_shape():
    from oracle.cogvideox_ref import DiTCfg, rope_tables
    cos, sin = rope_tables(DiTCfg())
    assert cos.shape == (13 * 30 * 45, 64) and torch.allclose(cos ** 2 + sin ** 2, torch.ones_like(cos), atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("cfg_kw", [dict(hidden=256, heads=4, text_length=10, frames=3, height=6, width=5),
                                    dict(hidden=3072, heads=48, text_length=226, frames=2, height=6, width=10)])
def test_dit_layer_gpu(cfg_kw):
    """full-width layer (3072 / 48 heads / 226 text tokens) at a short sequence, and a small layer"""
    from oracle.cogvideox_ref import dit_layer_forward
    from star_b200.cogvideox import DiTLayer
    cfg, sd, cos, sin, hidden, emb = _setup(cfg_kw)
    ref = dit_layer_forward(sd, hidden, emb, cfg, cos, sin)
    layer = DiTLayer(sd, cfg.hidden, cfg.heads, cfg.text_length, cfg.frames, cfg.height, cfg.width, cfg.ln_eps,
                     cfg.qk_ln_eps, cos, sin, device="cuda")
    out = layer.forward(hidden.cuda(), emb.cuda())
    torch.cuda.synchronize()
    err = rel_l2(out.cpu(), ref)
    print(f"DiT layer {cfg_kw}: rel-L2 vs fp32 oracle {err:.3e}")
    assert err < 3e-3
