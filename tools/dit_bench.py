"""STAR's CogVideoX-5B DiT layer at the BASELINE size (batch 2 = CFG pair, 226 text + 13*30*45 = 17 776 tokens, 3072 wide,
48 heads): parity against the (unpinned) fp32 restatement on the GPU and time per layer.  python tools/dit_bench.py"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from star_b200 import ops  # noqa: E402
from star_b200.cogvideox import DiTLayer  # noqa: E402
from star_b200.utils.synth import synth_state_dict  # noqa: E402


def main():
    from oracle.cogvideox_ref import DiTCfg, dit_layer_forward, layer_manifest, rope_tables
    cfg = DiTCfg()
    sd = synth_state_dict(layer_manifest(cfg), seed=3)
    cos, sin = rope_tables(cfg)
    S = cfg.text_length + cfg.frames * cfg.height * cfg.width
    g = torch.Generator().manual_seed(0)
    hidden = torch.randn(2, S, cfg.hidden, generator=g).cuda()
    emb = torch.randn(2, 512, generator=g).cuda()
    layer = DiTLayer(sd, cfg.hidden, cfg.heads, cfg.text_length, cfg.frames, cfg.height, cfg.width, cfg.ln_eps, cfg.qk_ln_eps,
                     cos, sin, device="cuda")
    out = layer.forward(hidden, emb)
    torch.cuda.synchronize()
    sdc = {k: v.cuda() for k, v in sd.items()}
    ref = dit_layer_forward(sdc, hidden, emb, cfg, cos.cuda(), sin.cuda())
    err = ((out.float() - ref).norm() / ref.norm()).item()
    print(f"DiT layer, batch 2 x {S} tokens x {cfg.hidden}: rel-L2 vs fp32 restatement (unpinned) {err:.3e}")
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        layer.forward(hidden, emb)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    d, ff = cfg.hidden, cfg.hidden * cfg.mlp_ratio
    flops = 2 * (2.0 * S * d * 3 * d + 2.0 * S * d * d + 4.0 * S * d * ff) + 2 * 4.0 * S * S * d
    print(f"{ms:.2f} ms per layer = {flops / ms / 1e9:.0f} TFLOP/s ({flops / 1e12:.2f} TFLOP); 42 layers x 50 steps = "
          f"{42 * 50 * ms / 1e3:.1f} s per 49-frame clip (both CFG branches in the batch)")
    ops.trace_begin()
    layer.forward(hidden, emb)
    agg = collections.defaultdict(float)
    for name, _s, t in ops.trace_end():
        agg[name] += t
    tot = sum(agg.values())
    print("  ops:", ", ".join(f"{k} {v:.2f} ms ({100 * v / tot:.0f}%)" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])))


if __name__ == "__main__":
    main()
