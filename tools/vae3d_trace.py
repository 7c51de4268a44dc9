#!/usr/bin/env python
"""Per-op device time of one CogVideoX 3-D VAE decode (13 latent -> 49 frames of 480x720) and one encode (49 frames -> latent), bf16:
CUDA events around every C-ABI call (star_b200.ops tracing).  python tools/vae3d_trace.py [--once: decode only, for ncu]"""
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from star_b200 import ops  # noqa: E402
from star_b200.cogvideox.vae3d import ContextParallelDecoder3D, ContextParallelEncoder3D  # noqa: E402
from star_b200.utils.synth import synth_tensor  # noqa: E402


def build(cls, seed, dev, dtype):
    with torch.device("meta"):
        m = cls()
    sd = {k: synth_tensor(k, v.shape, seed, dev) + (1.0 if ".conv_y.conv.bias" in k else 0.0) for k, v in m.state_dict().items()}
    m.load_state_dict(sd, assign=True)
    return m.to(dtype).eval()


def traced(name, fn):
    fn()
    torch.cuda.synchronize()
    ops.trace_begin()
    fn()
    agg, cnt = collections.defaultdict(float), collections.defaultdict(int)
    for op, _sig, ms in ops.trace_end():
        agg[op] += ms
        cnt[op] += 1
    tot = sum(agg.values())
    print(f"# {name}: {tot:.1f} ms in {sum(cnt.values())} op calls")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
        print(f"{k:24s} x{cnt[k]:4d} {v:9.2f} ms {100 * v / tot:6.2f} %")


def main():
    dev, dtype = torch.device("cuda", 0), torch.bfloat16
    dec = build(ContextParallelDecoder3D, 11, dev, dtype)
    z = torch.randn(1, 16, 13, 60, 90, generator=torch.Generator().manual_seed(5)).to(dev, dtype)
    if "--once" in sys.argv:
        dec.decode_latent(z)
        torch.cuda.synchronize()
        return
    traced("3-D VAE decode, 13 latent frames 60x90 -> 49 frames 480x720 (6 chunks)", lambda: dec.decode_latent(z))
    del dec
    torch.cuda.empty_cache()
    enc = build(ContextParallelEncoder3D, 13, dev, dtype)
    x = (torch.rand(1, 3, 49, 480, 720, generator=torch.Generator().manual_seed(6)) * 2 - 1).to(dev, dtype)
    traced("3-D VAE encode, 49 frames 480x720 -> moments (1, 32, 13, 60, 90)", lambda: enc(x))


if __name__ == "__main__":
    main()
