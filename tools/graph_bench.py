#!/usr/bin/env python
"""Eager vs CUDA-graphed CFG-pair forward of the FULL ControlledV2VUNet at small latents (BASELINE config 1 sizes): where the
~3 300 launches of a solver step take longer to issue than to run.  python tools/graph_bench.py"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import build_model  # noqa: E402
from star_b200 import ops  # noqa: E402
from star_b200.video_to_video.cuda_graph import GraphedCFGPair  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    net, _, _ = build_model(False, dev)
    g = GraphedCFGPair(net)
    print("# full ControlledV2VUNet (2.04 B params), one CFG-pair forward = the model work of one solver step")
    for (F, H, W) in ((8, 18, 16), (8, 34, 32), (8, 90, 160), (32, 122, 216)):
        gen = torch.Generator().manual_seed(0)
        x = torch.randn(1, 4, F, H, W, generator=gen).to(dev)
        hint = torch.randn(1, 4, F, H, W, generator=gen).to(dev)
        y, ny = torch.randn(1, 77, 1024, generator=gen).to(dev), torch.randn(1, 77, 1024, generator=gen).to(dev)
        t = torch.tensor([500], device=dev)
        res = {}
        for name, fn in (("eager", net.forward_cfg_pair), ("graph", g.forward_cfg_pair)):
            if name == "graph" and F * H * W > 8 * 90 * 160:
                continue                                            # GPU-bound shape: no graph (its pool would pin ~40 GB)
            for _ in range(2):
                out = fn(x, t, (y, ny), hint=hint)
            torch.cuda.synchronize()
            n0, w0 = ops.launch_count(), time.perf_counter()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                out = fn(x, t, (y, ny), hint=hint)
            b.record()
            torch.cuda.synchronize()
            res[name] = (a.elapsed_time(b) / 5, (time.perf_counter() - w0) / 5 * 1e3, (ops.launch_count() - n0) // 5, out)
        e = res["eager"]
        line = f"F={F:2d} latent {H:3d}x{W:3d}: eager {e[0]:8.2f} ms (host wall {e[1]:8.2f} ms, {e[2]} launches)"
        if "graph" in res:
            gq = res["graph"]
            same = torch.equal(gq[3][0], e[3][0]) and torch.equal(gq[3][1], e[3][1])
            line += f"   graph {gq[0]:8.2f} ms   speed-up {e[0] / gq[0]:.2f}x   bit-identical {same}"
        print(line, flush=True)


if __name__ == "__main__":
    main()
