#!/usr/bin/env python
"""Micro-benchmarks of the CogVideoX 3-D VAE kernels at the decoder's real shapes (one latent chunk: 9 frames at 480x720 / 240x360,
5 at 120x180, 3 at 60x90).  CUDA events, 2 warm-up + 3 timed launches.  python tools/vae3d_kbench.py   (ncu: -k regex:tapgemm)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from star_b200 import ops as O  # noqa: E402
from tools.kbench import PEAK, report, rnd, timeit  # noqa: E402

SHAPES = [(9, 480, 720, 128, 128), (9, 480, 720, 256, 128), (9, 240, 360, 256, 256), (5, 120, 180, 512, 256), (3, 60, 90, 512, 512)]


def main():
    print(f"# peaks used: {PEAK}")
    for (T, H, W, Cin, Cout) in SHAPES:
        xp = rnd((T + 2) * H * W, Cin)
        w = rnd(Cout, 3, 3, 3, Cin, scale=(27 * Cin) ** -0.5)
        b = rnd(Cout, scale=0.1)
        res = rnd(T * H * W, Cout)
        report(f"conv3d_causal 3x3x3 {Cin}->{Cout} @ {T}x{H}x{W} (+res)", timeit(lambda: O.conv3d_causal(xp, w, T, H, W, b, residual=res), 3),
               flops=2.0 * T * H * W * 27 * Cin * Cout, bytes_=2.0 * ((T + 2) * H * W * Cin + 2 * T * H * W * Cout))
        del xp, w, res
    for (T, H, W, C, Tl) in [(9, 480, 720, 128, 3), (9, 240, 360, 256, 3), (8, 480, 720, 128, 2)]:
        x = rnd(T * H * W, C)
        g, be = rnd(C), rnd(C)
        mod = rnd(Tl * 60 * 90, 2 * C)
        out = torch.empty_like(x)
        report(f"groupnorm_mod (SpatialNorm3D + SiLU) C={C} @ {T}x{H}x{W}",
               timeit(lambda: O.groupnorm_mod(x, g, be, mod[:, :C], mod[:, C:], T, H, W, Tl, 60, 90, 1e-6, True, out=out), 3),
               bytes_=3.0 * T * H * W * C * 2)
        del x, out
    x = rnd(49 * 480 * 720, 128)
    report("time_avgpool2 49 -> 25 frames @ 480x720, C=128", timeit(lambda: O.time_avgpool2(x, 49, 480 * 720), 3),
           bytes_=(49 + 25) * 480 * 720 * 128 * 2.0)


if __name__ == "__main__":
    main()
