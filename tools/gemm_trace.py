#!/usr/bin/env python
"""Timeline of one persistent tap-GEMM CTA (built with -DSTAR_GEMM_TRACE=1): per tile, when the TMA producer, the MMA-issuing thread
and the epilogue leader reach their milestones.  python tools/gemm_trace.py --lib tools/variants/libstar_trace.so [K N flags]"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import star_b200.lib as L  # noqa: E402

i = sys.argv.index("--lib")
L.LIB_PATH = os.path.abspath(sys.argv[i + 1])
del sys.argv[i:i + 2]
from star_b200 import ops as O  # noqa: E402

K, N, flags = (int(v) for v in (sys.argv[1:4] + ["320", "960", "0"][len(sys.argv) - 1:]))
R = 32 * 122 * 216
a = (torch.randn(R, K, device="cuda")).half()
w = (torch.randn(N, K, device="cuda") * K ** -0.5).half()
bias = torch.randn(N, device="cuda").half()
lib = L.get_lib()
buf = (ctypes.c_longlong * (3 * 4096))()
cnt = (ctypes.c_int * 3)()
for rep in range(2):                                   # first launch warms up, second is read back
    lib.star_debug_read_trace(buf, cnt)
    O.linear(a, w, bias, None, None, 1, flags)
    torch.cuda.synchronize()
lib.star_debug_read_trace(buf, cnt)
names = {0: {1: "tile_start", 2: "loads_issued"}, 1: {1: "wait_acc", 2: "acc_free", 3: "mma_issued"},
         2: {1: "wait_acc_full", 2: "acc_full", 3: "pass_wait_buf", 4: "buf_free", 5: "bar1", 6: "staged", 7: "bar2"}}
t0 = min(buf[r * 4096 + 1] for r in range(3) if cnt[r] >= 2)
print(f"# GEMM {R}x{K}->{N} flags={flags}: CTA 0 timelines, clocks relative to its first event; counts {list(cnt)}")
for role, rname in enumerate(("producer", "mma", "epilogue")):
    ev = [(buf[role * 4096 + j], buf[role * 4096 + j + 1] - t0) for j in range(0, cnt[role], 2)]
    # print the steady state: tiles 20..26
    starts = [k for k, (e, _) in enumerate(ev) if e == 1]
    lo, hi = (starts[20], starts[27]) if len(starts) > 27 else (0, len(ev))
    print(f"## {rname}: tiles 20..26")
    prev = None
    for e, t in ev[lo:hi]:
        d = "" if prev is None else f"  (+{t - prev})"
        print(f"  {names[role][e]:14s} {t:9d}{d}")
        prev = t
    if len(starts) > 40:
        per = (ev[starts[40]][1] - ev[starts[20]][1]) / 20.0
        print(f"  -> {per:.0f} clk per tile (tiles 20..40)")
