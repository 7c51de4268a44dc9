#!/usr/bin/env python
"""Softmax-phase timeline of CTA (0,0,0) of the spatial-attention kernel (built with -DSTAR_ATTN_TRACE=1): per KV step and
query tile, when warp 4 / warp 12 lane 0 passes: 1 wait S, 2 S ready, 3 S in registers, 4 row max agreed, 5 exponentials
done, 6 PV(j-1) retired, 7 P stored.   python tools/attn_trace.py --lib tools/variants/libstar_atrace.so"""
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

F, heads, N, C = 32, 5, 122 * 216, 320
qkv = torch.randn(F * N, 3 * C, device="cuda").half()
lib = L.get_lib()
buf = (ctypes.c_longlong * (2 * 8192))()
cnt = (ctypes.c_int * 2)()
for rep in range(2):
    lib.star_debug_read_attn_trace(buf, cnt)
    O.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], F, heads, N, N, 1, 0.125)
    torch.cuda.synchronize()
lib.star_debug_read_attn_trace(buf, cnt)
names = {1: "wait_S", 2: "S_ready", 3: "S_in_regs", 4: "max_agreed", 5: "exp_done", 6: "pv_prev_done", 7: "P_stored"}
t0 = min(buf[t * 8192 + 1] for t in range(2))
print(f"# attn4 (B*h={F * heads}, N={N}): CTA (0,0,0), clocks relative to its first event; counts {list(cnt)}")
ev = [[(buf[t * 8192 + j], buf[t * 8192 + j + 1] - t0) for j in range(0, cnt[t], 2)] for t in range(2)]
for t in range(2):
    starts = [k for k, (e, _) in enumerate(ev[t]) if e == 1]
    print(f"## tile {t}: KV steps 100..102")
    prev = None
    for e, c in ev[t][starts[100]:starts[103]]:
        print(f"  {names[e]:13s} {c:9d}" + ("" if prev is None else f"  (+{c - prev})"))
        prev = c
    # phase durations averaged over steps 20..180
    acc = {}
    for k in range(20, 180):
        seg = ev[t][starts[k]:starts[k + 1] + 1]
        for (e0, c0), (e1, c1) in zip(seg[:-1], seg[1:]):
            acc.setdefault((e0, e1), []).append(c1 - c0)
    print(f"## tile {t}: mean clocks per phase, KV steps 20..179")
    for (e0, e1), v in acc.items():
        print(f"  {names[e0]:13s} -> {names[e1]:13s} {sum(v) / len(v):8.0f}")
    per = (ev[t][starts[180]][1] - ev[t][starts[20]][1]) / 160.0
    print(f"  -> {per:.0f} clk per KV step")
s0 = [c for e, c in ev[0] if e == 4]
s1 = [c for e, c in ev[1] if e == 4]
n = min(len(s0), len(s1))
off = [s1[k] - s0[k] for k in range(n)]
print("## offset of tile 1's exp-phase start relative to tile 0's, every 20th KV step:", off[::20])
