"""Temporal VAE at the BASELINE size: decode of one 3-frame window (latent 122x216 -> 976x1728) and encode of one
frame, synthetic weights; prints per-op time shares (CUDA events).  python tools/vae_bench.py [h w]"""
import collections
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from star_b200 import ops  # noqa: E402
from star_b200.utils.synth import synth_tensor  # noqa: E402
from star_b200.video_to_video.modules.temporal_vae import AutoencoderKLTemporalDecoder  # noqa: E402


def build():
    with torch.device("meta"):
        vae = AutoencoderKLTemporalDecoder()
    sd = {k: synth_tensor(k, v.shape, 0, "cuda") for k, v in vae.state_dict().items()}
    vae.load_state_dict(sd, assign=True)
    return vae.eval().requires_grad_(False)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def shares(fn):
    ops.trace_begin()
    fn()
    agg = collections.defaultdict(float)
    for name, _sig, ms in ops.trace_end():
        agg[name] += ms
    tot = sum(agg.values())
    return ", ".join(f"{k} {v:.1f} ms ({100 * v / tot:.0f}%)" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]))


def main():
    h, w = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (122, 216)
    vae = build()
    z = torch.randn(3, 4, h, w, device="cuda")
    x = torch.rand(1, 3, 8 * h, 8 * w, device="cuda") * 2 - 1
    t0 = time.time()
    out = vae.decode(z, num_frames=3).sample
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    print(f"first decode {time.time() - t0:.2f} s, out {tuple(out.shape)} |x|max {out.abs().max().item():.3f}", flush=True)
    ms = timed(lambda: vae.decode(z, num_frames=3))
    print(f"decode 3 frames @ latent {h}x{w}: {ms:.1f} ms = {ms / 3:.1f} ms/frame "
          f"({20.8e12 * 3 * (h * w) / (122 * 216) / ms / 1e9:.0f} TFLOP/s on the analytic 20.8 TFLOP/frame)")
    print("  ops:", shares(lambda: vae.decode(z, num_frames=3)))
    m = vae.encode(x).latent_dist.parameters
    assert torch.isfinite(m).all()
    ms = timed(lambda: vae.encode(x))
    print(f"encode 1 frame @ {8 * h}x{8 * w}: {ms:.1f} ms ({8.5e12 * (h * w) / (122 * 216) / ms / 1e9:.0f} TFLOP/s on 8.5 TFLOP/frame)")
    print("  ops:", shares(lambda: vae.encode(x)))
    print(f"peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")


if __name__ == "__main__":
    main()
