"""VideoToVideo_sr.test() -- the REAL glue (pad_to_fit, upsample+pad, per-frame encode, chunking, sample_sr, 3-frame decode
windows, crop, layout) driven on CPU with injected stubs for the three heavy members (vae=, text_encoder=, generator=),
(a) against a straight restatement of the reference's test() (video_to_video_model.py:75-139) and
(b) on 2 gloo ranks: rank-local encode + latent all-gather, chunk-parallel denoise, rank-local decode + ONE all-gather of
    frames must reproduce the single-process result bit for bit (north-star end collective)."""
import os
import socket
from types import SimpleNamespace

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F
from einops import rearrange

from tests.util import FakeDenoiser


class StubVAE:
    """Deterministic stand-in with the three members the pipeline uses.  decode() depends on the WINDOW (mean over its
    frames) like the real decoder's temporal convs / 5-D norms do, so wrong 3-frame windowing changes the output."""
    config = SimpleNamespace(scaling_factor=0.18215)

    def encode(self, x):
        z = F.avg_pool2d(x, 8)
        z = torch.cat([z, z.mean(dim=1, keepdim=True)], dim=1)                       # (n, 4, h/8, w/8)
        return SimpleNamespace(latent_dist=SimpleNamespace(sample=lambda: z + 0.01 * torch.sin(z * 7.0)))

    def decode(self, z, num_frames):
        assert z.shape[0] == num_frames
        up = F.interpolate(z[:, :3], scale_factor=8, mode="nearest")
        return SimpleNamespace(sample=up + 0.1 * up.mean(dim=0, keepdim=True))


def make_pipe():
    from star_b200.video_to_video.video_to_video_model import VideoToVideo_sr
    emb = torch.randn(1, 77, 1024, generator=torch.Generator().manual_seed(3))
    pipe = VideoToVideo_sr(SimpleNamespace(model_path=None), device=torch.device("cpu"), text_encoder=lambda s: emb,
                           vae=StubVAE(), generator=FakeDenoiser())
    pipe.generator = pipe.generator.float()
    return pipe


def inputs(frames=5, h=24, w=40):
    g = torch.Generator().manual_seed(11)
    return {"video_data": torch.rand(frames, 3, h, w, generator=g) * 2 - 1, "y": "a cat", "target_res": (4 * h, 4 * w)}


def reference_test_restated(pipe, input, noise, sampler, steps, max_chunk_len):
    """video_to_video_model.py:75-139, line by line, on the same stubs"""
    from star_b200.video_to_video.video_to_video_model import make_chunks, pad_to_fit
    video_data = F.interpolate(input["video_data"], list(input["target_res"]), mode="bilinear")
    frames_num, _, h, w = video_data.shape
    padding = pad_to_fit(h, w)
    video_data = F.pad(video_data, padding, "constant", 1).unsqueeze(0)
    z = torch.cat([pipe.vae.encode(video_data[0, i:i + 1]).latent_dist.sample() for i in range(frames_num)])
    feat = rearrange(z, "(b f) c h w -> b c f h w", f=frames_num) * 0.18215
    y = pipe.text_encoder(input["y"])
    t = torch.LongTensor([999])
    noised = pipe.diffusion.diffuse(feat, t, noise=noise)
    chunk_inds = make_chunks(frames_num, 0, max_chunk_len) if frames_num > max_chunk_len else None
    gen = pipe.diffusion.sample_sr(noise=noised, model=pipe.generator, model_kwargs=[{"y": y}, {"y": pipe.negative_y}, {"hint": feat}],
                                   guide_scale=7.5, guide_rescale=0.2, solver="dpmpp_2m_sde", solver_mode="fast", steps=steps,
                                   t_max=999, t_min=0, discretization="trailing", chunk_inds=chunk_inds, noise_sampler=sampler)
    zf = rearrange(gen, "b c f h w -> (b f) c h w")
    vid = torch.cat([pipe.vae.decode(zf[i:i + 3] / 0.18215, num_frames=zf[i:i + 3].shape[0]).sample for i in range(0, zf.shape[0], 3)])
    w1, w2, h1, h2 = padding
    vid = vid[:, :, h1:h + h1, w1:w + w1]
    return rearrange(vid, "(b f) c h w -> b c f h w", b=1).float()


def _patched_run(pipe, inp, frames, steps, max_chunk_len):
    """test() with the two random inputs pinned (diffuse noise, SDE noise) through denoise_latents' own hooks"""
    lat = (1, 4, frames, 90, 160)                                    # 24x40 -> 96x160 -> padded to 720x1280 -> /8
    noise = torch.randn(lat, generator=torch.Generator().manual_seed(5))
    g = torch.Generator().manual_seed(6)
    sampler = lambda a, b: torch.randn(lat, generator=g)              # noqa: E731
    orig = pipe.denoise_latents

    def pinned(feat, y, neg, tnl, st, mode, gs, mcl):
        return orig(feat, y, neg, tnl, st, mode, gs, mcl, noise=noise, noise_sampler=sampler)
    pipe.denoise_latents = pinned
    out = pipe.test(inp, total_noise_levels=1000, steps=steps, solver_mode="fast", guide_scale=7.5, max_chunk_len=max_chunk_len)
    return out, noise


def test_pipeline_matches_restated_reference_single_and_chunked():
    for frames, mcl in ((5, 32), (7, 4)):                            # un-chunked; chunked (windows of 4, stride 2)
        pipe = make_pipe()
        inp = inputs(frames)
        out, noise = _patched_run(pipe, inp, frames, steps=15, max_chunk_len=mcl)
        g = torch.Generator().manual_seed(6)
        ref = reference_test_restated(make_pipe(), inp, noise, lambda a, b: torch.randn(noise.shape, generator=g), 15, mcl)
        assert out.shape == (1, 3, frames, 96, 160) and out.dtype == torch.float32 and out.device.type == "cpu"
        assert torch.allclose(out, ref, atol=1e-5, rtol=1e-5), float((out - ref).abs().max())


def _worker(rank, world, port, q, frames, mcl):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out, _ = _patched_run(make_pipe(), inputs(frames), frames, steps=15, max_chunk_len=mcl)
        q.put((rank, out.numpy()))            # by value: the worker may exit before the parent reads
    finally:
        dist.destroy_process_group()


def test_pipeline_two_ranks_equals_single_process():
    frames, mcl = 7, 4
    single, _ = _patched_run(make_pipe(), inputs(frames), frames, steps=15, max_chunk_len=mcl)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, frames, mcl)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, out in res:
        out = torch.from_numpy(out)
        assert torch.equal(out, single), f"rank {rank}: sharded pipeline differs from the single-process result"


def test_shard_helpers():
    from star_b200.video_to_video.video_to_video_model import _shard_bounds
    assert _shard_bounds(7, 2) == [(0, 4), (4, 7)]
    assert _shard_bounds(3, 8) == [(0, 1), (1, 2), (2, 3)] + [(3, 3)] * 5
    assert _shard_bounds(144, 8)[-1] == (126, 144)
