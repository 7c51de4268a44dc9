"""Sampler / scheduler / chunk index parity (host logic, CPU): star_b200's GaussianDiffusion against
golden vectors recorded from the real reference, plus the known-answer values of SURVEY App. B."""
import os

import pytest
import torch

from tests.util import FakeDenoiser, make_inputs

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gold():
    return torch.load(os.path.join(GOLD, "sampler.pt"))


@pytest.fixture(scope="module")
def diffusion():
    from star_b200.video_to_video.diffusion.diffusion_sdedit import GaussianDiffusion
    from star_b200.video_to_video.diffusion.schedules_sdedit import noise_schedule
    sig = noise_schedule(schedule="logsnr_cosine_interp", n=1000, zero_terminal_snr=True, scale_min=2.0, scale_max=4.0)
    return GaussianDiffusion(sigmas=sig)


def test_noise_schedule_bit_exact(gold, diffusion):
    assert torch.equal(diffusion.sigmas, gold["sigmas_full"])
    kat = {0: 0.001106, 1: 0.004252, 100: 0.322620, 250: 0.702541, 500: 0.943010, 750: 0.992548, 899: 0.999094,
           998: 1.0, 999: 1.0}
    for i, v in kat.items():
        assert abs(float(diffusion.sigmas[i]) - v) < 1e-6
    assert abs(float(diffusion.alphas[899]) - 0.0425660) < 1e-6


def test_diffuse(gold, diffusion):
    d = gold["diffuse"]
    out = diffusion.diffuse(d["x0"], torch.tensor([d["t"]]), noise=d["noise"])
    assert torch.equal(out, d["out"])


def _neg_randn_stream(seed, shape):
    g = torch.Generator().manual_seed(seed)
    return lambda a, b: -torch.randn(shape, generator=g)


def test_sample_sr_against_reference_runs(gold, diffusion):
    """timestep tables (50-step 'normal', 14-step 'fast', 2-step) and solver outputs, un-chunked and
    chunked, identical noise stream."""
    for run in gold["fake_runs"]:
        x, hint, y = make_inputs(21, 1, run["F"], 10, 8)
        _, _, ny = make_inputs(22, 1, run["F"], 10, 8)
        m = FakeDenoiser()
        out = diffusion.sample_sr(noise=x.clone(), model=m, model_kwargs=[{"y": y}, {"y": ny}, {"hint": hint}],
                                  guide_scale=7.5, guide_rescale=0.2, solver="dpmpp_2m_sde",
                                  solver_mode=run["mode"], steps=run["steps"], t_max=899, t_min=0,
                                  discretization="trailing", chunk_inds=run["chunks"],
                                  noise_sampler=_neg_randn_stream(77, x.shape))
        assert m.calls == run["all_calls"]
        assert torch.allclose(out, run["out"], rtol=1e-4, atol=1e-4), (run["F"], run["mode"], (out - run["out"]).abs().max())
    normal50 = [r for r in gold["fake_runs"] if r["steps"] == 50][0]["timesteps"]
    assert normal50[:5] == [899, 881, 864, 846, 828] and normal50[-3:] == [70, 52, 34] and len(normal50) == 50
    fast = [r for r in gold["fake_runs"] if r["mode"] == "fast"][0]["timesteps"]
    assert fast == [899, 799, 699, 599, 500, 454, 409, 363, 318, 272, 227, 181, 136, 90]


def test_pad_and_chunk_tables(gold):
    from star_b200.video_to_video.video_to_video_model import make_chunks, pad_to_fit
    for k, v in gold["pad_to_fit"].items():
        h, w = map(int, k.split("x"))
        assert list(pad_to_fit(h, w)) == v, k
    assert list(pad_to_fit(960, 1704)) == [0, 24, 0, 16]
    for k, v in gold["make_chunks"].items():
        assert [tuple(c) for c in make_chunks(int(k), 0, 32)] == [tuple(c) for c in v], k
    for k, v in gold["make_chunks_16"].items():
        assert [tuple(c) for c in make_chunks(int(k), 0, 16)] == [tuple(c) for c in v], k
    assert make_chunks(72, 0, 32) == [(0, 32), (16, 48), (32, 72)]


def test_single_window_is_unchunked(diffusion):
    """F in 33..40 gives one window; the reference raises IndexError there -- the
    product treats it as the un-chunked case."""
    x, hint, y = make_inputs(1, 1, 6, 10, 8)
    m = FakeDenoiser()
    out = diffusion.sample_sr(noise=x, model=m, model_kwargs=[{"y": y}, {"y": y}, {"hint": hint}], guide_scale=7.5,
                              guide_rescale=0.2, solver="dpmpp_2m_sde", solver_mode="normal", steps=2, t_max=899,
                              t_min=0, discretization="trailing", chunk_inds=[(0, 6)])
    assert out.shape == x.shape
