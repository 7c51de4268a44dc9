"""TEST INFRASTRUCTURE -- generate tests/golden/* by executing the UNMODIFIED reference
(/root/reference, build container only).  Run:  python -m oracle.make_golden

Outputs (all small; inputs/weights are regenerated from seeds by tests/util.py):
  unet_small.pt   reduced ControlledV2VUNet (dim_mult [1,2,1,4], 1 res block): fp32 output of the
                  real reference + its own fp16-autocast (CPU) output, two input shapes
  unet_full.pt    default ControlledV2VUNet (2.04 B parameters): fp32 output of the real reference
  sampler.pt      timestep / sigma tables, pad_to_fit / make_chunks tables, sample_sr results with a
                  cheap fake denoiser (un-chunked, chunked, 'fast' mode) and with the reduced UNet
  state_dict_manifest.json  {key: shape} of the default model
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_loader as R  # noqa: E402
from star_b200.utils.synth import synth_state_dict  # noqa: E402
from tests.util import SMALL_KW, FakeDenoiser, make_inputs  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def build_ref(kw, seed):
    ns = R.load_reference()
    U = ns.unet
    with torch.device("meta"):
        net = U.ControlledV2VUNet.__new__(U.ControlledV2VUNet)
        U.Vid2VidSDUNet.__init__(net, **kw)
        net.VideoControlNet = U.VideoControlNet(**kw)
    manifest = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    sd = synth_state_dict(manifest, seed=seed)
    net.load_state_dict(sd, assign=True)
    return net.eval(), manifest


def noise_stream(seed, shape):
    """The N(0,I) stream both samplers consume (reference: via oracle.ref_loader.InjectedBrownian,
    whose W(t0,t1) sign convention makes the solver see -randn)."""
    g = torch.Generator().manual_seed(seed)
    return lambda a, b: -torch.randn(shape, generator=g)


@torch.no_grad()
def main():
    os.makedirs(OUT, exist_ok=True)
    ns = R.load_reference()
    logging_off()

    # ---- UNet forward, reduced model --------------------------------------------------------
    net, _ = build_ref(SMALL_KW, seed=1)
    cases = []
    for (seed, B, F, H, W, t) in [(0, 1, 4, 18, 16, [899]), (3, 2, 5, 10, 8, [500, 34]), (5, 1, 8, 26, 24, [281])]:
        x, hint, y = make_inputs(seed, B, F, H, W)
        tt = torch.tensor(t)
        o32 = net(x, tt, y, hint=hint)
        cases.append(dict(seed=seed, B=B, F=F, H=H, W=W, t=t, out_fp32=o32.clone()))
        print("small", seed, B, F, H, W, float(o32.std()))
    neth = net.half()
    for c in cases:
        x, hint, y = make_inputs(c["seed"], c["B"], c["F"], c["H"], c["W"])
        with torch.autocast("cpu", dtype=torch.float16):
            o16 = neth(x, torch.tensor(c["t"]), y, hint=hint)
        c["out_ref_fp16"] = o16.clone()
        c["ref_fp16_rel_err"] = float((o16.float() - c["out_fp32"]).norm() / c["out_fp32"].norm())
        print("  ref fp16-autocast rel err", c["ref_fp16_rel_err"])
    torch.save(dict(kw=SMALL_KW, weight_seed=1, cases=cases), os.path.join(OUT, "unet_small.pt"))

    # ---- sampler with the reduced UNet (chunked, 2 steps) --------------------------------------
    net = net.float()
    sig = ns.schedules_sdedit.noise_schedule(schedule="logsnr_cosine_interp", n=1000, zero_terminal_snr=True,
                                             scale_min=2.0, scale_max=4.0)
    diff = ns.diffusion_sdedit.GaussianDiffusion(sigmas=sig)
    sampler = {}
    x, hint, y = make_inputs(11, 1, 12, 10, 8)
    _, _, ny = make_inputs(12, 1, 12, 10, 8)
    R.InjectedBrownian.seed = 99
    out = diff.sample_sr(noise=x.clone(), model=net, model_kwargs=[{"y": y}, {"y": ny}, {"hint": hint}],
                         guide_scale=7.5, guide_rescale=0.2, solver="dpmpp_2m_sde", solver_mode="normal", steps=2,
                         t_max=899, t_min=0, discretization="trailing", chunk_inds=[(0, 8), (4, 12)])
    sampler["unet_chunked"] = dict(input_seed=11, neg_seed=12, noise_seed=99, F=12, H=10, W=8, steps=2,
                                   chunk_inds=[(0, 8), (4, 12)], out=out.clone())
    print("sampler+unet chunked", float(out.std()))
    del net, neth

    # ---- sampler tables and fake-denoiser runs -------------------------------------------------
    sampler["sigmas_probe"] = {i: float(sig[i]) for i in (0, 1, 100, 250, 500, 750, 899, 998, 999)}
    sampler["sigmas_full"] = sig.clone()
    runs = []
    for (F, chunks, mode, steps) in [(8, None, "normal", 2), (12, [(0, 8), (4, 12)], "normal", 3),
                                     (8, None, "fast", 15), (20, [(0, 8), (4, 12), (8, 20)], "normal", 4),
                                     (8, None, "normal", 50)]:
        x, hint, y = make_inputs(21, 1, F, 10, 8)
        _, _, ny = make_inputs(22, 1, F, 10, 8)
        R.InjectedBrownian.seed = 77
        m = FakeDenoiser()
        o = diff.sample_sr(noise=x.clone(), model=m, model_kwargs=[{"y": y}, {"y": ny}, {"hint": hint}],
                           guide_scale=7.5, guide_rescale=0.2, solver="dpmpp_2m_sde", solver_mode=mode, steps=steps,
                           t_max=899, t_min=0, discretization="trailing", chunk_inds=chunks)
        runs.append(dict(F=F, chunks=chunks, mode=mode, steps=steps, timesteps=m.calls[::2 * (len(chunks) if chunks else 1)],
                         all_calls=m.calls, out=o.clone()))
        print("fake", F, chunks, mode, steps, m.calls[:6])
    sampler["fake_runs"] = runs
    sampler["pad_to_fit"] = {f"{h}x{w}": list(ns.pad_to_fit(h, w)) for (h, w) in
                             [(256, 256), (540, 960), (720, 1280), (768, 1360), (960, 1704), (1080, 1920), (1440, 2560),
                              (100, 3000), (721, 1281)]}
    sampler["make_chunks"] = {str(F): ns.make_chunks(F, 0, 32) for F in (33, 40, 41, 48, 64, 72, 78, 100, 144)}
    sampler["make_chunks_16"] = {str(F): ns.make_chunks(F, 0, 16) for F in (21, 24, 40)}
    x0 = torch.randn(1, 4, 3, 5, 4, generator=torch.Generator().manual_seed(5))
    nz = torch.randn(1, 4, 3, 5, 4, generator=torch.Generator().manual_seed(6))
    sampler["diffuse"] = dict(x0=x0, noise=nz, t=899, out=diff.diffuse(x0, torch.tensor([899]), noise=nz))
    torch.save(sampler, os.path.join(OUT, "sampler.pt"))

    # ---- UNet forward, full model ---------------------------------------------------------------
    net, manifest = build_ref({}, seed=2)
    json.dump({k: list(v) for k, v in manifest.items()}, open(os.path.join(OUT, "state_dict_manifest.json"), "w"))
    cases = []
    for (seed, B, F, H, W, t) in [(7, 1, 4, 18, 16, [793])]:
        x, hint, y = make_inputs(seed, B, F, H, W)
        o32 = net(x, torch.tensor(t), y, hint=hint)
        cases.append(dict(seed=seed, B=B, F=F, H=H, W=W, t=t, out_fp32=o32.clone()))
        print("full", float(o32.std()))
    torch.save(dict(kw={}, weight_seed=2, cases=cases), os.path.join(OUT, "unet_full.pt"))


def logging_off():
    import logging
    logging.getLogger("video_to_video").setLevel(logging.ERROR)
    for h in logging.getLogger("video_to_video").handlers:
        h.setLevel(logging.ERROR)


if __name__ == "__main__":
    main()
