"""STAR's CogVideoX-5B DiT layer at the BASELINE size (batch 2 = CFG pair, 226 text + 13*30*45 = 17 776 tokens, 3072 wide,
48 heads): parity against the (unpinned) fp32 restatement on the GPU and time per layer.  python tools/dit_bench.py"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from star_b200 import ops  # noqa: E402
from star_b200.cogvideox import DiTLayer  # noqa: E402
from star_b200.utils.synth import synth_state_dict  # noqa: E402


def main():
    from oracle.cogvideox_ref import DiTCfg, dit_layer_forward, layer_manifest, rope_tables
    cfg = DiTCfg()
    sd = synth_state_dict(layer_manifest(cfg), seed=3)
    cos, sin = rope_tables(cfg)
    S = cfg.text_length + cfg.frames * cfg.height * cfg.width
    g = torch.Generator().manual_seed(0)
    hidden = torch.randn(2, S, cfg.hidden, generator=g).cuda()
    emb = torch.randn(2, 512, generator=g).cuda()
    layer = DiTLayer(sd, cfg.hidden, cfg.heads, cfg.text_length, cfg.frames, cfg.height, cfg.width, cfg.ln_eps, cfg.qk_ln_eps,
                     cos, sin, device="cuda")
    out = layer.forward(hidden, emb)
    torch.cuda.synchronize()
    sdc = {k: v.cuda() for k, v in sd.items()}
    ref = dit_layer_forward(sdc, hidden, emb, cfg, cos.cuda(), sin.cuda())
    err = ((out.float() - ref).norm() / ref.norm()).item()
    print(f"DiT layer, batch 2 x {S} tokens x {cfg.hidden}: rel-L2 vs fp32 restatement (unpinned) {err:.3e}")
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        layer.forward(hidden, emb)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    d, ff = cfg.hidden, cfg.hidden * cfg.mlp_ratio
    flops = 2 * (2.0 * S * d * 3 * d + 2.0 * S * d * d + 4.0 * S * d * ff) + 2 * 4.0 * S * S * d
    print(f"{ms:.2f} ms per layer = {flops / ms / 1e9:.0f} TFLOP/s ({flops / 1e12:.2f} TFLOP); 42 layers x 50 steps = "
          f"{42 * 50 * ms / 1e3:.1f} s per 49-frame clip (both CFG branches in the batch)")
    ops.trace_begin()
    layer.forward(hidden, emb)
    agg = collections.defaultdict(float)
    for name, _s, t in ops.trace_end():
        agg[name] += t
    tot = sum(agg.values())
    print("  ops:", ", ".join(f"{k} {v:.2f} ms ({100 * v / tot:.0f}%)" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])))


def vae3d_encode_leg(dev, dtype, F, H, W):
    """the encode in front of config 4 (diffusion_video.py:279-282): the pre-upsampled LQ clip (49 x 480 x 720) -> latent"""
    from star_b200.cogvideox.vae3d import ContextParallelEncoder3D
    from star_b200.utils.synth import synth_tensor
    try:
        with torch.device("meta"):
            enc = ContextParallelEncoder3D()
        sd = {k: synth_tensor(k, v.shape, 13, dev) for k, v in enc.state_dict().items()}
        enc.load_state_dict(sd, assign=True)
        enc = enc.to(dtype).eval()
        x = (torch.rand(1, 3, F, H, W, generator=torch.Generator().manual_seed(6)) * 2 - 1).to(dev, dtype)
        m = enc(x)
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        m = enc(x)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        leg = {"frames": F, "encode_ms_per_clip": ms, "encode_ms_per_frame": ms / F, "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
               "finite": bool(torch.isfinite(m.float()).all())}
        try:
            from oracle.cogvideox_vae import build_reference_encoder, reference_encode_moments, vae_reference_available
            if vae_reference_available():
                renc = build_reference_encoder({k: v.float() for k, v in sd.items()}).to(dev, dtype)
                del sd
                reference_encode_moments(renc, x)
                torch.cuda.synchronize()
                r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                r0.record()
                rm = reference_encode_moments(renc, x)
                r1.record()
                torch.cuda.synchronize()
                leg["gpu_reference"] = {"encode_ms_per_clip": r0.elapsed_time(r1), "star_over_reference": r0.elapsed_time(r1) / ms,
                                        "rel_l2_star_vs_reference_same_dtype": ((m.float() - rm.float()).norm() / rm.float().norm()).item()}
        except Exception as e:
            leg["gpu_reference"] = {"unavailable": repr(e)[:200]}
        return leg
    except Exception as e:                                                # a leg must never break the bench line
        return {"unavailable": repr(e)[:300]}


def vae3d_leg(dev, dtype, T, H, W, step_ms):
    """the decode tail of config 4 (sample_sr.py:206-230): 13 latent frames of 60x90 -> 49 frames of 480x720 through the 3-D
    causal VAE decoder in the reference's chunk protocol (3 + 5 x 2 latent frames), causal-conv context kept on the GPU"""
    from star_b200.cogvideox.vae3d import ContextParallelDecoder3D
    from star_b200.utils.synth import synth_tensor
    with torch.device("meta"):
        dec = ContextParallelDecoder3D()
    sd = {k: synth_tensor(k, v.shape, 11, dev) + (1.0 if ".conv_y.conv.bias" in k else 0.0) for k, v in dec.state_dict().items()}
    dec.load_state_dict(sd, assign=True)
    dec = dec.to(dtype).eval()
    z = torch.randn(1, 16, T, H, W, generator=torch.Generator().manual_seed(5)).to(dev, dtype)
    frames = dec.decode_latent(z)                                   # warm-up (packs the weights)
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    n0 = ops.launch_count()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(2):
        frames = dec.decode_latent(z)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 2
    nf = frames.shape[2]
    # 27-tap convs dominate: 2 * rows * 27 * Cin * Cout summed over the decoder, per clip
    flops = 0.0
    chans = {3: (512, 512), 2: (512, 256), 1: (256, 256), 0: (256, 128)}
    for (t_l, first) in [(3, True)] + [(2, False)] * ((T - 1) // 2 - 1):
        tt, hh, ww = t_l, H, W
        flops += 2.0 * tt * hh * ww * 27 * 64 * 512 + 2 * 2 * 2.0 * tt * hh * ww * 27 * 512 * 512
        for lvl in (3, 2, 1, 0):
            cin, cout = chans[lvl]
            rows = tt * hh * ww
            flops += 2.0 * rows * 27 * (cin * cout + cout * cout) + 3 * 2 * 2.0 * rows * 27 * cout * cout
            if lvl:
                if lvl >= 2:
                    tt = 2 * tt - 1 if (tt % 2 == 1 and tt > 1) else 2 * tt
                hh, ww = 2 * hh, 2 * ww
                flops += 2.0 * tt * hh * ww * 9 * cout * cout
        flops += 2.0 * tt * hh * ww * 27 * 128 * 3
    # the reference's own decoder (unmodified cp_enc_dec.py from the staged tree) on the same GPU, same latent, same dtype
    gpu_ref = None
    try:
        from oracle.cogvideox_vae import build_reference_decoder, reference_decode_latent, vae_reference_available
        if vae_reference_available():
            rdec = build_reference_decoder({k: v.float() for k, v in sd.items()}).to(dev, dtype)
            del sd
            reference_decode_latent(rdec, z)                              # warm-up (cuDNN autotune)
            torch.cuda.synchronize()
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0.record()
            rframes = reference_decode_latent(rdec, z)
            r1.record()
            torch.cuda.synchronize()
            rms = r0.elapsed_time(r1)
            diff = ((frames.float() - rframes.float()).norm() / rframes.float().norm()).item()
            gpu_ref = {"decode_ms_per_clip": rms, "star_over_reference": rms / ms, "rel_l2_star_vs_reference_same_dtype": diff,
                       "how": "ContextParallelDecoder3D of the reference tree, " + str(dtype) + ", cuDNN Conv3d, its CPU cache round trip included"}
            del rdec, rframes
        else:
            gpu_ref = {"unavailable": "oracle/_ref not staged on this box"}
    except Exception as e:                                                # the baseline must never break the bench line
        gpu_ref = {"unavailable": repr(e)[:200]}
    finite = bool(torch.isfinite(frames.float()).all())
    launches_per_clip = int((ops.launch_count() - n0) // 2)
    peak_gb = torch.cuda.max_memory_allocated() / 2 ** 30
    del frames
    torch.cuda.empty_cache()
    enc_leg = vae3d_encode_leg(dev, dtype, (T - 1) * 4 + 1, 8 * H, 8 * W)
    return {"what": "CogVideoX 3-D causal VAE decode of the clip (SURVEY 8 f4), reference chunk protocol, context frames on the GPU",
            "encode": enc_leg,
            "gpu_reference": gpu_ref,
            "frames": int(nf), "decode_ms_per_clip": ms, "decode_ms_per_frame": ms / nf, "decode_tflops_per_s": flops / ms / 1e9,
            "algorithmic_tflop_per_clip": flops / 1e12, "peak_mem_gb": peak_gb,
            "launches_per_clip": launches_per_clip, "finite": finite,
            "frames_per_s_denoise_plus_decode": nf / ((50 * step_ms + ms) / 1e3),
            "frames_per_s_encode_denoise_decode": (nf / ((50 * step_ms + ms + enc_leg["encode_ms_per_clip"]) / 1e3)) if enc_leg.get("encode_ms_per_clip") else None,
            "parity": "tests/test_cogvideox_vae.py (reference's unmodified cp_enc_dec.py)"}


def run_config4(args):
    """bench.py --workload cogvideox: BASELINE config 4 -- CogVideoX-5B heavy-deg 4x, 49 frames 720x480 (latent 13 x 60 x 90,
    patch 2 -> 17 550 image + 226 text tokens), the whole 42-layer DiffusionTransformer with LoRA r = 512 merged, bf16 (the
    reference's dtype), CFG pair as batch 2.  A "step" = one step of the shipped sampler (star_b200/cogvideox/sampling.py:
    DiT forward of the CFG pair, VideoScaling, DynamicCFG, DPM-Solver++(2M) SDE update)."""
    import json
    import torch.distributed as dist
    from bench import ClockSampler, measured_peaks
    from star_b200.cogvideox import DiffusionTransformer
    from star_b200.utils.synth import synth_tensor
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        if world != 2:
            raise SystemExit("bench.py --workload cogvideox: one clip splits over exactly 2 GPUs (CFG pair); more GPUs are replicas")
    dtype = torch.bfloat16
    layers = 4 if args.small else 42
    with torch.device("meta"):
        net = DiffusionTransformer(num_layers=layers, dtype=dtype)
    sd = {}
    for k, v in net.state_dict().items():
        t = synth_tensor(k, v.shape, 7, dev)
        sd[k] = t * (0.5 if "local.conv1" in k else 1.0)
    net.load_state_dict(sd, assign=True)
    net._pack()
    del sd
    for p in list(net.parameters()):                      # packed bf16 copies are what runs; drop the fp32 masters
        p.data = torch.empty(0, device=dev)
    torch.cuda.empty_cache()
    T, H, W, tl = 13, 60, 90, 226
    from star_b200.cogvideox.sampling import VPSDEDPMPP2MSampler
    sampler = VPSDEDPMPP2MSampler(num_steps=50, dtype=dtype)      # the shipped schedule / DynamicCFG / VideoScaling (yaml :20-33,:157-174)
    st_mid = sampler.plan.steps[25]                              # a mid-schedule step: two noise draws, 2M history term
    torch.manual_seed(0)                                         # same solver noise on every rank
    g = torch.Generator().manual_seed(0)
    # CFG: split the pair across 2 ranks when available (the only parallel axis of one clip); else batch 2
    split = world >= 2
    bsz = 1 if split else 2
    x = torch.randn(1, T, 16, H, W, generator=g)                 # noisy latent (fp32, sampler state)
    lq1 = torch.randn(1, T, 16, H, W, generator=g)               # VAE-encoded LQ clip
    ctx = torch.randn(2, tl, 4096, generator=g)                  # [uncond, cond] T5 embeddings
    x_pin, lq_pin, ctx_pin = x.pin_memory(), lq1.pin_memory(), ctx.pin_memory()
    xd, lqd, cd = x.to(dev), torch.cat((lq1, lq1), 0).to(dev), ctx.to(dev, dtype)
    old = torch.randn(1, T, 16, H, W, generator=g).to(dev)       # previous step's denoised latent

    from star_b200.cogvideox.sampling import split_cfg_pair
    net_pair = split_cfg_pair(net) if split else net          # one branch per rank + one all-gather of the 1.4 MB prediction

    def step(xx, lq2, cc):
        return sampler.step(net_pair, xx, old, st_mid, cc, lq2)[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(max(args.warmup, 1)):
        step(xd, lqd, cd)
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    n0 = ops.launch_count()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        out = step(xd, lqd, cd)
    b.record()
    barrier()
    launches = ops.launch_count() - n0
    ms = a.elapsed_time(b) / args.steps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        lq_d = lq_pin.to(dev, non_blocking=True)
        res = step(x_pin.to(dev, non_blocking=True), torch.cat((lq_d, lq_d), 0), ctx_pin.to(dev, non_blocking=True).to(dtype)).cpu()
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        tt = torch.tensor([ms, e2e_ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, e2e_ms = tt.tolist()
    ops.trace_begin()
    step(xd, lqd, cd)
    trace = ops.trace_end()
    att = [t_ms for name, sig, t_ms in trace if name == "attention"]
    per_op = {}
    for name, _sig, t_ms in trace:
        per_op[name] = per_op.get(name, 0.0) + t_ms
    peaks = measured_peaks()
    S = tl + T * (H // 2) * (W // 2)
    aflops = 4.0 * S * S * 64 * 48 * bsz
    d, ff = 3072, 4 * 3072
    layer_flops = 2 * (2.0 * S * d * 3 * d + 2.0 * S * d * d + 4.0 * S * d * ff) + 2 * 4.0 * S * S * d   # both CFG branches
    vae_leg = None
    if world == 1 and not args.small:
        vae_leg = vae3d_leg(dev, dtype, T, H, W, ms)
    if rank == 0:
        line = {"metric": "upscaled frames/sec, CogVideoX-5B heavy-deg 4x, 49-frame 720x480, 50 steps", "value": 49.0 / (50 * ms / 1e3),
                "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                "higher_is_better": True, "scaling": "strong" if split and world == 2 else "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic",
                "config": {"workload": "BASELINE config 4: CogVideoX-5B DiT, 49 frames 720x480 (latent 13x60x90, 17 776 tokens), 50 steps, CFG pair",
                           "model": f"DiffusionTransformer {layers} layers x 3072, 48 heads, LoRA r=512 merged, synthetic weights"
                                    + (" (REDUCED DEPTH, debug)" if args.small else ""),
                           "parallelism": "CFG pair split over 2 ranks, one all-gather of the 1.4 MB prediction per step" if split else "single GPU, CFG pair as batch 2",
                           "step": "one VPSDEDPMPP2MSampler step (index 25 of 50): DiT forward of the CFG pair on cat(noisy, LQ) latents, VideoScaling, DynamicCFG, DPM-Solver++(2M) SDE update with its two noise draws",
                           "l2": "activations (109 MB per 3072-wide token matrix per sample) exceed the L2 across a layer; no explicit flush"},
                "clocks": clocks.stop(), "gpu_launches": int(launches),
                "e2e": {"value": 49.0 / (50 * e2e_ms / 1e3), "unit": "frames/s", "ms_per_step": e2e_ms,
                        "h2d_bytes_per_step": int((x.numel() + lq1.numel() + ctx.numel()) * 4), "d2h_bytes_per_step": int(res.numel() * 4),
                        "api": "VPSDEDPMPP2MSampler.step(host latent, LQ latent, text embeddings) -> next latent .cpu()"},
                "roofline": {"kernel": "attn4_fwd_kernel (3-D full attention, 48 heads, N = 17 776)", "bound": "tensor",
                             "achieved": aflops / (sum(att) / max(len(att), 1) * 1e-3) / 1e12 if att else None, "peak": peaks["tflops"],
                             "unit": "TFLOP/s", "frac": (aflops / (sum(att) / max(len(att), 1) * 1e-3) / 1e12 / peaks["tflops"]) if att else None,
                             "algorithmic_flops_per_launch": aflops, "launches_timed": len(att), "traffic": None},
                "model_tflops_per_s": layers * layer_flops * (bsz / 2.0) / (ms * 1e-3) / 1e12,
                "op_time_share": {k: round(v / sum(per_op.values()), 4) for k, v in sorted(per_op.items(), key=lambda kv: -kv[1])},
                "out_checksum": {"finite": bool(torch.isfinite(out).all()), "abs_mean": float(out.abs().mean())},
                "pipeline": vae_leg,
                "parity": "tests/test_cogvideox.py::test_dit_model_gpu_vs_reference_files (reference files behind the sat shim)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
