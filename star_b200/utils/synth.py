"""Deterministic synthetic checkpoints for the STAR denoiser.

There are no pretrained weights in the build/bench environment (no network),
and the reference zero-initialises many layers (zero_module / nn.init.zeros_,
e.g. video_to_video/modules/unet_v2v.py:289-294, :638-639, :1223-1224, :1555,
:2128-2132) so a stock random init exercises almost nothing.  This module
produces a reproducible, fully non-zero state_dict from a {key: shape}
manifest; every tensor depends only on (seed, key) so the oracle and the
product can be given bit-identical weights without sharing a module tree.
"""
import zlib

import torch


def _key_seed(seed, key):
    return (int(seed) * 1000003 + zlib.crc32(key.encode())) & 0x7FFFFFFFFFFF


def synth_tensor(key, shape, seed=0, device="cpu"):
    shape = tuple(int(s) for s in shape)
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(_key_seed(seed, key))
    r = torch.randn(shape, generator=g, device=dev, dtype=torch.float32)
    if len(shape) >= 2:                       # linear / conv kernels: 1/sqrt(fan_in)
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        return r * (1.0 / fan_in ** 0.5)
    if key.endswith("weight"):                # all 1-D weights are norm scales
        return 1.0 + 0.1 * r
    return 0.02 * r                           # biases


def synth_state_dict(manifest, seed=0, device="cpu", dtype=torch.float32):
    """manifest: {key: shape}.  CPU generation is bit-reproducible across
    machines with the same torch build (fixtures rely on it); CUDA generation
    is only used for benchmarking."""
    return {k: synth_tensor(k, s, seed, device).to(dtype) for k, s in manifest.items()}
