"""TEST INFRASTRUCTURE -- execute the reference's CogVideoX 3-D VAE decoder file UNMODIFIED.

cogvideox-based/sat/vae_modules/cp_enc_dec.py (ContextParallelDecoder3D and everything under it) is loaded by path from the
reference tree (/root/reference, or the staged git-ignored copy oracle/_ref written by oracle/stage_reference.py).  Its three
foreign imports are shimmed with what they are at context-parallel size 1 (yaml :103 `cp_size: 1`):
    sgm.util.get_context_parallel_{group,rank,world_size,group_rank}  -> None, 0, 1, 0
    vae_modules.utils.SafeConv3d                                     -> restated (Conv3d that splits inputs above 2 GB along T with the
                                                                        exact overlap, utils.py:72-91; its module imports fsspec / PIL)
    torch.distributed.get_rank() / get_world_size()                  -> 0 / 1 while the decoder runs (no process group in a test)
PINNING STATUS: the decoder arithmetic of SURVEY row f4 is pinned to the reference's own file through this module.
Never imported by the product.
"""
import importlib.util
import os
import sys
import types
from contextlib import contextmanager
from unittest import mock

import torch

from . import ref_loader

_ns = {}


def _mod(name):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        sys.modules[name] = m
        parent, _, leaf = name.rpartition(".")
        if parent:
            setattr(_mod(parent), leaf, m)
    return m


def vae_reference_available():
    return os.path.isfile(os.path.join(ref_loader.REF_ROOT, "cogvideox-based", "sat", "vae_modules", "cp_enc_dec.py"))


def load_reference_vae():
    """the reference's cp_enc_dec module"""
    if "m" in _ns:
        return _ns["m"]
    path = os.path.join(ref_loader.REF_ROOT, "cogvideox-based", "sat", "vae_modules", "cp_enc_dec.py")
    if not os.path.isfile(path):
        raise RuntimeError("CogVideoX VAE reference file not present under %s" % ref_loader.REF_ROOT)
    u = _mod("sgm.util")
    u.get_context_parallel_group = lambda: None
    u.get_context_parallel_rank = lambda: 0
    u.get_context_parallel_world_size = lambda: 1
    u.get_context_parallel_group_rank = lambda: 0

    class SafeConv3d(torch.nn.Conv3d):
        """restated from vae_modules/utils.py:72-91 (that file imports fsspec / PIL / safetensors and cannot be loaded here):
        inputs above 2 GB are convolved in chunks along T, each chunk prefixed with the kernel_size - 1 frames before it"""

        def forward(self, input):
            gb = torch.prod(torch.tensor(input.shape)).item() * 2 / 1024 ** 3
            if gb <= 2:
                return super().forward(input)
            k = self.kernel_size[0]
            chunks = torch.chunk(input, int(gb / 2) + 1, dim=2)
            if k > 1:
                chunks = [chunks[0]] + [torch.cat((chunks[i - 1][:, :, -k + 1:], chunks[i]), dim=2) for i in range(1, len(chunks))]
            return torch.cat([super(SafeConv3d, self).forward(c) for c in chunks], dim=2)
    _mod("vae_modules.utils").SafeConv3d = SafeConv3d
    spec = importlib.util.spec_from_file_location("vae_modules.cp_enc_dec", path)
    m = importlib.util.module_from_spec(spec)
    sys.modules["vae_modules.cp_enc_dec"] = m
    spec.loader.exec_module(m)
    _ns["m"] = m
    return m


@contextmanager
def single_rank():
    with mock.patch.object(torch.distributed, "get_rank", lambda *a, **k: 0), \
            mock.patch.object(torch.distributed, "get_world_size", lambda *a, **k: 1):
        yield


DECODER_KW = dict(double_z=True, z_channels=16, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=(1, 2, 2, 4),
                  attn_resolutions=[], num_res_blocks=3, dropout=0.0, gather_norm=False)     # cogvideox_5b_infer_sr.yaml:128-141


def build_reference_decoder(state_dict=None, **overrides):
    import contextlib
    import io
    m = load_reference_vae()
    kw = dict(DECODER_KW)
    kw.update(overrides)
    with contextlib.redirect_stdout(io.StringIO()):           # the constructor prints its z shape
        dec = m.ContextParallelDecoder3D(**kw)
    if state_dict is not None:
        dec.load_state_dict(state_dict)
    return dec.eval()


@torch.no_grad()
def reference_decode_latent(dec, latent):
    """the serial chunk loop of sample_sr.py:212-227 on the reference decoder (cache moved through the CPU as it does)"""
    Tl = latent.shape[2]
    loops = (Tl - 1) // 2
    out = []
    with single_rank():
        for i in range(loops):
            a, b = (0, 3) if i == 0 else (2 * i + 1, 2 * i + 3)
            out.append(dec(latent[:, :, a:b].contiguous(), clear_fake_cp_cache=(i == loops - 1)))
    return torch.cat(out, dim=2)


ENCODER_KW = dict(double_z=True, z_channels=16, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=(1, 2, 2, 4),
                  attn_resolutions=[], num_res_blocks=3, dropout=0.0, gather_norm=True)      # cogvideox_5b_infer_sr.yaml:113-126


def build_reference_encoder(state_dict=None, **overrides):
    m = load_reference_vae()
    kw = dict(ENCODER_KW)
    kw.update(overrides)
    enc = m.ContextParallelEncoder3D(**kw)
    if state_dict is not None:
        enc.load_state_dict(state_dict)
    return enc.eval()


@torch.no_grad()
def reference_encode_moments(enc, x):
    with single_rank():
        return enc(x)
