"""TEST INFRASTRUCTURE -- load the UNMODIFIED reference (NJU-PCALab/STAR) hot path.

Usable where /root/reference exists (the build container) or where oracle/stage_reference.py
has staged the same unmodified files into the git-ignored oracle/_ref/ (the GPU box: GPU-side
reference baseline and config-2 parity).  No reference source enters the repository: the files
are executed where they lie, after five tiny import shims for packages absent from this image:

  xformers.ops.memory_efficient_attention -> F.scaled_dot_product_attention
      (call sites: video_to_video/modules/unet_v2v.py:179,184; semantics are
      plain softmax(QK^T/sqrt(d))V with no mask/bias, so SDPA is an exact
      mathematical stand-in)
  fairscale.nn.checkpoint.checkpoint_wrapper -> identity (unet_v2v.py:13)
  timm.models.vision_transformer.Mlp -> empty module (only CaptionEmbedder, unused)
  torchsde.BrownianTree -> injectable seeded sampler (solvers_sdedit.py:94-97)
  easydict.EasyDict -> dict with attribute access (utils/config.py:9)
"""
import importlib
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")     # oracle/stage_reference.py (git-ignored)


def _find_root():
    env = os.environ.get("STAR_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", _STAGED):
        if os.path.isfile(os.path.join(cand, "video_to_video/modules/unet_v2v.py")):
            return cand
    return "/root/reference"


REF_ROOT = _find_root()


def reference_available():
    return os.path.isfile(os.path.join(REF_ROOT, "video_to_video/modules/unet_v2v.py"))


class InjectedBrownian:
    """Stand-in for torchsde.BrownianTree: W(t1)-W(t0) drawn from a shared,
    seeded stream so that two samplers given the same seed consume identical
    increments (north_star: parity on identical noise)."""

    seed = 1234

    def __init__(self, t0, w0, t1, entropy=None, **kw):
        self.shape, self.dtype, self.device = w0.shape, w0.dtype, w0.device
        self.gen = torch.Generator().manual_seed(int(InjectedBrownian.seed))

    def __call__(self, t0, t1):
        n = torch.randn(self.shape, generator=self.gen, dtype=torch.float32)
        return n.to(self.device, self.dtype) * (torch.as_tensor(t1) - torch.as_tensor(t0)).abs().sqrt().to(self.device)


def _exact_attention_fp32(q, k, v, budget_bytes=8 << 30):
    """softmax(q k^T / sqrt(d)) v in true fp32 (no fused kernel, whose fp32 path may run TF32 tensor ops):
    batched matmuls over query-row chunks sized so the score block stays under ``budget_bytes``."""
    Bh, Nq, d = q.shape
    Nk = k.shape[1]
    rows = max(1, min(Nq, budget_bytes // (4 * Bh * Nk)))
    if rows < 16:                                   # many small heads (temporal attention): chunk the batch instead
        out = torch.empty_like(q)
        step = max(1, budget_bytes // (4 * Nq * Nk))
        for b in range(0, Bh, step):
            s = torch.baddbmm(q.new_zeros(()), q[b:b + step], k[b:b + step].transpose(1, 2), beta=0, alpha=d ** -0.5)
            out[b:b + step] = torch.bmm(torch.softmax(s, dim=-1), v[b:b + step])
        return out
    out = torch.empty_like(q)
    kt = k.transpose(1, 2)
    for r in range(0, Nq, rows):
        s = torch.bmm(q[:, r:r + rows], kt) * (d ** -0.5)
        out[:, r:r + rows] = torch.bmm(torch.softmax(s, dim=-1), v)
    return out


def _mea(q, k, v, attn_bias=None, op=None):
    assert attn_bias is None
    if q.is_cuda and q.dtype == torch.float32:
        return _exact_attention_fp32(q, k, v)
    return F.scaled_dot_product_attention(q[None], k[None], v[None])[0]


def _install_shims():
    def mod(name):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            sys.modules[name] = m
        return m

    xf = mod("xformers"); xo = mod("xformers.ops")
    xo.memory_efficient_attention = _mea
    xf.ops = xo
    fs = mod("fairscale"); fsn = mod("fairscale.nn"); fsc = mod("fairscale.nn.checkpoint")
    fsc.checkpoint_wrapper = lambda m, *a, **k: m
    fs.nn = fsn; fsn.checkpoint = fsc
    tm = mod("timm"); tmm = mod("timm.models"); tmv = mod("timm.models.vision_transformer")

    class Mlp(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

    tmv.Mlp = Mlp
    tm.models = tmm; tmm.vision_transformer = tmv
    ts = mod("torchsde")
    ts.BrownianTree = InjectedBrownian
    ed = mod("easydict")

    class EasyDict(dict):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            self.__dict__ = self

    ed.EasyDict = EasyDict


_cache = {}


def load_reference():
    """Returns a namespace with .unet (unet_v2v module), .diffusion_sdedit,
    .solvers_sdedit, .schedules_sdedit and the pure helpers of
    video_to_video_model.py (pad_to_fit, make_chunks, sliding_windows_1d)."""
    if "ns" in _cache:
        return _cache["ns"]
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    _install_shims()
    ns = types.SimpleNamespace()
    spec = importlib.util.spec_from_file_location(
        "ref_unet_v2v", os.path.join(REF_ROOT, "video_to_video/modules/unet_v2v.py"))
    ns.unet = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ns.unet)
    # the diffusion package imports video_to_video.utils.logger -> needs the
    # package on sys.path; video_to_video/__init__.py is empty so this does not
    # pull open_clip / diffusers.
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    ns.diffusion_sdedit = importlib.import_module("video_to_video.diffusion.diffusion_sdedit")
    ns.solvers_sdedit = importlib.import_module("video_to_video.diffusion.solvers_sdedit")
    ns.schedules_sdedit = importlib.import_module("video_to_video.diffusion.schedules_sdedit")
    # video_to_video_model.py has a top-level `from diffusers import ...`:
    # exec only its pure helper functions from source text.
    src = open(os.path.join(REF_ROOT, "video_to_video/video_to_video_model.py")).read()
    helpers = "def pad_to_fit" + src.split("def pad_to_fit")[1]
    g = {}
    exec(compile(helpers, "video_to_video_model_helpers", "exec"), g)
    ns.pad_to_fit, ns.make_chunks, ns.sliding_windows_1d = g["pad_to_fit"], g["make_chunks"], g["sliding_windows_1d"]
    _cache["ns"] = ns
    return ns


def build_reference_unet(dim=320, state_dict=None):
    """ControlledV2VUNet exactly as the reference builds it (unet_v2v.py:1712).
    ``dim`` != 320 builds a narrow variant through the base-class ctor args
    (Vid2VidSDUNet.__init__ / VideoControlNet.__init__ accept ``dim``) -- used
    only to keep CPU tests fast."""
    ns = load_reference()
    U = ns.unet
    if dim == 320:
        net = U.ControlledV2VUNet()
    else:
        net = U.ControlledV2VUNet.__new__(U.ControlledV2VUNet)
        U.Vid2VidSDUNet.__init__(net, dim=dim)
        net.VideoControlNet = U.VideoControlNet(dim=dim)
    net.eval()
    if state_dict is not None:
        missing = net.load_state_dict(state_dict, strict=True)
    return net
