"""TEST INFRASTRUCTURE -- execute the reference's CogVideoX sampler stack UNMODIFIED (cogvideox-based/sat/sgm/modules/diffusionmodules):
    sampling.py (VPSDEDPMPP2MSampler), guiders.py (DynamicCFG), denoiser.py (DiscreteDenoiser), denoiser_scaling.py (VideoScaling),
    denoiser_weighting.py, discretizer.py (ZeroSNRDDPMDiscretization), sampling_utils.py, util.py (make_beta_schedule), wrappers.py
loaded by path from the reference tree (or the staged oracle/_ref).  Shimmed imports: ``omegaconf`` (type names only), ``sgm.util``
(append_dims / default / append_zero / instantiate_from_config restated -- four one-liners of sgm/util.py:233-283; SeededNoise unused).
PINNING STATUS: star_b200/cogvideox/sampling.py is pinned to the reference's own files through this module.
Never imported by the product.
"""
import importlib
import importlib.util
import os
import sys
import types
from inspect import isfunction

import torch

from . import ref_loader

_ns = {}
_DIR = ("cogvideox-based", "sat", "sgm", "modules", "diffusionmodules")
FILES = ["util", "sampling_utils", "guiders", "discretizer", "denoiser_scaling", "denoiser_weighting", "denoiser", "wrappers", "sampling"]


def _mod(name):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        sys.modules[name] = m
        parent, _, leaf = name.rpartition(".")
        if parent:
            setattr(_mod(parent), leaf, m)
    return m


def sampler_reference_available():
    return os.path.isfile(os.path.join(ref_loader.REF_ROOT, *_DIR, "sampling.py"))


def instantiate_from_config(config, **extra):
    module, cls = config["target"].rsplit(".", 1)
    return getattr(sys.modules[module] if module in sys.modules else importlib.import_module(module), cls)(
        **config.get("params", {}), **extra)


def load_reference_sampler_modules():
    if "mods" in _ns:
        return _ns["mods"]
    if "omegaconf" not in sys.modules:
        try:
            import omegaconf  # noqa: F401
        except ImportError:
            oc = _mod("omegaconf")
            oc.ListConfig = type("ListConfig", (list,), {})
            oc.OmegaConf = type("OmegaConf", (dict,), {})
    u = _mod("sgm.util")
    u.append_dims = lambda x, target_dims: x[(...,) + (None,) * (target_dims - x.ndim)]
    u.default = lambda val, d: val if val is not None else (d() if isfunction(d) else d)
    u.append_zero = lambda x: torch.cat([x, x.new_zeros([1])])
    u.instantiate_from_config = instantiate_from_config
    u.SeededNoise = type("SeededNoise", (), {})
    mods = {}
    for f in FILES:
        name = "sgm.modules.diffusionmodules." + f
        spec = importlib.util.spec_from_file_location(name, os.path.join(ref_loader.REF_ROOT, *_DIR, f + ".py"))
        m = importlib.util.module_from_spec(spec)
        _mod("sgm.modules.diffusionmodules")
        sys.modules[name] = m
        setattr(sys.modules["sgm.modules.diffusionmodules"], f, m)
        spec.loader.exec_module(m)
        mods[f] = m
    _ns["mods"] = mods
    return mods


def build_reference_sampler(num_steps=50, device="cpu"):
    """(sampler, denoiser) exactly as configs/cogvideox_5b/cogvideox_5b_infer_sr.yaml:20-33,:157-174 instantiates them"""
    mods = load_reference_sampler_modules()
    P = "sgm.modules.diffusionmodules."
    disc = {"target": P + "discretizer.ZeroSNRDDPMDiscretization", "params": {"shift_scale": 1.0}}
    sampler = mods["sampling"].VPSDEDPMPP2MSampler(
        num_steps=num_steps, verbose=False, device=device, discretization_config=disc,
        guider_config={"target": P + "guiders.DynamicCFG", "params": {"scale": 6, "exp": 5, "num_steps": 50}})
    denoiser = mods["denoiser"].DiscreteDenoiser(
        num_idx=1000, quantize_c_noise=False, weighting_config={"target": P + "denoiser_weighting.EpsWeighting"},
        scaling_config={"target": P + "denoiser_scaling.VideoScaling"}, discretization_config=disc)
    return sampler, denoiser


@torch.no_grad()
def reference_sample(network, sampler, denoiser, randn, cond, uc, lq_latent, dtype=torch.float32):
    """SATVideoDiffusionEngine.sample_sr after the LQ encode (diffusion_video.py:270-292): OpenAIWrapper around the network,
    the denoiser lambda, the LQ latent doubled for CFG"""
    mods = load_reference_sampler_modules()
    model = mods["wrappers"].OpenAIWrapper(network, dtype=dtype)

    def den(inp, sigma, c, **kw):
        return denoiser(model, inp, sigma, c, concat_images=None, **kw)
    lq = torch.cat((lq_latent, lq_latent), dim=0)
    return sampler(den, randn, cond, uc=uc, scale=None, scale_emb=None, lq=lq)
