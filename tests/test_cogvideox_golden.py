"""CogVideoX path against committed golden vectors (tests/golden/cogvideox_path.pt, written by oracle/make_golden_cogvideox.py from
the UNMODIFIED reference files): these run with no reference tree at all -- CPU (host graph on emulated kernels) and GPU."""
import os

import pytest
import torch

from tests.util import rel_l2

GOLD = os.path.join(os.path.dirname(__file__), "golden", "cogvideox_path.pt")


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD)


def _models(device, dtype):
    from oracle.make_golden_cogvideox import SMALL
    from star_b200.cogvideox.vae3d import ContextParallelDecoder3D, ContextParallelEncoder3D
    from star_b200.utils.synth import synth_state_dict
    dec, enc = ContextParallelDecoder3D(**SMALL), ContextParallelEncoder3D(**SMALL)
    sd = synth_state_dict({k: tuple(v.shape) for k, v in dec.state_dict().items()}, seed=3)
    for k in sd:
        if ".conv_y.conv.bias" in k:
            sd[k] = sd[k] + 1.0
    dec.load_state_dict(sd)
    enc.load_state_dict(synth_state_dict({k: tuple(v.shape) for k, v in enc.state_dict().items()}, seed=6))
    return dec.to(device=device, dtype=dtype).eval(), enc.to(device=device, dtype=dtype).eval()


def _check_vae(gold, device, tol):
    from oracle.make_golden_cogvideox import vae_inputs
    dec, enc = _models(device, torch.float16)
    z, x = vae_inputs()
    e_dec = rel_l2(dec.decode_latent(z.to(device)).cpu(), gold["vae_dec"])
    e_enc = rel_l2(enc(x.to(device)).cpu(), gold["vae_enc"])
    print(f"[cogvideox golden, {device}] decoder {e_dec:.2e} encoder {e_enc:.2e}")
    assert e_dec < tol and e_enc < tol


def test_vae3d_vs_golden_on_emulated_kernels(gold, monkeypatch):
    from oracle import kernel_ref as KR
    from star_b200 import ops
    for name in dir(KR):
        if not name.startswith("_") and callable(getattr(KR, name)) and hasattr(ops, name):
            monkeypatch.setattr(ops, name, getattr(KR, name))
    _check_vae(gold, "cpu", 3e-3)


def test_sampler_vs_golden(gold):
    from oracle.make_golden_cogvideox import sampler_inputs
    from star_b200.cogvideox.sampling import StepPlan, VPSDEDPMPP2MSampler
    from tests.test_cogvideox_sampler import FakeDiT
    plan = StepPlan()
    assert torch.equal(plan.alphas_cumprod_sqrt, gold["acs"]) and plan.timesteps == gold["timesteps"]
    assert [st.c_skip for st in plan.steps] == [float(v) for v in gold["sigma_q"]]
    assert [st.cfg_scale for st in plan.steps] == gold["cfg"]
    lq, randn, cond, uc = sampler_inputs()
    torch.manual_seed(123)
    got = VPSDEDPMPP2MSampler(num_steps=6, dtype=torch.float32)(FakeDiT(), randn.clone(), cond, uc=uc, lq=torch.cat((lq, lq), 0))
    assert rel_l2(got, gold["run6"]) < 2e-5


@pytest.mark.gpu
def test_vae3d_vs_golden_gpu(gold):
    _check_vae(gold, "cuda", 3e-3)
