"""One tiny invocation of the hot path on cuda:0, checked against the oracle (used by
__graft_entry__.smoke()).  The oracle import lives here on purpose: smoke() is one of the three
places allowed to use it (as the checker)."""
import torch


def run_smoke():
    if not torch.cuda.is_available():
        raise RuntimeError("smoke() needs a CUDA device: star_b200 has no CPU path")
    from oracle.unet_ref import UNetCfg, controlled_unet_forward
    from star_b200 import ops
    from star_b200.utils.synth import synth_state_dict
    from star_b200.video_to_video.modules.unet_v2v import ControlledV2VUNet

    kw = dict(dim_mult=[1, 2, 1, 4], num_res_blocks=1)
    with torch.device("meta"):
        net = ControlledV2VUNet(**kw)
    sd = synth_state_dict({k: tuple(v.shape) for k, v in net.state_dict().items()}, seed=1)
    net.load_state_dict(sd, assign=True)
    net = net.half().eval().to("cuda:0")
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 4, 4, 18, 16, generator=g)
    hint = 0.5 * torch.randn(1, 4, 4, 18, 16, generator=g)
    y = torch.randn(1, 77, 1024, generator=g)
    t = torch.tensor([899])
    n0 = ops.launch_count()
    out = net(x.cuda(), t.cuda(), y.cuda(), hint=hint.cuda())
    torch.cuda.synchronize()
    launches = ops.launch_count() - n0
    ref = controlled_unet_forward(sd, x, t, y, hint, UNetCfg(**kw))
    err = ((out.float().cpu() - ref).norm() / ref.norm()).item()
    print(f"smoke: ControlledV2VUNet forward on cuda:0, {launches} star kernels, rel-L2 vs fp32 oracle = {err:.3e}")
    if not (err < 4e-3) or launches == 0:
        raise RuntimeError(f"smoke failed: rel-L2 {err:.3e}, launches {launches}")
    # the solver step's CFG pair (shared text-independent prefix) must reproduce the single forwards bit for bit
    pa, pb = net.forward_cfg_pair(x.cuda(), t.cuda(), (y.cuda(), -y.cuda()), hint=hint.cuda())
    other = net(x.cuda(), t.cuda(), -y.cuda(), hint=hint.cuda())
    torch.cuda.synchronize()
    if not (torch.equal(pa, out) and torch.equal(pb, other)):
        raise RuntimeError("smoke failed: forward_cfg_pair differs from two forwards")
