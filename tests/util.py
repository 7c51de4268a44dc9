import torch


def rel_l2(got, ref):
    got, ref = got.float(), ref.float()
    return ((got - ref).norm() / ref.norm().clamp_min(1e-12)).item()


def assert_close(got, ref, rel=2e-3, max_rel=2e-2, what=""):
    """fp16-output comparison: relative L2 error and max-abs error relative to max|ref|."""
    assert got.shape == ref.shape, f"{what}: shape {tuple(got.shape)} vs {tuple(ref.shape)}"
    g, r = got.float(), ref.float()
    assert torch.isfinite(g).all(), f"{what}: non-finite values in result"
    e = rel_l2(g, r)
    m = ((g - r).abs().max() / r.abs().max().clamp_min(1e-12)).item()
    assert e <= rel and m <= max_rel, f"{what}: rel-L2 {e:.3e} (tol {rel:.1e}), max-abs/max|ref| {m:.3e} (tol {max_rel:.1e})"
    return e, m
