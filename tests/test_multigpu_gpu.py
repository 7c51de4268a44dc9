"""NCCL checks on >= 2 GPUs of one node (skipped on single-GPU boxes): tools/mgpu_check.py under torchrun -- chunk-parallel /
CFG-split denoise on the real kernels, sharded VAE decode and the sharded entry against their single-GPU counterparts."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_nccl_sharded_pipeline(nproc):
    if torch.cuda.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    only = os.environ.get("STAR_MGPU_ONLY")                  # e.g. "8": run just that world size (GPU-minute budget)
    if only and int(only) != nproc:
        pytest.skip("STAR_MGPU_ONLY")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
                        "--master-port", str(29500 + nproc), os.path.join(ROOT, "tools", "mgpu_check.py")], cwd=ROOT, capture_output=True,
                       text=True, timeout=900)
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("== serial loop: True") == nproc
