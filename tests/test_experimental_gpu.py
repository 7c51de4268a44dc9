"""Opt-in checks of code paths that are compiled but not yet validated on hardware (run with STAR_TEST_EXPERIMENTAL=1).
They are skipped in the regular GPU suite: an experimental kernel must not be able to turn the suite red."""
import os
import subprocess
import sys

import pytest

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("STAR_TEST_EXPERIMENTAL") != "1", reason="set STAR_TEST_EXPERIMENTAL=1 to run")]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("pair", ["1", "2"])
def test_cta_pair_gemm(pair):
    """tapgemm2_pair_kernel (cta_group::2): the linear / conv parity tests in a child process with STAR_GEMM_PAIR set
    (the switch is read once at star_init)."""
    env = dict(os.environ, STAR_GEMM_PAIR=pair, STAR_GEMM_BN256="0" if pair == "2" else "1")
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_kernels_gpu.py", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider",
                        "-k", "linear or conv"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
