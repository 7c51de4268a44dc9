#!/bin/bash
# round-2 job 1: reference on the B200 (baseline + config-2 parity), then the CTA-pair GEMM verdict
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/j1_smi.txt
timeout 600 python tools/ref_gpu.py --small --shape 4,18,16 --out gpurun_out/ref_gpu_small.json > gpurun_out/j1_ref_small.log 2>&1
echo "small rc=$?"
timeout 1200 python tools/ref_gpu.py --shape 32,122,216 --out gpurun_out/ref_gpu_c2.json > gpurun_out/j1_ref_c2.log 2>&1
echo "c2 rc=$?"
tail -3 gpurun_out/j1_ref_c2.log | cut -c1-1500
timeout 300 python tools/kbench.py linear conv > gpurun_out/j1_kbench_base.log 2>&1
echo "kbench base rc=$?"
STAR_GEMM_PAIR=1 timeout 300 python tools/kbench.py linear conv > gpurun_out/j1_kbench_pair1.log 2>&1
echo "kbench pair rc=$?"
STAR_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_experimental_gpu.py -q -x -p no:cacheprovider > gpurun_out/j1_pair_tests.log 2>&1
echo "pair tests rc=$?"
tail -5 gpurun_out/j1_pair_tests.log
