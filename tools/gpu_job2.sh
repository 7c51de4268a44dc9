#!/bin/bash
# round-2 job 2: full GPU suite after the cleanup (+ config-2-shape tests), TMA store-rate and MUFU f16 micro-benchmarks
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider --durations=15 > gpurun_out/j2_gpu_tests.log 2>&1
echo "gpu tests rc=$?"; tail -4 gpurun_out/j2_gpu_tests.log
timeout 120 tools/micro/tma_store_rate > gpurun_out/j2_tma_store_rate.log 2>&1; echo "tma rc=$?"
cat gpurun_out/j2_tma_store_rate.log
timeout 120 tools/micro/xu_rate > gpurun_out/j2_xu_rate.log 2>&1; echo "xu rc=$?"
cat gpurun_out/j2_xu_rate.log
