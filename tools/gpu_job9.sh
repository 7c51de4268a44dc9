#!/bin/bash
# round-2 job 9 (2 GPUs): re-run of the NCCL checks after the upload fix; new single-GPU tests ride along on GPU 0
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu_gpu.py -x -q -s -p no:cacheprovider > gpurun_out/j9_mgpu_tests.log 2>&1
echo "mgpu tests rc=$?"; grep -o "\[rank [0-9]\][^\\\\]*" gpurun_out/j9_mgpu_tests.log | sort -u | head -12; tail -3 gpurun_out/j9_mgpu_tests.log
CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider -s -k "adain or pixel_pipeline or cfg_x0 or bilinear" > gpurun_out/j9_tests.log 2>&1
echo "tests rc=$?"; grep -E "passed|failed|rel-L2|Error" gpurun_out/j9_tests.log | tail -6
