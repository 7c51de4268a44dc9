#!/bin/bash
# round-2 job 3: new kernels / pipeline-vs-oracle tests, then the N=1 bench line with the gpu_reference block
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider -k "bilinear or cfg_x0 or pixel_pipeline or unet or sampler" -s > gpurun_out/j3_tests.log 2>&1
echo "tests rc=$?"; grep -E "passed|failed|rel-L2|Error" gpurun_out/j3_tests.log | tail -15
timeout 1500 python bench.py --steps 3 --warmup 3 --trace-out gpurun_out/j3_optrace.txt > gpurun_out/j3_bench.json 2> gpurun_out/j3_bench.err
echo "bench rc=$?"; tail -c 3000 gpurun_out/j3_bench.json; tail -5 gpurun_out/j3_bench.err
