#!/bin/bash
# round-2 job 6: bf16 library + CogVideoX DiT vs the reference files (sat shim), config-4 bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cogvideox.py tests/test_kernels_gpu.py -m gpu -x -q -p no:cacheprovider -s -k "dit or bf16 or linear_ex or qk_ln" > gpurun_out/j6_tests.log 2>&1
echo "tests rc=$?"; grep -E "passed|failed|rel-L2|Error|error" gpurun_out/j6_tests.log | tail -12
timeout 600 python bench.py --workload cogvideox --small --steps 2 --warmup 1 > gpurun_out/j6_dit_small.json 2> gpurun_out/j6_dit_small.err
echo "small rc=$?"; tail -3 gpurun_out/j6_dit_small.err; cut -c1-600 gpurun_out/j6_dit_small.json
timeout 1200 python bench.py --workload cogvideox --steps 3 --warmup 3 > gpurun_out/j6_dit_config4.json 2> gpurun_out/j6_dit_config4.err
echo "config4 rc=$?"; tail -3 gpurun_out/j6_dit_config4.err; cut -c1-3000 gpurun_out/j6_dit_config4.json
