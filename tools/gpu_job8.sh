#!/bin/bash
# round-2 job 8 (2 GPUs): NCCL checks of the sharded pipeline, N=2 bench lines (denoise weak scaling, CogVideoX CFG split)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_multigpu_gpu.py -x -q -s -p no:cacheprovider > gpurun_out/j8_mgpu_tests.log 2>&1
echo "mgpu tests rc=$?"; grep -E "rank|passed|failed|Error" gpurun_out/j8_mgpu_tests.log | tail -14
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/j8_bench_n2.json 2> gpurun_out/j8_bench_n2.err
echo "bench n2 rc=$?"; tail -2 gpurun_out/j8_bench_n2.err; tail -1 gpurun_out/j8_bench_n2.json | cut -c1-1500
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --workload cogvideox --gpus 2 --steps 3 --warmup 3 > gpurun_out/j8_dit_n2.json 2> gpurun_out/j8_dit_n2.err
echo "dit n2 rc=$?"; tail -2 gpurun_out/j8_dit_n2.err; tail -1 gpurun_out/j8_dit_n2.json | cut -c1-900
