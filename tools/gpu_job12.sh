#!/bin/bash
# round-2 job 12 (8 GPUs): NCCL checks on 8 ranks (CFG split), the N=8 bench line with config 3 and the sharded VAE legs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
nvidia-smi -L | wc -l
STAR_MGPU_ONLY=8 timeout 900 python -m pytest tests/test_multigpu_gpu.py -x -q -s -p no:cacheprovider > gpurun_out/j12_mgpu_tests.log 2>&1
echo "mgpu tests rc=$?"; grep -o "\[rank [0-9]\][^\\\\]*" gpurun_out/j12_mgpu_tests.log | sort -u | head -40; tail -3 gpurun_out/j12_mgpu_tests.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/j12_bench_n8.json 2> gpurun_out/j12_bench_n8.err
echo "bench n8 rc=$?"; tail -3 gpurun_out/j12_bench_n8.err; tail -1 gpurun_out/j12_bench_n8.json | cut -c1-3000
nvidia-smi --query-gpu=index,clocks.sm,memory.used --format=csv,noheader | head -8
python -c "import torch; [torch.zeros(1, device=f'cuda:{i}') for i in range(torch.cuda.device_count())]; torch.cuda.synchronize(); print('all GPUs answer')"
