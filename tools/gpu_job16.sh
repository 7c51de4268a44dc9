#!/bin/bash
# round-2 job 16 (2 GPUs): tools/mgpu_check.py with its full output
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 tools/mgpu_check.py > gpurun_out/j16_mgpu.out 2> gpurun_out/j16_mgpu.err
echo "rc=$?"; grep -v "^$" gpurun_out/j16_mgpu.out | tail -12; grep -n "Error\|error\|Traceback" -A6 gpurun_out/j16_mgpu.err | grep -v "errors.html\|error_file" | head -60
python -c "import torch; torch.zeros(1, device='cuda'); torch.cuda.synchronize(); print('GPU answers')"
