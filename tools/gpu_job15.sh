#!/bin/bash
# round-2 job 15: ncu evidence with the final library -- full captures of the attention kernel and of the short-K GEMM, launch list of one solver step
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
K='regex:^(tapgemm|attn|temporal_attn|gn_|layernorm|liem|concat_add|add_kernel|upsample|s2_split|im2col|pad_w36|nchw5|tokens_to|sinusoidal|silu|cfg_|row_gate|softmax_rows|vae_head|bilinear|plane_stats|adain|qk_ln)'
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn4_fwd_kernel --launch-skip 2 --launch-count 1 -f -o gpurun_out/r02_attn4 python tools/kbench.py attention > gpurun_out/j15_ncu_attn.log 2>&1
echo "ncu attn rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tapgemm2_kernel --launch-skip 7 --launch-count 1 -f -o gpurun_out/r02_gemm_qkv python tools/kbench.py linear > gpurun_out/j15_ncu_gemm.log 2>&1
echo "ncu gemm rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tapgemm2_kernel --launch-skip 14 --launch-count 1 -f -o gpurun_out/r02_gemm_geglu python tools/kbench.py linear > gpurun_out/j15_ncu_geglu.log 2>&1
echo "ncu geglu rc=$?"
timeout 1100 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" --launch-skip 3214 --launch-count 3214 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-vae --no-gpu-reference > gpurun_out/j15_ncu_launches.log 2>&1
echo "ncu launches rc=$?"; wc -l gpurun_out/r02_launches.csv
ls -la gpurun_out/*.ncu-rep
nvidia-smi --query-gpu=index,clocks.sm,memory.used --format=csv,noheader
python -c "import torch; torch.zeros(1, device='cuda'); torch.cuda.synchronize(); print('GPU answers')"
