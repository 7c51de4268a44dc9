#!/bin/bash
# round-2 job 10: role timeline of the short-K persistent GEMM (trace variant)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 300 python tools/gemm_trace.py --lib tools/variants/libstar_trace.so 320 960 0 > gpurun_out/j10_trace_qkv.log 2>&1; echo "rc=$?"
timeout 300 python tools/gemm_trace.py --lib tools/variants/libstar_trace.so 320 2560 1 > gpurun_out/j10_trace_geglu.log 2>&1; echo "rc=$?"
cat gpurun_out/j10_trace_qkv.log | head -120
