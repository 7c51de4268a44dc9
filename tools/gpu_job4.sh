#!/bin/bash
# round-2 job 4: attention A/B -- exp-phase ping-pong and FMA-pipe polynomial share (compile-time variants)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
for v in default pp1 poly4 pp1poly4 pp1poly3; do
  if [ $v = default ]; then LIB=""; else LIB="--lib tools/variants/libstar_$v.so"; fi
  echo "=== $v" >> gpurun_out/j4_attn_ab.log
  timeout 300 python tools/kbench.py attention $LIB >> gpurun_out/j4_attn_ab.log 2>&1
done
cat gpurun_out/j4_attn_ab.log | grep -E "===|attention self"
for v in pp1 pp1poly4; do
  STAR_LIB_VARIANT=tools/variants/libstar_$v.so timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_config2_gpu.py -q -x -p no:cacheprovider -k "attention" > gpurun_out/j4_tests_$v.log 2>&1
  echo "tests $v rc=$?"; tail -2 gpurun_out/j4_tests_$v.log
done
