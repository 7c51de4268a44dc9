#!/bin/bash
# round-2 job 14: staging depth once more on the specialised epilogues (linear only), then the N=1 bench line (CFG-pair forward, new GEMM epilogue)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
for v in default dbuf0; do
  if [ $v = default ]; then LIB=""; else LIB="--lib tools/variants/libstar_$v.so"; fi
  echo "=== $v" >> gpurun_out/j14_ab.log
  timeout 300 python tools/kbench.py linear $LIB >> gpurun_out/j14_ab.log 2>&1
done
grep -E "===|linear L0|L1 qkv|nvidia" gpurun_out/j14_ab.log
timeout 1500 python bench.py --steps 5 --warmup 3 --trace-out gpurun_out/j14_optrace.txt > gpurun_out/j14_bench.json 2> gpurun_out/j14_bench.err
echo "bench rc=$?"; tail -2 gpurun_out/j14_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/j14_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['clocks'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'])
print(d['gpu_reference']); print(d['op_time_share'])
PY
