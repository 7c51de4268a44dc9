#!/bin/bash
# round-2 job 13: compile-time specialised tapgemm2 epilogues (plain / +residual / GEGLU): parity, kbench, role timeline
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_config2_gpu.py tests/test_unet_gpu.py tests/test_cogvideox.py tests/test_vae.py -q -x -p no:cacheprovider -m gpu > gpurun_out/j13_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/j13_tests.log
for v in default dbufall; do
  if [ $v = default ]; then LIB=""; else LIB="--lib tools/variants/libstar_$v.so"; fi
  echo "=== $v" >> gpurun_out/j13_ab.log
  timeout 300 python tools/kbench.py linear conv $LIB >> gpurun_out/j13_ab.log 2>&1
done
python - <<'PY'
import re,collections
rows=collections.OrderedDict(); cur=None
for line in open('gpurun_out/j13_ab.log'):
    if line.startswith('==='): cur=line.split()[1]; continue
    m=re.match(r'(.{58})\s+([0-9.]+) ms\s+([0-9.]+) TFLOP',line)
    if m: rows.setdefault(m.group(1).strip(),{})[cur]=(float(m.group(2)),float(m.group(3)))
print('%-58s %9s %9s %9s'%('shape','default','TFLOP/s','dbufall'))
for k,v in rows.items(): print('%-58s %9.3f %9.0f %9.3f'%(k,v.get('default',(0,0))[0],v.get('default',(0,0))[1],v.get('dbufall',(0,0))[0]))
PY
timeout 300 python tools/gemm_trace.py --lib tools/variants/libstar_trace.so 320 960 0 > gpurun_out/j13_trace_qkv.log 2>&1
grep -A14 "## epilogue" gpurun_out/j13_trace_qkv.log | head -18; grep "clk per tile" gpurun_out/j13_trace_qkv.log
timeout 300 python tools/gemm_trace.py --lib tools/variants/libstar_trace.so 320 2560 1 > gpurun_out/j13_trace_geglu.log 2>&1
grep -A8 "## epilogue" gpurun_out/j13_trace_geglu.log | head -10; grep "clk per tile" gpurun_out/j13_trace_geglu.log
