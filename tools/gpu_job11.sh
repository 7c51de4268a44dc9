#!/bin/bash
# round-2 job 11: tapgemm2 with a dedicated TMA-store thread (mbarrier hand-off, no bar.sync in the epilogue): parity, A/B of
# the staging depth, role timeline; CFG-pair forward on the real kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_config2_gpu.py tests/test_unet_gpu.py -q -x -p no:cacheprovider -s -k "linear or conv or unet or cfg_pair or single_head" > gpurun_out/j11_tests.log 2>&1
echo "tests rc=$?"; grep -E "passed|failed|launches|Error" gpurun_out/j11_tests.log | tail -5
STAR_LIB_VARIANT=tools/variants/libstar_st_dbufall.so timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_config2_gpu.py -q -x -p no:cacheprovider -k "linear or conv" > gpurun_out/j11_tests_dbufall.log 2>&1
echo "tests dbufall rc=$?"; tail -2 gpurun_out/j11_tests_dbufall.log
for v in default st_dbuf0 st_dbufall; do
  if [ $v = default ]; then LIB=""; else LIB="--lib tools/variants/libstar_$v.so"; fi
  echo "=== $v" >> gpurun_out/j11_ab.log
  timeout 300 python tools/kbench.py linear conv $LIB >> gpurun_out/j11_ab.log 2>&1
done
python - <<'PY'
import re,collections
rows=collections.OrderedDict(); cur=None
for line in open('gpurun_out/j11_ab.log'):
    if line.startswith('==='): cur=line.split()[1]; continue
    m=re.match(r'(.{58})\s+([0-9.]+) ms',line)
    if m: rows.setdefault(m.group(1).strip(),{})[cur]=float(m.group(2))
print('%-58s %9s %9s %9s'%('shape','default','st_dbuf0','st_dbufall'))
for k,v in rows.items(): print('%-58s %9.3f %9.3f %9.3f'%(k,v.get('default',0),v.get('st_dbuf0',0),v.get('st_dbufall',0)))
PY
timeout 300 python tools/gemm_trace.py --lib tools/variants/libstar_trace.so 320 960 0 > gpurun_out/j11_trace_qkv.log 2>&1
grep -A24 "## epilogue" gpurun_out/j11_trace_qkv.log | head -30; grep "clk per tile" gpurun_out/j11_trace_qkv.log
