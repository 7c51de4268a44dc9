#!/bin/bash
# round-2 job 7: double-buffered output staging in tapgemm2 -- A/B by reduction length, parity on the variant
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
for v in default dbuf700 dbuf1300 dbufall; do
  if [ $v = default ]; then LIB=""; else LIB="--lib tools/variants/libstar_$v.so"; fi
  echo "=== $v" >> gpurun_out/j7_dbuf_ab.log
  timeout 300 python tools/kbench.py linear conv $LIB >> gpurun_out/j7_dbuf_ab.log 2>&1
done
python - <<'PY'
import re,collections
rows=collections.OrderedDict(); cur=None
for line in open('gpurun_out/j7_dbuf_ab.log'):
    if line.startswith('==='): cur=line.split()[1]; continue
    m=re.match(r'(.{58})\s+([0-9.]+) ms',line)
    if m: rows.setdefault(m.group(1).strip(),{})[cur]=float(m.group(2))
print('%-58s %9s %9s %9s %9s'%('shape','default','dbuf700','dbuf1300','dbufall'))
for k,v in rows.items(): print('%-58s %9.3f %9.3f %9.3f %9.3f'%(k,v.get('default',0),v.get('dbuf700',0),v.get('dbuf1300',0),v.get('dbufall',0)))
PY
STAR_LIB_VARIANT=tools/variants/libstar_dbufall.so timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_config2_gpu.py tests/test_unet_gpu.py -q -x -p no:cacheprovider -k "linear or conv or unet" > gpurun_out/j7_tests_dbufall.log 2>&1
echo "tests dbufall rc=$?"; tail -3 gpurun_out/j7_tests_dbufall.log
