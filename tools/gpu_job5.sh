#!/bin/bash
# round-2 job 5: short-K GEMM attribution (store path / epilogue math removed), config-5 VAE decode sweep
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
for v in default gexp1 gexp2; do
  if [ $v = default ]; then LIB=""; else LIB="--lib tools/variants/libstar_$v.so"; fi
  echo "=== $v" >> gpurun_out/j5_gemm_attr.log
  timeout 300 python tools/kbench.py linear conv $LIB >> gpurun_out/j5_gemm_attr.log 2>&1
done
grep -E "===|linear L0|conv2d 3x3 320" gpurun_out/j5_gemm_attr.log
timeout 900 python bench.py --workload vae --steps 3 --warmup 3 > gpurun_out/j5_vae_sweep.json 2> gpurun_out/j5_vae_sweep.err
echo "vae rc=$?"; cut -c1-2500 gpurun_out/j5_vae_sweep.json; tail -3 gpurun_out/j5_vae_sweep.err
