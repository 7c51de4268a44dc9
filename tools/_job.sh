mkdir -p gpurun_out
timeout 900 python bench.py --workload cogvideox --steps 3 --warmup 1 > gpurun_out/k6_bench_cogvideox.json 2> gpurun_out/k6_bench_cogvideox.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/k6_bench_cogvideox.json').read().strip().splitlines()[-1])
    print(d['ms_per_step'], d['value'], d['e2e'], json.dumps(d['pipeline'], indent=1))
except Exception as e:
    print("bench parse failed", e)
PY
tail -5 gpurun_out/k6_bench_cogvideox.err
timeout 300 python tools/vae3d_kbench.py > gpurun_out/k6_vae3d_kbench.log 2>&1; cat gpurun_out/k6_vae3d_kbench.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tapgemm2 --launch-skip 2 -c 1 -o gpurun_out/k6_conv3d_l0 python tools/vae3d_kbench.py > gpurun_out/k6_ncu_conv3d.log 2>&1; tail -3 gpurun_out/k6_ncu_conv3d.log; ls -la gpurun_out/*.ncu-rep
