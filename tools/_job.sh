mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q -s tests/test_text_tower.py tests/test_cogvideox_vae.py > gpurun_out/k4_tests_text_tower.log 2>&1; echo "EXIT $?" >> gpurun_out/k4_tests_text_tower.log)
tail -12 gpurun_out/k4_tests_text_tower.log
timeout 600 python bench.py --workload cogvideox --steps 3 --warmup 1 > gpurun_out/k4_bench_cogvideox.json 2> gpurun_out/k4_bench_cogvideox.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/k4_bench_cogvideox.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], json.dumps(d['pipeline'], indent=1))
PY
tail -5 gpurun_out/k4_bench_cogvideox.err
