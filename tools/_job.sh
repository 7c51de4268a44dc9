mkdir -p gpurun_out
(timeout 900 python -m pytest -m gpu -x -q -s tests/test_cogvideox_vae.py tests/test_cogvideox_sampler.py tests/test_text_tower.py tests/test_cogvideox.py > gpurun_out/k5_tests_cogvideox_path.log 2>&1; echo "EXIT $?" >> gpurun_out/k5_tests_cogvideox_path.log)
grep -n "cogvideox\|passed\|failed\|Error\|EXIT\|text tower" gpurun_out/k5_tests_cogvideox_path.log | tail -20
timeout 900 python bench.py --workload cogvideox --steps 3 --warmup 1 > gpurun_out/k5_bench_cogvideox.json 2> gpurun_out/k5_bench_cogvideox.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/k5_bench_cogvideox.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e'], json.dumps(d['pipeline'], indent=1))
PY
tail -5 gpurun_out/k5_bench_cogvideox.err
