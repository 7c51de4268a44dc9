mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q -s -k "cogvideox_vae" > gpurun_out/k3_tests_vae3d.log 2>&1; echo "EXIT $?" >> gpurun_out/k3_tests_vae3d.log)
tail -25 gpurun_out/k3_tests_vae3d.log
(timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/k3_gpu_tests.log 2>&1; echo "EXIT $?" >> gpurun_out/k3_gpu_tests.log)
tail -4 gpurun_out/k3_gpu_tests.log
timeout 600 python bench.py --workload cogvideox --steps 3 --warmup 1 > gpurun_out/k3_bench_cogvideox.json 2> gpurun_out/k3_bench_cogvideox.err; tail -c 1500 gpurun_out/k3_bench_cogvideox.json; tail -5 gpurun_out/k3_bench_cogvideox.err
