mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --workload cogvideox --gpus 2 --steps 3 --warmup 1 > gpurun_out/k8_bench_cogvideox_n2.json 2> gpurun_out/k8_bench_cogvideox_n2.err
tail -c 1200 gpurun_out/k8_bench_cogvideox_n2.json; tail -5 gpurun_out/k8_bench_cogvideox_n2.err
