mkdir -p gpurun_out
(timeout 600 python -m pytest -m gpu -x -q tests/test_unet_gpu.py -k "graph or cfg_pair" > gpurun_out/k9_tests_graph.log 2>&1; echo "EXIT $?" >> gpurun_out/k9_tests_graph.log); tail -15 gpurun_out/k9_tests_graph.log
timeout 600 python tools/graph_bench.py > gpurun_out/k9_graph_bench.log 2>&1; tail -12 gpurun_out/k9_graph_bench.log
