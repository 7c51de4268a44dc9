mkdir -p gpurun_out
for v in pipe1 pipe2; do echo "=== $v"; STAR_LIB_VARIANT=tools/variants/libstar_$v.so timeout 300 python -m pytest tests -m gpu -q -k "attention or unet_full or cfg_pair" 2>&1 | tail -4; done > gpurun_out/k2_tests_pipe.log 2>&1
(echo "=== default"; timeout 200 python tools/kbench.py attention; for v in pipe1 pipe2; do echo "=== $v"; timeout 200 python tools/kbench.py --lib tools/variants/libstar_$v.so attention; done) > gpurun_out/k2_kbench_attn_pipe.log 2>&1
for v in atrace atrace_pipe2; do echo "=== $v"; timeout 120 python tools/attn_trace.py --lib tools/variants/libstar_$v.so; done > gpurun_out/k2_attn_trace.log 2>&1
cat gpurun_out/k2_tests_pipe.log; grep "attention self\|===" gpurun_out/k2_kbench_attn_pipe.log; grep -v "^  [a-zA-Z_]* *[0-9]* *(\|^  [a-zA-Z_]* *[0-9]*$" gpurun_out/k2_attn_trace.log
