mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_golden.py tests/test_gpu_vec_ops.py -x -q 2>&1 | tail -5 > gpurun_out/t_m31.log; cat gpurun_out/t_m31.log
timeout 600 python tools/widen_probe.py > gpurun_out/widen_probe.txt 2>&1; cat gpurun_out/widen_probe.txt
