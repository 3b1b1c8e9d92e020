# GPU-box scratch runner used during development (gpurun -- 'bash tools/gpu_call.sh'): edit per experiment.
mkdir -p gpurun_out
./build/host_copy_probe > gpurun_out/host_copy_probe.txt 2>&1; cat gpurun_out/host_copy_probe.txt
timeout 1200 python -m pytest tests/test_gpu_golden.py tests/test_gpu_multi.py tests/test_gpu_ntt.py tests/test_gpu_vec_ops.py -x -q 2>&1 | tail -15 > gpurun_out/t_new.log; cat gpurun_out/t_new.log
