# GPU-box scratch runner used during development (gpurun -- 'bash tools/gpu_call.sh'): edit per experiment.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_dropin.py -x -q 2>&1 | tail -12 > gpurun_out/t_dropin.log; cat gpurun_out/t_dropin.log
