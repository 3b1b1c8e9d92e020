# GPU-box scratch runner used during development (gpurun -- 'bash tools/gpu_call.sh'): edit per experiment.
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_multi.py "tests/test_gpu_fullsize.py::test_bn254_msm_vs_reference" -x -q 2>&1 | tail -4 > gpurun_out/t_last.log; cat gpurun_out/t_last.log; echo "exit=$?"
