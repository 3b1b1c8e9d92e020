mkdir -p gpurun_out
./build/gather_bench > gpurun_out/gather_bench.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_msm.py tests/test_gpu_golden.py tests/test_gpu_ecntt.py tests/test_gpu_dropin.py -x -q 2>&1 | tail -12 > gpurun_out/t_msm.log; cat gpurun_out/t_msm.log
timeout 400 python tools/e2e_probe.py 26 4,8,16 > gpurun_out/e2e_probe.txt 2>&1; cat gpurun_out/e2e_probe.txt
cat gpurun_out/gather_bench.txt
