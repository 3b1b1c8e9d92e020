mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_msm.py tests/test_gpu_golden.py -x -q 2>&1 | tail -6 > gpurun_out/t_msm.log; cat gpurun_out/t_msm.log
timeout 600 python tools/stage_profile.py 26 20 > gpurun_out/stage26.txt 2>&1; cat gpurun_out/stage26.txt
timeout 600 python tools/stage_profile.py 24 20 0,3 > gpurun_out/stage24.txt 2>&1; cat gpurun_out/stage24.txt
timeout 600 python tools/stage_profile.py 20 0 > gpurun_out/stage20.txt 2>&1; cat gpurun_out/stage20.txt
