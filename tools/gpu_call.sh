# GPU-box scratch runner used during development (gpurun -- 'bash tools/gpu_call.sh'): edit per experiment.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_msm.py -x -q -k "bn254 or linearity or pipelined or reference" 2>&1 | tail -8 > gpurun_out/t_sort.log; cat gpurun_out/t_sort.log
(for lg in 20 22 24 26; do python tools/stage_profile.py $lg 0; B200_MSM_SORT=1 python tools/stage_profile.py $lg 0; done) 2>&1 | grep "^msm" | tee gpurun_out/sort_stage.txt
