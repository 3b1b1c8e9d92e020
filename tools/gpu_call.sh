mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 100 python -m pytest tests/test_gpu_msm.py -x -q -k "pipelined or golden or sizes" 2>&1 | tail -2
