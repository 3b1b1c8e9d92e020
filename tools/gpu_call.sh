mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -12 > gpurun_out/t_full.log; cat gpurun_out/t_full.log
timeout 400 python tools/e2e_probe.py 26 0,4,2 > gpurun_out/e2e_probe2.txt 2>&1; cat gpurun_out/e2e_probe2.txt
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
