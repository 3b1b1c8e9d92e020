# GPU-box check used during development: full -m gpu suite, then the default bench line (outputs under gpurun_out/).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -8 > gpurun_out/t_full.log; cat gpurun_out/t_full.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_pairs.json 2> gpurun_out/bench_pairs.err; tail -c 3000 gpurun_out/bench_pairs.json; tail -3 gpurun_out/bench_pairs.err
