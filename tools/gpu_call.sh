# GPU-box scratch runner used during development (gpurun -- 'bash tools/gpu_call.sh'): edit per experiment.  Default = the round check.
mkdir -p gpurun_out
(time timeout 2400 python -m pytest tests/ -x -q -m gpu --durations=8 2>&1 | tail -16 > gpurun_out/t_full.log) 2> gpurun_out/t_full.time; cat gpurun_out/t_full.log; tail -3 gpurun_out/t_full.time
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
(time timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err) 2> gpurun_out/bench.time; tail -c 1500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err; tail -3 gpurun_out/bench.time
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; cat gpurun_out/bench_ref.json | cut -c1-400
