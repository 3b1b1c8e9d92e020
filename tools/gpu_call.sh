# GPU-box scratch runner used during development (gpurun -- 'bash tools/gpu_call.sh'): edit per experiment.
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/gpus8.txt; cat /sys/fs/cgroup/cpu.max >> gpurun_out/gpus8.txt; cat gpurun_out/gpus8.txt
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err) 2> gpurun_out/bench_n8.time; tail -c 5000 gpurun_out/bench_n8.json; tail -4 gpurun_out/bench_n8.err; tail -3 gpurun_out/bench_n8.time
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -4
