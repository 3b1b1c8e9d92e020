# GPU-box scratch runner used during development (gpurun -- 'bash tools/gpu_call.sh'): edit per experiment.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 1200 python -m pytest tests/test_gpu_multi.py tests/test_gpu_dropin.py tests/test_gpu_msm.py tests/test_gpu_vec_ops.py -x -q 2>&1 | tail -15 > gpurun_out/t_multi.log; cat gpurun_out/t_multi.log
for T in 8 12 16; do B200_COPIER_THREADS=$T timeout 300 python bench.py --steps 3 --warmup 3 --no-configs --no-ntt --no-cpu-baseline 2>gpurun_out/e2e_$T.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('T=$T', d['ms_per_step'], json.dumps(d['e2e']['variants']))"; done 2>&1 | tee gpurun_out/e2e_threads.txt
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err) 2> gpurun_out/bench_n2.time; tail -c 7000 gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err; cat gpurun_out/bench_n2.time
