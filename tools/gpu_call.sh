# GPU-box scratch runner used during development (gpurun -- 'bash tools/gpu_call.sh'): edit per experiment.
mkdir -p gpurun_out
for NB in 0 1; do
if [ $NB = 1 ]; then export B200_BENCH_NO_NUMA_BIND=1; fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 3 --warmup 3 --no-configs 2>gpurun_out/n4_nb$NB.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=4 no_bind=$NB', round(d['ms_per_step'],1), {k: round(v['ms_per_step'],1) for k,v in d['e2e']['variants'].items()}, d['config'].get('cpu_affinity'))"
done 2>&1 | tee gpurun_out/e2e_n4_numa.txt
