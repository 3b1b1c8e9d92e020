# GPU-box scratch runner used during development (gpurun -- 'bash tools/gpu_call.sh'): edit per experiment.
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err) 2> gpurun_out/bench_n$N.time
python -c "
import json; d=json.load(open('gpurun_out/bench_n$N.json')); print($N, d['value'], d['ms_per_step'], {k: round(v['ms_per_step'],1) for k,v in d['e2e']['variants'].items()}, d['e2e']['path'])
for c in d['configs']: print(c['config'][:70], c.get('value'), c.get('ms_per_pass'), c.get('parity',{}).get('ok'), c.get('all_to_all_ms'), c.get('nvlink_gbs_per_gpu_each_way'), c.get('error'))"
tail -3 gpurun_out/bench_n$N.err; tail -3 gpurun_out/bench_n$N.time
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | cut -c1-200
