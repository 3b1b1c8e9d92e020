mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 300 gpurun_out/bench_final.json; tail -2 gpurun_out/bench_final.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active
timeout 300 ncu --metrics $M --clock-control none --csv -k regex:k_ntt --log-file gpurun_out/ncu_ntt_bn254_2p24.csv python tools/one_ntt.py bn254 24 1 > gpurun_out/ncu_n1.log 2>&1
timeout 300 ncu --metrics $M --clock-control none --csv -k regex:k_ntt --log-file gpurun_out/ncu_ntt_babybear_2p27.csv python tools/one_ntt.py babybear 27 2 > gpurun_out/ncu_n2.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:k_ntt31 --launch-skip 3 -c 1 -o /tmp/r1_k_ntt31 -f python tools/one_ntt.py babybear 27 2 > gpurun_out/ncu_n3.log 2>&1; ncu -i /tmp/r1_k_ntt31.ncu-rep --page raw --csv > gpurun_out/r1_ncu_full_k_ntt31_babybear_2p27_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none -k regex:k_ntt_tile --launch-skip 3 -c 1 -o /tmp/r1_k_ntt_tile -f python tools/one_ntt.py bn254 24 1 > gpurun_out/ncu_n4.log 2>&1; ncu -i /tmp/r1_k_ntt_tile.ncu-rep --page raw --csv > gpurun_out/r1_ncu_full_k_ntt_tile_bn254_2p24_raw.csv 2>/dev/null
ls -la gpurun_out; du -sh gpurun_out
