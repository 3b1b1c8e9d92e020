mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -6 > gpurun_out/t_full.log; cat gpurun_out/t_full.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 600 gpurun_out/bench_final.json; tail -2 gpurun_out/bench_final.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active
timeout 300 ncu --metrics $M --clock-control none --csv -k regex:k_ntt --log-file gpurun_out/ncu_ntt_bn254_2p24.csv python tools/one_ntt.py bn254 24 1 > gpurun_out/ncu_n1.log 2>&1
timeout 300 ncu --metrics $M --clock-control none --csv -k regex:k_ntt --log-file gpurun_out/ncu_ntt_babybear_2p27.csv python tools/one_ntt.py babybear 27 2 > gpurun_out/ncu_n2.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_ntt31 --launch-skip 3 -c 1 -o gpurun_out/r1_k_ntt31_babybear_2p27 -f python tools/one_ntt.py babybear 27 2 > gpurun_out/ncu_n3.log 2>&1; tail -1 gpurun_out/ncu_n3.log
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_ntt_tile --launch-skip 3 -c 1 -o gpurun_out/r1_k_ntt_tile_bn254_2p24 -f python tools/one_ntt.py bn254 24 1 > gpurun_out/ncu_n4.log 2>&1; tail -1 gpurun_out/ncu_n4.log
