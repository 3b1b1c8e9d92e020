# GPU-box scratch runner used during development (gpurun -- 'bash tools/gpu_call.sh'): edit per experiment.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ntt.py tests/test_gpu_golden.py -x -q 2>&1 | tail -4 > gpurun_out/t_ntt.log; cat gpurun_out/t_ntt.log
timeout 600 python tools/ntt_probe2.py 2>&1 | grep "G elem" | tee gpurun_out/ntt_probe2b.txt
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active
timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2_launches_msm_2p26.csv python tools/one_msm.py 26 1 > gpurun_out/ncu1.log 2>&1
timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2_launches_ntt_bn254_2p24.csv python tools/one_ntt.py bn254 24 1 > gpurun_out/ncu2.log 2>&1
timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2_launches_ntt_babybear_2p27.csv python tools/one_ntt.py babybear 27 2 > gpurun_out/ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ntt31 -s 3 -c 1 -o gpurun_out/r2_full_ntt31 python tools/one_ntt.py babybear 27 2 > gpurun_out/ncu4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ntt_tile -s 3 -c 1 -o gpurun_out/r2_full_ntt_tile python tools/one_ntt.py bn254 24 1 > gpurun_out/ncu5.log 2>&1
for f in r2_full_ntt31 r2_full_ntt_tile; do ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/${f}_raw.csv 2>/dev/null; done
timeout 900 compute-sanitizer --tool racecheck --print-limit 10 python tools/race_probe.py > gpurun_out/r2_racecheck.txt 2>&1; tail -5 gpurun_out/r2_racecheck.txt
ls -la gpurun_out | head -40
