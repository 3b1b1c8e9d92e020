# GPU-box scratch runner used during development (gpurun -- 'bash tools/gpu_call.sh'): edit per experiment.
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active
timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2_launches_msm_2p26.csv python tools/one_msm.py 26 1 > gpurun_out/ncu1.log 2>&1
timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2_launches_ntt_bn254_2p24.csv python tools/one_ntt.py bn254 24 1 > gpurun_out/ncu2.log 2>&1
timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2_launches_ntt_babybear_2p27.csv python tools/one_ntt.py babybear 27 2 > gpurun_out/ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k_ntt31 -s 3 -c 1 -o /tmp/r2_full_ntt31 python tools/one_ntt.py babybear 27 2 > gpurun_out/ncu4.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k_ntt_tile -s 3 -c 1 -o /tmp/r2_full_ntt_tile python tools/one_ntt.py bn254 24 1 > gpurun_out/ncu5.log 2>&1
for f in r2_full_ntt31 r2_full_ntt_tile; do ncu -i /tmp/$f.ncu-rep --page raw --csv > gpurun_out/${f}_raw.csv 2>/dev/null; done
timeout 900 compute-sanitizer --tool racecheck --print-limit 10 python tools/race_probe.py > gpurun_out/r2_racecheck.txt 2>&1
timeout 900 compute-sanitizer --tool memcheck --print-limit 10 python tools/race_probe.py > gpurun_out/r2_memcheck.txt 2>&1
du -sh gpurun_out
