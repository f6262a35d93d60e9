# Side-path evidence: FFT bench with the scipy CPU timing beside it, and an ncu --set full summary of the K8 kernels.
set -x
mkdir -p gpurun_out
python scripts/bench_fft.py 4096 8192 3 3 > gpurun_out/fft_c5_small.json 2>&1; cat gpurun_out/fft_c5_small.json
python scripts/bench_fft.py 16384 32768 3 2 > gpurun_out/fft_c5.json 2>&1; cat gpurun_out/fft_c5.json
ncu --set full --clock-control none --import-source on -k regex:"k_gen_" -c 6 -o /tmp/gen_full python scripts/bench_general.py 4096 4608 9216 0 > gpurun_out/bench_under_ncu3.log 2>&1
python scripts/ncu_summary.py full /tmp/gen_full.ncu-rep > gpurun_out/gen_full.txt 2>&1
grep -n "====\|gpu__time_duration\|dram__bytes\|sm__throughput\|warps_active\|registers" gpurun_out/gen_full.txt | head -60
