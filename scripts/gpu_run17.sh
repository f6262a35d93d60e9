set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_fft_axis -s 2 -c 2 -o /tmp/r1i_fft python scripts/bench_fft.py 16384 32768 1 1 > gpurun_out/ncu_fft.log 2>&1
python scripts/ncu_summary.py full /tmp/r1i_fft.ncu-rep > gpurun_out/r1i_fft_full.txt 2>&1
python scripts/ncu_source.py /tmp/r1i_fft.ncu-rep 10 > gpurun_out/r1i_fft_stalls.txt 2>&1
