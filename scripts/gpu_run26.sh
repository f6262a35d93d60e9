set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_fft_cl -s 4 -c 2 -o /tmp/fftcl python scripts/bench_fft.py 4096 8192 1 1 > gpurun_out/ncu_fftcl.log 2>&1
python scripts/ncu_summary.py full /tmp/fftcl.ncu-rep > gpurun_out/r1n_fftcl_full.txt 2>&1
python scripts/ncu_source.py /tmp/fftcl.ncu-rep 10 > gpurun_out/r1n_fftcl_stalls.txt 2>&1
