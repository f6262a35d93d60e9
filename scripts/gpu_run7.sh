set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mc_gpu.py -m gpu -x -q > gpurun_out/pytest_mc.txt 2>&1; tail -15 gpurun_out/pytest_mc.txt
python scripts/bench_fft.py 4096 8192 3 3 > gpurun_out/fft_c5_small.json 2>&1; cat gpurun_out/fft_c5_small.json
python scripts/bench_fft.py 16384 32768 3 2 > gpurun_out/fft_c5.json 2>&1; cat gpurun_out/fft_c5.json
ncu --set full --clock-control none --import-source on -k regex:k_fft_axis -s 2 -c 2 -o gpurun_out/r1f_fft_c5 python scripts/bench_fft.py 16384 32768 1 1 > gpurun_out/ncu_fft.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_resamp -s 10 -c 5 -o gpurun_out/r1f_resamp python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_resamp.log 2>&1
ls -la gpurun_out | tail -8
