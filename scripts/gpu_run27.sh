set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fft_gpu.py -x -q -m gpu > gpurun_out/pytest_fft.txt 2>&1; tail -15 gpurun_out/pytest_fft.txt
timeout 300 python scripts/bench_fft.py 4096 8192 3 3 > gpurun_out/fft_c5_small.json 2>&1; cat gpurun_out/fft_c5_small.json
timeout 300 python scripts/bench_fft.py 16384 32768 3 2 > gpurun_out/fft_c5.json 2>&1; cat gpurun_out/fft_c5.json
