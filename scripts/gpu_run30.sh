set -x
mkdir -p gpurun_out
timeout 600 python scripts/bench_general.py 2048 2304 4608 2 > gpurun_out/general_2048.json 2>&1; cat gpurun_out/general_2048.json
timeout 900 python scripts/bench_general.py 4096 4608 9216 2 > gpurun_out/general_4096.json 2>&1; cat gpurun_out/general_4096.json
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.txt 2>&1; tail -5 gpurun_out/pytest_gpu.txt
