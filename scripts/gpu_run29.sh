set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_general_gpu.py -x -q -m gpu > gpurun_out/pytest_general.txt 2>&1; tail -30 gpurun_out/pytest_general.txt
