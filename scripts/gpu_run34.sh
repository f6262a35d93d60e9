set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_uharm_gpu.py -x -q -m gpu > gpurun_out/pytest_uharm.txt 2>&1; tail -40 gpurun_out/pytest_uharm.txt
