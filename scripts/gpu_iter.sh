set -x
mkdir -p gpurun_out
T=${1:-it}
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/${T}_pytest_gpu.txt 2>&1; tail -3 gpurun_out/${T}_pytest_gpu.txt
python scripts/e2e_probe.py 2>/dev/null
