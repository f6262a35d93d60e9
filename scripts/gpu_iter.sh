set -x
mkdir -p gpurun_out
T=${1:-it}
timeout 1500 python -m pytest tests -q -x -m gpu > gpurun_out/${T}_pytest_gpu.txt 2>&1; tail -5 gpurun_out/${T}_pytest_gpu.txt
python scripts/tune_legendre.py c3 1 0123 > gpurun_out/${T}_tune.txt 2>&1; cat gpurun_out/${T}_tune.txt
python scripts/e2e_probe.py > gpurun_out/${T}_probe.json 2> gpurun_out/${T}_probe.err; cat gpurun_out/${T}_probe.json
