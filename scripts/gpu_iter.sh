set -x
mkdir -p gpurun_out
T=${1:-it}
timeout 1500 python -m pytest tests -q -x -m gpu > gpurun_out/${T}_pytest_gpu.txt 2>&1; tail -5 gpurun_out/${T}_pytest_gpu.txt
B2_TRACE=1 python scripts/e2e_probe.py > gpurun_out/${T}_probe.json 2> gpurun_out/${T}_trace.txt; cat gpurun_out/${T}_probe.json; grep -n "map -> alm" -A20 gpurun_out/${T}_trace.txt | tail -21;  grep -n "alm -> map" -A35 gpurun_out/${T}_trace.txt | tail -36
