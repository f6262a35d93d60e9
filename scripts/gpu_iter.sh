# scratch script of the current experiment: rewritten before every `gpurun -- 'bash scripts/gpu_iter.sh <tag>'`
set -x
mkdir -p gpurun_out
T=${1:-it}
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.txt 2>&1; tail -1 gpurun_out/${T}_smoke.txt
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/${T}_pytest_gpu.txt 2>&1; tail -3 gpurun_out/${T}_pytest_gpu.txt
