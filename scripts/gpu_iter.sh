# scratch script of the current experiment: rewritten before every `gpurun -- 'bash scripts/gpu_iter.sh <tag>'`
set -x
mkdir -p gpurun_out
T=${1:-it}
python scripts/e2e_probe.py > gpurun_out/${T}_probe.json 2>/dev/null; cat gpurun_out/${T}_probe.json
