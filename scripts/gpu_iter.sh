# scratch script of the current experiment: rewritten before every `gpurun -- 'bash scripts/gpu_iter.sh <tag>'`
set -x
mkdir -p gpurun_out
T=${1:-it}
python scripts/tune_legendre.py c3 13 3 > gpurun_out/${T}_tune.txt 2>&1; grep "variant \(0\|7\|9\|10\|11\|12\):" gpurun_out/${T}_tune.txt
