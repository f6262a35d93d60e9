set -x
mkdir -p gpurun_out
T=${1:-it}
python scripts/tune_legendre.py c3 10 02 > gpurun_out/${T}_tune.txt 2>&1; cat gpurun_out/${T}_tune.txt
