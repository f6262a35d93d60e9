set -x
mkdir -p gpurun_out
python scripts/tune_legendre.py c3 21 3 > gpurun_out/tune_c3_g.txt 2>&1
grep -v unknown gpurun_out/tune_c3_g.txt
