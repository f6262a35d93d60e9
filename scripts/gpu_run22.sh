set -x
mkdir -p gpurun_out
for w in c0 c1 c2 c3; do python scripts/tune_legendre.py $w 5 01; done > gpurun_out/tune_even.txt 2>&1
cat gpurun_out/tune_even.txt
