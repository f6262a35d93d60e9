set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; tail -4 gpurun_out/pytest_gpu.txt | cut -c1-300
python scripts/tune_legendre.py c3 10 13 > gpurun_out/tune_c3_f.txt 2>&1
grep -v unknown gpurun_out/tune_c3_f.txt
