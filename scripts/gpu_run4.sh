set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; tail -3 gpurun_out/pytest_gpu.txt
python scripts/tune_legendre.py c3 10 13 > gpurun_out/tune_c3_d.txt 2>&1
cat gpurun_out/tune_c3_d.txt
