set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; tail -4 gpurun_out/pytest_gpu.txt | cut -c1-300
python scripts/tune_legendre.py c3 1 0123 > gpurun_out/tune_c3_h.txt 2>&1; cat gpurun_out/tune_c3_h.txt
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; cut -c1-400 gpurun_out/bench_quick.json; tail -3 gpurun_out/bench_quick.err
