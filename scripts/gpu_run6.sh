set -x
mkdir -p gpurun_out
./scratch/ubench_dfma > gpurun_out/ubench_dfma.txt 2>&1; cat gpurun_out/ubench_dfma.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; tail -25 gpurun_out/pytest_gpu.txt
B2_LEG_VARIANT=4,10,5,12 timeout 900 python -m pytest tests/test_sht_gpu.py tests/test_curvedsky_gpu.py -m gpu -x -q > gpurun_out/pytest_gpu_sys.txt 2>&1; tail -5 gpurun_out/pytest_gpu_sys.txt
python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; cat gpurun_out/bench_quick.json; tail -3 gpurun_out/bench_quick.err
