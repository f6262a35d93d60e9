set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; tail -3 gpurun_out/pytest_gpu.txt
B2_LEG_VARIANT=4,10,5,12 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_sys.txt 2>&1; tail -15 gpurun_out/pytest_gpu_sys.txt
python scripts/tune_legendre.py c3 16 13 > gpurun_out/tune_c3_e.txt 2>&1
grep -v unknown gpurun_out/tune_c3_e.txt
B2_LEG_VARIANT=4,10,5,12 ncu --set full --clock-control none --import-source on -k regex:k_adj2 -s 1 -c 1 -o gpurun_out/r1e_adj2_sys12 python scripts/tune_legendre.py c3 0 3 > gpurun_out/ncu_adj2.log 2>&1
python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; cat gpurun_out/bench_quick.json; tail -3 gpurun_out/bench_quick.err
