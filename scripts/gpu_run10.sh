set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.txt 2>&1; tail -25 gpurun_out/pytest_gpu.txt | cut -c1-300
python scripts/bench_mc.py --nsim 16 > gpurun_out/mc_c4_1gpu.json 2> gpurun_out/mc_c4.err; cat gpurun_out/mc_c4_1gpu.json; tail -3 gpurun_out/mc_c4.err
