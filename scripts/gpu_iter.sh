set -x
mkdir -p gpurun_out
T=${1:-it}
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.txt 2>&1; tail -2 gpurun_out/${T}_smoke.txt
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/${T}_pytest_gpu.txt 2>&1; tail -3 gpurun_out/${T}_pytest_gpu.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_c3_1gpu.json 2> gpurun_out/${T}_bench.err; tail -2 gpurun_out/${T}_bench.err; cut -c1-300 gpurun_out/${T}_bench_c3_1gpu.json
