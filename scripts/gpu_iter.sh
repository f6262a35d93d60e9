set -x
mkdir -p gpurun_out
T=${1:-it}
B2_TRACE=1 python scripts/e2e_probe.py > gpurun_out/${T}_probe.json 2> gpurun_out/${T}_trace.txt; cat gpurun_out/${T}_probe.json; grep -n "map -> alm" -A20 gpurun_out/${T}_trace.txt | tail -21 | grep "K5 done\|K2 done"
B2_LEG_VARIANT=0,9,5,8 python scripts/e2e_probe.py 2>/dev/null
python -m pytest tests/test_sht_gpu.py tests/test_curvedsky_gpu.py tests/test_baseline_parity_gpu.py -x -q -m gpu 2>&1 | tail -2
