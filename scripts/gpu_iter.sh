set -x
python scripts/tune_legendre.py c3 0 0123 2>&1 | tail -4
python scripts/e2e_probe.py 2>/dev/null
python -m pytest tests/test_sht_gpu.py tests/test_curvedsky_gpu.py tests/test_baseline_parity_gpu.py tests/test_mc_gpu.py -x -q -m gpu 2>&1 | tail -2
