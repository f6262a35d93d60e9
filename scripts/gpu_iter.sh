set -x
mkdir -p gpurun_out
T=${1:-it}
python scripts/e2e_probe.py 2>/dev/null
python -m pytest tests/test_fft_gpu.py tests/test_sht_gpu.py tests/test_curvedsky_gpu.py tests/test_baseline_parity_gpu.py -x -q -m gpu 2>&1 | tail -2
