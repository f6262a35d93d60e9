set -x
mkdir -p gpurun_out
T=${1:-it}
python -m pytest tests/test_fft_gpu.py -x -q -m gpu -k "dct" 2>&1 | tail -15
