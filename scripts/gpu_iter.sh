set -x
python -m pytest tests/test_fft_gpu.py -x -q -m gpu -k "shift_and_resample or engine_object or dct" 2>&1 | tail -3
