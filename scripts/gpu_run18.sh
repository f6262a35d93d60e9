set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; tail -4 gpurun_out/pytest_gpu.txt | cut -c1-300
python scripts/bench_fft.py 4096 8192 3 3 > gpurun_out/fft_c5_small.json 2>&1; cat gpurun_out/fft_c5_small.json
python scripts/bench_fft.py 16384 32768 3 2 > gpurun_out/fft_c5.json 2>&1; cat gpurun_out/fft_c5.json
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; python -c "
import json; d=json.load(open('gpurun_out/bench_quick.json')); print(d['value'], d['ms_per_step'], d['roundtrip_rel_err'], d['e2e']['value'], d['stage_ms_last_map2alm_group'])"
