# scratch script of the current experiment: rewritten before every `gpurun -- 'bash scripts/gpu_iter.sh <tag>'`
set -x
mkdir -p gpurun_out
T=${1:-it}
python -m pytest tests/test_fft_gpu.py -x -q -m gpu -k "fourier_filter or smooth or window or map2harm" 2>&1 | tail -2
python bench.py --steps 3 --warmup 3 --no-cpu --no-parity > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -2 gpurun_out/${T}_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/'+__import__('sys').argv[1]+'_bench.json')) if False else json.load(open([f for f in __import__('glob').glob('gpurun_out/*_bench.json')][-1]))
print(d['value'], d['e2e']['value']); print(json.dumps(d['configs']['c5']))
PY
