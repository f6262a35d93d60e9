# scratch script of the current experiment: rewritten before every `gpurun -- 'bash scripts/gpu_iter.sh <tag>'`
set -x
mkdir -p gpurun_out
T=${1:-it}
python -m pytest tests/test_healpix_gpu.py tests/test_sht_gpu.py tests/test_general_gpu.py -x -q -m gpu 2>&1 | tail -2
python scripts/bench_healpix.py 2048 4096 2 2>/dev/null | tee gpurun_out/${T}_healpix_pack.json
B2_NO_PACK=1 python scripts/bench_healpix.py 2048 4096 2 2>/dev/null | tee gpurun_out/${T}_healpix_nopack.json
python scripts/bench_healpix.py 512 1024 2 2>/dev/null
python scripts/e2e_probe.py 2>/dev/null | cut -c1-120
