set -x
mkdir -p gpurun_out
timeout 600 python scripts/bench_healpix.py 512 1024 2 > gpurun_out/healpix_512.json 2>&1; cat gpurun_out/healpix_512.json
timeout 900 python scripts/bench_healpix.py 2048 4096 2 > gpurun_out/healpix_2048.json 2>&1; cat gpurun_out/healpix_2048.json
