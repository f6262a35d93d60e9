# scratch script of the current experiment: rewritten before every `gpurun -- 'bash scripts/gpu_iter.sh <tag>'`
set -x
mkdir -p gpurun_out
T=${1:-it}
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_healpix_gpu.py -x -q -m gpu > gpurun_out/${T}_sanitize_memcheck_healpix_pack.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/${T}_sanitize_memcheck_healpix_pack.txt; tail -4 gpurun_out/${T}_sanitize_memcheck_healpix_pack.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_mc_gpu.py -x -q -m gpu -k "batched and not 2304" > gpurun_out/${T}_sanitize_memcheck_batched.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/${T}_sanitize_memcheck_batched.txt; tail -4 gpurun_out/${T}_sanitize_memcheck_batched.txt
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_mc_gpu.py -x -q -m gpu -k "batched and 256" > gpurun_out/${T}_sanitize_racecheck_batched.txt 2>&1; echo "racecheck rc=$?" >> gpurun_out/${T}_sanitize_racecheck_batched.txt; tail -4 gpurun_out/${T}_sanitize_racecheck_batched.txt
