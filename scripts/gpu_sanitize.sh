# compute-sanitizer passes over the small GPU tests of the newer kernels (memcheck: out-of-bounds / misaligned accesses;
# racecheck: shared-memory hazards, including the thread-block-cluster FFT path).  Slow: keep to the small cases.
set -x
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_general_gpu.py tests/test_healpix_gpu.py -x -q -m gpu -k "not at_scale" > gpurun_out/sanitize_memcheck.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/sanitize_memcheck.txt; tail -6 gpurun_out/sanitize_memcheck.txt
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_fft_gpu.py -x -q -m gpu -k "queb or window or c2c_2d or r2c_c2r_1d or strided_views" > gpurun_out/sanitize_memcheck_fft.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/sanitize_memcheck_fft.txt; tail -6 gpurun_out/sanitize_memcheck_fft.txt
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_fft_gpu.py -x -q -m gpu -k "cluster_split or r2c_c2r_2d" > gpurun_out/sanitize_racecheck_fft.txt 2>&1; echo "racecheck rc=$?" >> gpurun_out/sanitize_racecheck_fft.txt; tail -6 gpurun_out/sanitize_racecheck_fft.txt
