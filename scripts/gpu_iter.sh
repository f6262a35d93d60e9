set -x
mkdir -p gpurun_out
T=${1:-it}
python -m pytest tests/test_curvedsky_gpu.py -x -q -m gpu -k "one_plan or grouped_host or streamed_host" 2>&1 | tail -3
for i in 1 2 3; do timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_curvedsky_gpu.py -x -q -m gpu -k "grouped_host or one_plan" 2>&1 | grep "passed\|failed\|ERROR SUMMARY" | head -4; done
timeout 1500 python -m pytest tests -q -x -m gpu 2>&1 | tail -2
