# Round-2 evidence run on the GPU box (one B200): smoke, the GPU test suite, both bench arms, the ncu launch list of the
# bench command, --set full captures of the Legendre kernels, the DFMA operand microbenchmark, the stage timeline of the
# host-memory pair, compute-sanitizer over the new host-memory paths (completion flags, arrival gates).
set -x
mkdir -p gpurun_out
T=${1:-r2s}
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.txt 2>&1; tail -2 gpurun_out/${T}_smoke.txt
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/${T}_pytest_gpu.txt 2>&1; tail -3 gpurun_out/${T}_pytest_gpu.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_c3_1gpu.json 2> gpurun_out/${T}_bench.err; tail -2 gpurun_out/${T}_bench.err; cut -c1-400 gpurun_out/${T}_bench_c3_1gpu.json
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${T}_bench_c3_reference_arm.json 2> gpurun_out/${T}_bench_ref.err; cut -c1-300 gpurun_out/${T}_bench_c3_reference_arm.json
B2_TRACE=1 python scripts/e2e_probe.py > gpurun_out/${T}_e2e_probe.json 2> gpurun_out/${T}_e2e_trace_full.txt; cat gpurun_out/${T}_e2e_probe.json
(grep -n "map -> alm" -A20 gpurun_out/${T}_e2e_trace_full.txt | tail -21; grep -n "alm -> map" -A36 gpurun_out/${T}_e2e_trace_full.txt | tail -37) > gpurun_out/${T}_e2e_stage_trace.txt
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench_dfma_operands scripts/ubench/ubench_dfma_operands.cu && /tmp/ubench_dfma_operands > gpurun_out/${T}_ubench_dfma_operands.txt
# launch list of the bench command (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file /tmp/launches_c3.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-configs > gpurun_out/${T}_bench_under_ncu.log 2>&1
python scripts/ncu_summary.py launches /tmp/launches_c3.csv > gpurun_out/${T}_launches_c3.txt 2>&1; head -14 gpurun_out/${T}_launches_c3.txt
# the Legendre kernels of one step
ncu --set full --clock-control none --import-source on -k regex:"k_adj|k_synth" -s 8 -c 4 -o /tmp/leg_full python bench.py --steps 1 --warmup 1 --no-cpu --no-configs > gpurun_out/${T}_bench_under_ncu2.log 2>&1
python scripts/ncu_summary.py full /tmp/leg_full.ncu-rep > gpurun_out/${T}_leg_full.txt 2>&1
# sanitizer: the host-memory paths (single-warp CTAs publishing completion flags to mapped memory, CTAs waiting on arrival flags)
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_curvedsky_gpu.py -x -q -m gpu -k "streamed_host or grouped_host" > gpurun_out/${T}_sanitize_memcheck_hostpath.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/${T}_sanitize_memcheck_hostpath.txt; tail -4 gpurun_out/${T}_sanitize_memcheck_hostpath.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_sht_gpu.py -x -q -m gpu -k "(test_synthesis_2d or test_adjoint_synthesis_2d) and (F1-32 or CC-258 or MW-31)" > gpurun_out/${T}_sanitize_racecheck_legendre.txt 2>&1; echo "racecheck rc=$?" >> gpurun_out/${T}_sanitize_racecheck_legendre.txt; tail -4 gpurun_out/${T}_sanitize_racecheck_legendre.txt
