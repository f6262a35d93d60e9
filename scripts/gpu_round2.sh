# Round-2 evidence run on the GPU box: smoke, the GPU test suite, both bench arms, the ncu launch list of the bench command,
# --set full captures of the Legendre kernels and of the TMA FFT kernels, compute-sanitizer passes over the Legendre kernels
# and the TMA FFT path.  Reports too large for gpurun_out/ stay in /tmp on the box; text summaries come back.
set -x
mkdir -p gpurun_out
T=${1:-r2}
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.txt 2>&1; tail -2 gpurun_out/${T}_smoke.txt
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/${T}_pytest_gpu.txt 2>&1; tail -3 gpurun_out/${T}_pytest_gpu.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_c3_1gpu.json 2> gpurun_out/${T}_bench.err; tail -2 gpurun_out/${T}_bench.err; cut -c1-300 gpurun_out/${T}_bench_c3_1gpu.json
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${T}_bench_c3_reference_arm.json 2> gpurun_out/${T}_bench_ref.err; cut -c1-300 gpurun_out/${T}_bench_c3_reference_arm.json
# launch list of the bench command (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file /tmp/launches_c3.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-configs > gpurun_out/${T}_bench_under_ncu.log 2>&1
python scripts/ncu_summary.py launches /tmp/launches_c3.csv > gpurun_out/${T}_launches_c3.txt 2>&1; head -14 gpurun_out/${T}_launches_c3.txt
# the Legendre kernels of one step
ncu --set full --clock-control none --import-source on -k regex:"k_adj|k_synth" -s 8 -c 4 -o /tmp/leg_full python bench.py --steps 1 --warmup 1 --no-cpu --no-configs > gpurun_out/${T}_bench_under_ncu2.log 2>&1
python scripts/ncu_summary.py full /tmp/leg_full.ncu-rep > gpurun_out/${T}_leg_full.txt 2>&1
# the TMA FFT kernels at C5 size (one component): X1, X2, Y1 (mirrored, untangle), Y2 of rfft2 and the four of irfft2
ncu --set full --clock-control none --import-source on -k regex:k_tfft -c 8 -o /tmp/tfft_full python scripts/c5dbg.py 1 16384 32768 fi > gpurun_out/${T}_tfft_under_ncu.log 2>&1
python scripts/ncu_summary.py full /tmp/tfft_full.ncu-rep > gpurun_out/${T}_tfft_full.txt 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -k regex:k_tfft --csv --log-file gpurun_out/${T}_tfft_c5_launches.csv python scripts/c5dbg.py 3 16384 32768 fi > /dev/null 2>&1
# sanitizer: the warp-synchronous Legendre kernels (cp.async double buffer) and the mbarrier / TMA pipeline
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_sht_gpu.py -x -q -m gpu -k "test_synthesis_2d or test_adjoint_synthesis_2d or test_analysis_2d" > gpurun_out/${T}_sanitize_memcheck_legendre.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/${T}_sanitize_memcheck_legendre.txt; tail -4 gpurun_out/${T}_sanitize_memcheck_legendre.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_sht_gpu.py -x -q -m gpu -k "(test_synthesis_2d or test_adjoint_synthesis_2d) and (F1-32 or CC-258 or MW-31)" > gpurun_out/${T}_sanitize_racecheck_legendre.txt 2>&1; echo "racecheck rc=$?" >> gpurun_out/${T}_sanitize_racecheck_legendre.txt; tail -4 gpurun_out/${T}_sanitize_racecheck_legendre.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_fft_gpu.py -x -q -m gpu -k "c2c_2d or r2c_c2r_2d" > gpurun_out/${T}_sanitize_memcheck_tfft.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/${T}_sanitize_memcheck_tfft.txt; tail -4 gpurun_out/${T}_sanitize_memcheck_tfft.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/check_tfft.py > gpurun_out/${T}_sanitize_racecheck_tfft.txt 2>&1; echo "racecheck rc=$?" >> gpurun_out/${T}_sanitize_racecheck_tfft.txt; tail -6 gpurun_out/${T}_sanitize_racecheck_tfft.txt
