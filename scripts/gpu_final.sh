# End-of-session validation on the GPU box: smoke, GPU test suite, bench lines (both arms), ncu launch list and the
# --set full summaries of the Legendre kernels, and the side benches (FFT, alm helpers, arbitrary positions, HEALPix,
# multi-scale loop).  Reports too large for gpurun_out/ stay in /tmp on the box; text summaries come back.
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; tail -2 gpurun_out/smoke.txt
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.txt 2>&1; tail -3 gpurun_out/pytest_gpu.txt
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err; cut -c1-300 gpurun_out/bench_final.json
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; cut -c1-300 gpurun_out/bench_final_ref.json
python scripts/bench_fft.py 4096 8192 3 3 > gpurun_out/fft_c5_small.json 2>&1; cat gpurun_out/fft_c5_small.json
python scripts/bench_fft.py 16384 32768 3 2 > gpurun_out/fft_c5.json 2>&1; cat gpurun_out/fft_c5.json
python scripts/bench_almops.py > gpurun_out/almops_c3.json 2>&1; tail -1 gpurun_out/almops_c3.json | cut -c1-400
python scripts/bench_general.py 4096 4608 9216 2 > gpurun_out/general_4096.json 2>&1; cat gpurun_out/general_4096.json
python scripts/bench_healpix.py 2048 4096 2 > gpurun_out/healpix_2048.json 2>&1; cat gpurun_out/healpix_2048.json
python scripts/bench_multiscale.py 4096 4608 9216 2 > gpurun_out/multiscale_4096.json 2>&1; cat gpurun_out/multiscale_4096.json
python scripts/bench_mc.py > gpurun_out/mc_c4_1gpu.json 2>&1; tail -1 gpurun_out/mc_c4_1gpu.json | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file /tmp/launches_c3.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
python scripts/ncu_summary.py launches /tmp/launches_c3.csv > gpurun_out/launches_c3.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_adj|k_synth" -s 8 -c 4 -o /tmp/leg_full python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/bench_under_ncu2.log 2>&1
python scripts/ncu_summary.py full /tmp/leg_full.ncu-rep > gpurun_out/leg_full.txt 2>&1
python scripts/ncu_source.py /tmp/leg_full.ncu-rep 8 > gpurun_out/leg_stalls.txt 2>&1
