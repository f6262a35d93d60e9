set -x
mkdir -p gpurun_out
# launch list of one bench run (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file /tmp/launches_c3.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
python scripts/ncu_summary.py launches /tmp/launches_c3.csv > gpurun_out/launches_c3.txt 2>&1
# full metric set on the engine's kernels of one step (reports stay on the box; summaries come back)
ncu --set full --clock-control none --import-source on -k regex:"k_adj|k_synth|k_resamp|k_leg2map|k_map2leg" -s 40 -c 24 -o /tmp/step_full python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/bench_under_ncu2.log 2>&1
python scripts/ncu_summary.py full /tmp/step_full.ncu-rep > gpurun_out/step_full.txt 2>&1
