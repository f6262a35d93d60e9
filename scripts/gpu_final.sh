set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; tail -2 gpurun_out/smoke.txt
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err; cut -c1-300 gpurun_out/bench_final.json
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; cut -c1-300 gpurun_out/bench_final_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file /tmp/launches_c3.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
python scripts/ncu_summary.py launches /tmp/launches_c3.csv > gpurun_out/launches_c3.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_adj|k_synth" -s 8 -c 4 -o /tmp/leg_full python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/bench_under_ncu2.log 2>&1
python scripts/ncu_summary.py full /tmp/leg_full.ncu-rep > gpurun_out/leg_full.txt 2>&1
python scripts/ncu_source.py /tmp/leg_full.ncu-rep 8 > gpurun_out/leg_stalls.txt 2>&1
