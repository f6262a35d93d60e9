mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_resamp -s 12 -c 2 -o /tmp/rs python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_rs.log 2>&1
python scripts/ncu_summary.py full /tmp/rs.ncu-rep > gpurun_out/r1k_resamp_full.txt 2>&1
python scripts/ncu_source.py /tmp/rs.ncu-rep 10 > gpurun_out/r1k_resamp_stalls.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_leg2map|k_map2leg" -s 2 -c 2 -o /tmp/rf python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_rf.log 2>&1
python scripts/ncu_summary.py full /tmp/rf.ncu-rep > gpurun_out/r1k_ringfft_full.txt 2>&1
python scripts/ncu_source.py /tmp/rf.ncu-rep 8 > gpurun_out/r1k_ringfft_stalls.txt 2>&1
