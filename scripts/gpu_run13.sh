set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_adj2 -s 1 -c 1 -o gpurun_out/r1g_adj2_v7 python scripts/tune_legendre.py c3 0 3 > gpurun_out/ncu_adj2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_synth2 -s 1 -c 1 -o gpurun_out/r1g_synth2_v5 python scripts/tune_legendre.py c3 0 2 > gpurun_out/ncu_synth2.log 2>&1
