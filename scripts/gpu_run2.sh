set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; tail -3 gpurun_out/pytest_gpu.txt
python scripts/tune_legendre.py c3 6 0123 > gpurun_out/tune_c3_b.txt 2>&1
cat gpurun_out/tune_c3_b.txt
B2_LEG_VARIANT=0,0,0,2 ncu --set full --clock-control none --import-source on -k regex:k_adj2 -s 1 -c 1 -o gpurun_out/r1c_adj2_v2 python scripts/tune_legendre.py c3 0 3 > gpurun_out/ncu_adj2.log 2>&1
B2_LEG_VARIANT=0,0,4,0 ncu --set full --clock-control none --import-source on -k regex:k_synth2 -s 1 -c 1 -o gpurun_out/r1c_synth2_v4 python scripts/tune_legendre.py c3 0 2 > gpurun_out/ncu_synth2.log 2>&1
B2_LEG_VARIANT=0,0,2,0 ncu --set full --clock-control none --import-source on -k regex:k_synth2 -s 1 -c 1 -o gpurun_out/r1c_synth2_v2 python scripts/tune_legendre.py c3 0 2 >> gpurun_out/ncu_synth2.log 2>&1
ls -la gpurun_out
