set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; cat gpurun_out/bench_2gpu.json | cut -c1-600; tail -5 gpurun_out/bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/bench_mc.py --nsim 32 > gpurun_out/mc_c4_2gpu.json 2> gpurun_out/mc_c4_2gpu.err; cat gpurun_out/mc_c4_2gpu.json; tail -3 gpurun_out/mc_c4_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 --cpu-seconds 8 > gpurun_out/bench_ref_2gpu.json 2> gpurun_out/bench_ref_2gpu.err; cat gpurun_out/bench_ref_2gpu.json | cut -c1-500; tail -3 gpurun_out/bench_ref_2gpu.err
