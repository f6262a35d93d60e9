# multi-GPU bench line (N ranks on one box): python -m torch.distributed.run as the driver launches it
set -x
N=${1:-2}; T=${2:-r2w}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/${T}_bench_${N}gpu.json 2> gpurun_out/${T}_bench_${N}gpu.err; tail -3 gpurun_out/${T}_bench_${N}gpu.err; cut -c1-1500 gpurun_out/${T}_bench_${N}gpu.json
nvidia-smi topo -m > gpurun_out/${T}_topo_${N}gpu.txt 2>&1; head -14 gpurun_out/${T}_topo_${N}gpu.txt
