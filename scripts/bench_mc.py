#!/usr/bin/env python
"""Configuration C4: a batch of 64 (T,Q,U) rand_map realisations from C_l at lmax 4096 on the
4608 x 9216 Fejer-1 grid, block-partitioned over the ranks of a torchrun job (1/2/4/8 GPUs of one box).
  python scripts/bench_mc.py [--nsim 64] [--rng device|reference] [--lmax 4096 --ny 4608 --nx 9216]
  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_mc.py ...
The only communication is the broadcast of the input C_l from rank 0.  Prints realisations/s (max over ranks)."""
import argparse, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def main():
	ap = argparse.ArgumentParser()
	ap.add_argument("--nsim", type=int, default=64); ap.add_argument("--rng", default="device")
	ap.add_argument("--lmax", type=int, default=4096); ap.add_argument("--ny", type=int, default=4608); ap.add_argument("--nx", type=int, default=9216)
	args = ap.parse_args()
	import torch.distributed as dist
	rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
	torch.cuda.set_device(local)
	if world > 1: dist.init_process_group("nccl", device_id=torch.device("cuda", local))
	from pixell_b200 import mc, geometry, _lib as L
	L.init(local)
	lmax = args.lmax
	ps = None
	if rank == 0:
		l = np.arange(lmax+1.0); tt = np.where(l >= 2, 1.0/np.maximum(l*(l+1), 1), 0.0)
		ps = np.zeros((3, 3, lmax+1)); ps[0, 0] = tt; ps[1, 1] = 0.3*tt; ps[2, 2] = 0.1*tt; ps[0, 1] = ps[1, 0] = 0.5*np.sqrt(ps[0, 0]*ps[1, 1])
	ps = mc.broadcast_ps(ps, src=0)
	shape, wcs = geometry.fullsky_geometry(shape=(args.ny, args.nx))
	seeds = list(range(1000, 1000+args.nsim))
	mine = mc.partition(len(seeds))
	out = torch.empty((len(mine), 3)+shape, dtype=torch.float64, device="cuda")
	mc.rand_maps((3,)+shape, wcs, ps, seeds[:world], lmax=lmax, rng=args.rng, out=out[:1])      # warm-up: plans, tables
	torch.cuda.synchronize()
	if world > 1: dist.barrier()
	e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	e0.record()
	mc.rand_maps((3,)+shape, wcs, ps, seeds, lmax=lmax, rng=args.rng, out=out)
	e1.record(); torch.cuda.synchronize()
	ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
	if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
	var_t = float(out[:, 0].var().item())
	if rank == 0:
		print(json.dumps({"workload": "C4: %d (T,Q,U) rand_map realisations, lmax %d, %dx%d, rng=%s" % (args.nsim, lmax, args.ny, args.nx, args.rng),
			"n_gpus": world, "ms_total": ms.item(), "realisations_per_s": args.nsim/(ms.item()*1e-3),
			"bytes_written_per_realisation": 8*3*args.ny*args.nx, "var_T_rank0": var_t}), flush=True)
	if world > 1: dist.destroy_process_group()

if __name__ == "__main__":
	main()
