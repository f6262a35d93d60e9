"""Host logic of the multi-GPU Monte-Carlo sharding (pixell_b200.mc) on CPU: block partition and the
C_l broadcast / result gather over torch.distributed with the gloo backend, world_size 2."""
import os, sys, socket
import numpy as np, pytest
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

def test_partition_covers_everything():
	from pixell_b200 import mc
	for n in (0, 1, 7, 64, 65):
		for w in (1, 2, 3, 8):
			got = [i for r in range(w) for i in mc.partition(n, w, r)]
			assert got == list(range(n))
			sizes = [len(mc.partition(n, w, r)) for r in range(w)]
			assert max(sizes)-min(sizes) <= 1

def _worker(rank, world, port, q):
	os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
	sys.path.insert(0, ROOT)
	import torch.distributed as dist
	from pixell_b200 import mc
	dist.init_process_group("gloo", rank=rank, world_size=world)
	try:
		ps = np.arange(3*3*17, dtype=np.float64).reshape(3, 3, 17) if rank == 0 else None
		got = mc.broadcast_ps(ps, src=0)
		mine = mc.partition(5)
		local = np.array([[k, 10*k] for k in mine], dtype=np.float64).reshape(len(mine), 2)
		allrows = mc.gather_rows(local)
		q.put((rank, got.shape, float(got.sum()), list(mine), allrows.tolist()))
	finally:
		dist.destroy_process_group()

def test_broadcast_and_gather_gloo_world2():
	import torch.multiprocessing as mp
	s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
	ctx = mp.get_context("spawn")
	q = ctx.Queue()
	procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
	for p in procs: p.start()
	res = sorted(q.get(timeout=120) for _ in range(2))
	for p in procs: p.join(timeout=60)
	want_sum = float(np.arange(3*3*17).sum())
	assert [r[3] for r in res] == [[0, 1, 2], [3, 4]]
	for rank, shape, tot, mine, rows in res:
		assert shape == (3, 3, 17) and tot == want_sum
		assert rows == [[k, 10.0*k] for k in range(5)]
