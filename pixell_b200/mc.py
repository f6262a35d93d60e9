"""pixell_b200.mc -- batches of curvedsky.rand_map realisations sharded over the GPUs of one box.

The reference draws realisations one at a time (pixell/curvedsky.py:17-36 -> rand_alm_healpy / rand_alm
:44-77 -> alm2map) and leaves any parallelism to the caller (MPI scripts).  The units are independent,
so the multi-GPU form is plain sharding: one process per GPU (torchrun), realisations
block-partitioned over ranks, the input C_l broadcast from rank 0 (NCCL through torch.distributed; gloo
in the CPU tests), maps and alm never leave their GPU.  Optionally the per-realisation power spectra are
gathered.  Nothing here is a data-path collective.

Random streams:
  rng="reference"  numpy's legacy global stream on the host, seeded per realisation (curvedsky.rand_alm_healpy): for a
                   scalar spectrum the alm are healpy.synalm's, i.e. the reference rand_map's realisation for that seed
                   (pinned by the reference's golden MM_041121.pkl); for T,Q,U the stream is pixell's rand_alm one (l-major
                   fill, symmetric square root), NOT healpy's polarised synalm -- same covariance, other numbers.
  rng="device"     the engine's own kernel (b2_rand_alm): counter-based Philox4x32-10 normals in the reference's fill
                   order, coloured with the symmetric square root of C_l and fixed at m = 0 in the same kernel
                   (curvedsky.rand_alm :61-77, rand_alm_white :620-628): same statistics, different numbers.
"""
import numpy as np
from . import _lib as L, curvedsky, geometry, sht

def world():
	"""(rank, world_size) of the torch.distributed job, (0, 1) outside one"""
	try:
		import torch.distributed as dist
		if dist.is_available() and dist.is_initialized(): return dist.get_rank(), dist.get_world_size()
	except ImportError: pass
	return 0, 1

def partition(n, nrank=None, rank=None):
	"""Block partition of n units: the half-open index range of `rank` (first n % nrank ranks get one extra)."""
	if nrank is None or rank is None:
		r, w = world(); rank = r if rank is None else rank; nrank = w if nrank is None else nrank
	base, extra = divmod(int(n), int(nrank))
	lo = rank*base + min(rank, extra)
	return range(lo, lo + base + (1 if rank < extra else 0))

def broadcast_ps(ps, src=0, device=None):
	"""The input spectrum is known on rank `src` only (e.g. read from disk there): every rank gets a copy.
	ps may be None on the other ranks; the shape travels first."""
	rank, nrank = world()
	if nrank == 1: return np.asarray(ps, dtype=np.float64)
	import torch, torch.distributed as dist
	dev = device if device is not None else (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu"))
	hdr = torch.zeros(5, dtype=torch.int64, device=dev)
	if rank == src:
		ps = np.ascontiguousarray(ps, dtype=np.float64)
		hdr[0] = ps.ndim; hdr[1:1+ps.ndim] = torch.tensor(ps.shape, dtype=torch.int64)
	dist.broadcast(hdr, src)
	shape = tuple(int(v) for v in hdr[1:1+int(hdr[0])].tolist())
	buf = torch.from_numpy(ps).to(dev) if rank == src else torch.empty(shape, dtype=torch.float64, device=dev)
	dist.broadcast(buf, src)
	return buf.cpu().numpy()

def gather_rows(local, counts=None):
	"""All ranks get the concatenation (rank order) of every rank's rows: used for the small per-realisation
	results (spectra), never for maps.  local: numpy [nlocal, ...]."""
	rank, nrank = world()
	if nrank == 1: return np.asarray(local)
	import torch, torch.distributed as dist
	dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
	local = np.ascontiguousarray(local)
	n = torch.tensor([local.shape[0]], dtype=torch.int64, device=dev)
	ns = [torch.zeros_like(n) for _ in range(nrank)]
	dist.all_gather(ns, n)
	ns = [int(v.item()) for v in ns]
	nmax = max(ns)
	pad = np.zeros((nmax,)+local.shape[1:], local.dtype); pad[:local.shape[0]] = local
	bufs = [torch.empty(pad.shape, dtype=torch.from_numpy(pad).dtype, device=dev) for _ in range(nrank)]
	dist.all_gather(bufs, torch.from_numpy(pad).to(dev))
	return np.concatenate([b.cpu().numpy()[:k] for b, k in zip(bufs, ns)], 0)

def _wps(ps, ncomp, lmax):
	ps = np.asarray(ps, dtype=np.float64)
	if ps.ndim == 1: ps = ps[None, None]
	elif ps.ndim == 2: ps = curvedsky.sym_expand(ps)
	ps = ps[:ncomp, :ncomp]
	if ps.shape[-1] < lmax+1: ps = curvedsky.pad_spectrum(ps, lmax)
	return ps[..., :lmax+1]

def rand_alm_device(ps12, ainfo, seed, device, dtype=None, out=None):
	"""Coloured Gaussian alm on the GPU in one kernel launch (b2_rand_alm): Philox normals in the reference's fill order,
	alm <- ps^(1/2)/sqrt(2) alm, m = 0 made real with the sqrt(2) restored (curvedsky.rand_alm :61-77).
	ps12: [ncomp, ncomp, lmax+1] float64 torch CUDA tensor (or None for white alm)."""
	import torch, ctypes
	ncomp = 1 if ps12 is None else ps12.shape[0]
	dt = torch.complex128 if dtype is None else dtype
	if out is None: out = torch.zeros((ncomp, ainfo.nelem), dtype=dt, device=device)
	else: out.zero_()
	ms = L.as_i64(ainfo.mstart)
	if ainfo.stride != 1: raise NotImplementedError("rand_alm_device needs a unit-stride alm layout")
	if ps12 is not None:
		ps12 = ps12.contiguous()
		if ps12.shape[-1] != ainfo.lmax+1 or ps12.dtype != torch.float64: raise ValueError("ps12 must be float64 [ncomp, ncomp, lmax+1]")
	L.check(L.lib().b2_rand_alm(ainfo.lmax, ainfo.mmax, L.p_i64(ms), ncomp, ctypes.c_uint64(int(seed) & (2**64-1)),
		None if ps12 is None else ps12.data_ptr(), L.F64 if dt == torch.complex128 else L.F32, out.data_ptr(), out.stride(0),
		L.MEM_DEVICE, L.current_stream(out)))
	return out

def rand_maps(shape, wcs, ps, seeds, lmax=None, spin=[0, 2], rng="reference", device=None, out=None, return_alm=False, batch=4):
	"""This rank's share of the realisations `seeds` (block partition): a torch CUDA tensor
	[nlocal, ncomp, ny, nx] of maps (float64), realisation i of the share = seed seeds[partition[i]].
	ps: [ncomp,ncomp,nl], [nspec,nl] or [nl], already present on every rank (see broadcast_ps)."""
	import torch
	L.init()
	if device is None: device = torch.device("cuda", torch.cuda.current_device())
	seeds = list(seeds)
	mine = partition(len(seeds))
	ncomp = 1 if len(shape) == 2 else shape[-3]
	ny, nx = shape[-2:]
	if lmax is None: lmax = np.asarray(ps).shape[-1]-1
	wps = _wps(ps, ncomp, lmax)
	ainfo = curvedsky.alm_info(lmax)
	if out is None: out = torch.empty((len(mine), ncomp, ny, nx), dtype=torch.float64, device=device)
	ps12 = None
	alms = []
	def draw(k, dst=None):
		nonlocal ps12
		if rng == "reference":
			a = curvedsky.rand_alm_healpy(wps[0, 0] if ncomp == 1 else wps, lmax=lmax, seed=seeds[k])
			alm = torch.from_numpy(np.atleast_2d(a)).to(device)
			if dst is not None: dst.copy_(alm); alm = dst
			return alm
		if rng == "device":
			if ps12 is None: ps12 = torch.as_tensor(curvedsky.multi_pow_half(wps), device=device)
			return rand_alm_device(ps12, ainfo, seeds[k], device, out=dst)
		raise ValueError("rng must be 'reference' or 'device'")
	# blocks of realisations through the batched synthesis (the members of a block share the Legendre recurrence: about a
	# quarter fewer FP64 instructions per T,Q,U realisation); anything the batch path does not cover goes one by one
	plan = _batch_plan(out.shape[1:], wcs, ainfo) if batch > 1 else None
	i = 0
	while i < len(mine):
		nb = min(batch, len(mine)-i) if plan is not None else 1
		if nb >= 2:
			blk = torch.empty((nb, ncomp, ainfo.nelem), dtype=torch.complex128, device=device)
			for b in range(nb): draw(mine[i+b], blk[b])
			for s, j1, j2 in curvedsky.spin_helper(spin, ncomp): sht.run_batch(plan, s, blk[:, j1:j2], out[i:i+nb, j1:j2])
			if return_alm: alms.extend(blk[b] for b in range(nb))
		else:
			alm = draw(mine[i])
			curvedsky.alm2map(alm, out[i], spin=spin, ainfo=ainfo, wcs=wcs)
			if return_alm: alms.append(alm)
		i += nb
	return (out, alms) if return_alm else out

def _batch_plan(shape, wcs, ainfo):
	"""the engine plan curvedsky.alm2map would use for maps of this geometry, or None when the batched call does not apply
	(non-cylindrical geometry)"""
	minfo = curvedsky.analyse_geometry(shape, wcs)
	method = curvedsky.get_method(shape, wcs, minfo=minfo)
	if method not in ("2d", "cyl"): return None
	return curvedsky._plan_of(curvedsky._plan_kwargs(shape, wcs, minfo, method, ainfo, ainfo.lmax, ainfo.mmax))
