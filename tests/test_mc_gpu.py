"""GPU tests of the sharded rand_map batches (pixell_b200.mc): the reference random stream reproduces
curvedsky.rand_map seed by seed; the device stream has the right covariance."""
import numpy as np, pytest
pytestmark = pytest.mark.gpu

def _ps(lmax):
	l = np.arange(lmax+1.0)
	tt = np.where(l >= 2, 1.0/np.maximum(l*(l+1), 1), 0.0)
	ps = np.zeros((3, 3, lmax+1))
	ps[0, 0] = tt; ps[1, 1] = 0.3*tt; ps[2, 2] = 0.1*tt; ps[0, 1] = ps[1, 0] = 0.5*np.sqrt(ps[0, 0]*ps[1, 1])
	return ps

def test_reference_stream_matches_rand_map():
	from pixell_b200 import mc, curvedsky, geometry
	lmax = 64
	shape, wcs = geometry.fullsky_geometry(shape=(80, 160))
	ps = _ps(lmax)
	seeds = [1000, 1001, 1002]
	maps = mc.rand_maps((3,)+shape, wcs, ps, seeds, lmax=lmax, rng="reference").cpu().numpy()
	assert maps.shape == (3, 3)+shape
	for i, s in enumerate(seeds):
		want = curvedsky.rand_map((3,)+shape, wcs, ps, lmax=lmax, seed=s)
		assert np.abs(maps[i]-np.asarray(want)).max() <= 1e-12*np.abs(want).max()

def test_device_stream_covariance():
	import torch
	from pixell_b200 import mc, curvedsky
	lmax = 200
	ps = _ps(lmax)
	ainfo = curvedsky.alm_info(lmax)
	ps12 = torch.as_tensor(curvedsky.multi_pow_half(ps), device="cuda")
	acc = np.zeros((3, 3, lmax+1)); n = 40
	for s in range(n):
		alm = mc.rand_alm_device(ps12, ainfo, 77+s, torch.device("cuda"))
		assert float(alm[:, :lmax+1].imag.abs().max()) == 0.0
		acc += curvedsky.alm2cl(alm[:, None], alm[None, :]).cpu().numpy()
	acc /= n
	l = np.arange(20, lmax+1)
	for (i, j) in [(0, 0), (1, 1), (2, 2), (0, 1)]:
		ratio = acc[i, j, l]/ps[i, j, l]
		# each C_l estimate averages n (2l+1) modes: relative scatter sqrt(2/(n(2l+1))) (x ~1.5 for the cross term)
		sig = np.sqrt(2.0/(n*(2*l+1)))*(2.0 if i != j else 1.0)
		assert np.all(np.abs(ratio-1) < 6*sig)
	# different seeds give different realisations, the same seed the same one
	a = mc.rand_alm_device(ps12, ainfo, 5, torch.device("cuda")); b = mc.rand_alm_device(ps12, ainfo, 5, torch.device("cuda"))
	c = mc.rand_alm_device(ps12, ainfo, 6, torch.device("cuda"))
	assert torch.equal(a, b) and not torch.equal(a, c)

@pytest.mark.parametrize("lmax,mmax,ncomp", [(50, 50, 3), (40, 25, 2), (7, 7, 1)])
def test_rand_alm_kernel_matches_the_numpy_restatement(lmax, mmax, ncomp):
	"""b2_rand_alm against oracle/alm_oracle.rand_alm_philox: Philox4x32-10 counters in the reference's fill order
	(pixell/curvedsky.py:602-628), Box-Muller, colouring and m = 0 fix (:61-77); white and coloured, host and device memory"""
	import torch, ctypes
	from pixell_b200 import _lib as L
	from oracle import alm_oracle as ao
	L.init()
	ai = ao.AlmInfo(lmax, mmax)
	rng = np.random.default_rng(3)
	A = rng.standard_normal((ncomp, ncomp, lmax+1)); ps12 = np.einsum("acl,bcl->abl", A, A)      # any symmetric matrix per l
	ms = L.as_i64(ai.mstart)
	for p12 in (None, ps12):
		want = ao.rand_alm_philox(ai, ncomp, 12345678901234, p12)
		got = np.full((ncomp, ai.nelem), 7+7j)
		L.check(L.lib().b2_rand_alm(lmax, mmax, L.p_i64(ms), ncomp, ctypes.c_uint64(12345678901234), None if p12 is None else p12.ctypes.data,
			L.F64, got.ctypes.data, ai.nelem, L.MEM_HOST, None))
		assert np.abs(got-want).max() <= 1e-13*np.abs(want).max()
		t = torch.zeros((ncomp, ai.nelem), dtype=torch.complex128, device="cuda")
		tp = None if p12 is None else torch.as_tensor(p12, device="cuda")
		L.check(L.lib().b2_rand_alm(lmax, mmax, L.p_i64(ms), ncomp, ctypes.c_uint64(12345678901234), None if tp is None else tp.data_ptr(),
			L.F64, t.data_ptr(), ai.nelem, L.MEM_DEVICE, None))
		assert np.array_equal(t.cpu().numpy(), got)

def test_rand_alm_kernel_shares_large_scales_between_lmax():
	"""the property the reference's l-major fill order exists for (pixell/curvedsky.py:62-66): realisations at two lmax agree
	on the common multipoles (first component; later components start after a longer block, as in the reference)"""
	import torch
	from pixell_b200 import mc, curvedsky
	lo, hi = curvedsky.alm_info(30), curvedsky.alm_info(60)
	a = mc.rand_alm_device(None, lo, 9, torch.device("cuda")).cpu().numpy()[0]
	b = mc.rand_alm_device(None, hi, 9, torch.device("cuda")).cpu().numpy()[0]
	for m in range(31):
		l = np.arange(m, 31)
		assert np.array_equal(a[int(lo.mstart[m])+l], b[int(hi.mstart[m])+l])
	# unit normals: mean 0, variance 1 per real and imaginary part
	big = mc.rand_alm_device(None, curvedsky.alm_info(400), 10, torch.device("cuda")).cpu().numpy()[0]
	assert abs(big.real.mean()) < 0.02 and abs(big.real.var()-1) < 0.03 and abs(big.imag.var()-1) < 0.03

@pytest.mark.parametrize("geom,ny,nx,lmax,ncomp", [("F1", 400, 800, 300, 3), ("CC", 301, 600, 280, 3), ("F1", 256, 512, 200, 1), ("F1", 2304, 1024, 500, 3)])
def test_batched_synthesis_is_bit_identical(geom, ny, nx, lmax, ncomp):
	"""mc.rand_maps runs blocks of realisations through the batched Legendre kernels (k_synth0b / k_synth2b: the members of
	a block share the recurrence); every member's accumulations are the single-map kernels' sequence of FMAs, so the maps
	must be bit-identical to the one-by-one path.  7 realisations: blocks of 4, 2 and a single one."""
	import torch
	from pixell_b200 import mc, geometry
	if geom == "F1": shape, wcs = geometry.fullsky_geometry(shape=(ny, nx))
	else: shape, wcs = geometry.fullsky_geometry(shape=(ny, nx), variant="CC")
	ps = _ps(lmax) if ncomp == 3 else _ps(lmax)[0, 0]
	seeds = list(range(40, 47))
	full = (ncomp,)+tuple(shape) if ncomp > 1 else tuple(shape)
	one = mc.rand_maps(full, wcs, ps, seeds, lmax=lmax, rng="device", batch=1)
	blk, alms = mc.rand_maps(full, wcs, ps, seeds, lmax=lmax, rng="device", batch=4, return_alm=True)
	assert torch.equal(one, blk) and len(alms) == len(seeds)
	three = mc.rand_maps(full, wcs, ps, seeds[:3], lmax=lmax, rng="device", batch=3)      # one pair + one single inside the engine call
	assert torch.equal(three, one[:3])
