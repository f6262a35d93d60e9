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
