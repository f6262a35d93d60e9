"""GPU tests of the pixell-shaped layer (pixell_b200.curvedsky / cmisc), written after the reference's
own tests for this path (reference tests/test_pixell.py) and checked against the CPU oracle."""
import os
import numpy as np, pytest
from conftest import GOLDEN

pytestmark = pytest.mark.gpu
from oracle import sht_oracle as so, alm_oracle as ao, pixell_ref as pr

@pytest.fixture(scope="module")
def cs():
	from pixell_b200 import curvedsky
	return curvedsky

@pytest.fixture(scope="module")
def geom():
	from pixell_b200 import geometry
	return geometry

def relerr(a, b): return np.abs(np.asarray(a)-np.asarray(b)).max()/max(np.abs(b).max(), 1e-300)
def geo_of(shape, wcs): return pr.Geo(shape[-2:], wcs.wcs.crval, wcs.wcs.cdelt, wcs.wcs.crpix)

def test_alm2map_2d_roundtrip(cs, geom):
	"""reference tests/test_pixell.py:870-965: F1 32x61 (prime nphi), lmax 30, single alm (30,30)=1+1j,
	spin 0 and spin 1, f64 and f32, 1/2/3-d alm, with and without a preallocated output"""
	lmax = 30
	ainfo = cs.alm_info(lmax=lmax)
	nrings, nphi = lmax+2, 2*lmax+1
	shape, wcs = geom.fullsky_geometry(shape=(nrings, nphi))
	for spin in (0, 1):
		for rdt in (np.float64, np.float32):
			ctype = np.result_type(rdt, 0j)
			for pre in ((), (2,), (3, 2)) if spin == 0 else ((2,), (3, 2)):
				alm = np.zeros(pre+(ainfo.nelem,), ctype)
				i = ainfo.lm2ind(lmax, lmax)
				alm[..., i] = 1+1j
				omap = geom.zeros(pre+shape, wcs, rdt)
				cs.alm2map(alm, omap, spin=spin)
				alm_out = cs.map2alm(omap, spin=spin, ainfo=ainfo)
				np.testing.assert_array_almost_equal(alm_out, alm, decimal=6)
				alm_out = np.zeros_like(alm)
				cs.map2alm(omap, alm=alm_out, spin=spin, ainfo=ainfo)
				np.testing.assert_array_almost_equal(alm_out, alm, decimal=6)
				# tight check of the map against the oracle
				if rdt == np.float64 and len(pre) == 1:
					want = np.zeros(pre+shape); pr.alm2map(alm.astype(np.complex128), want, geo_of(shape, wcs), spin=[spin])
					assert relerr(omap, want) < 1e-12

def test_alm_conversion(cs, geom):
	"""reference tests/test_pixell.py:1028-1046"""
	lmax = 16
	shape, wcs = geom.fullsky_geometry(shape=(lmax+1, 2*lmax+2))
	ainfo = cs.alm_info(lmax)
	alm = np.zeros(ainfo.nelem, np.complex64); alm[5] = 1
	omap = geom.zeros(shape, wcs, np.float64)
	cs.alm2map(alm, omap, spin=0)                       # c64 alm into an f64 map is converted
	assert np.abs(omap).max() > 0
	with pytest.raises(ValueError):
		cs.map2alm(omap, alm=np.zeros(ainfo.nelem, np.complex64), spin=0)

def _zip_alm(alm, ainfo):
	"""real orthonormal alm basis of the reference's adjointness helpers (tests/test_pixell.py:218-230):
	m = 0 real parts, then sqrt(2) (re, im) of the m > 0 entries"""
	n = int(ainfo.lm2ind(1, 1))
	return np.concatenate([alm[..., :n].real, alm[..., n:].view(np.float64)*2**0.5], -1)

def _unzip_alm(zalm, ainfo):
	n = int(ainfo.lm2ind(1, 1))
	oalm = np.zeros(zalm.shape[:-1]+(ainfo.nelem,), np.complex128)
	oalm[..., :n] = zalm[..., :n]
	oalm[..., n:] = zalm[..., n:].view(np.complex128)/2**0.5
	return oalm

@pytest.mark.parametrize("variant,ny,nx", [("fejer1", 6, 12), ("cc", 7, 12)])
@pytest.mark.parametrize("ncomp", [1, 3])
def test_adjointness(cs, geom, variant, ny, nx, ncomp):
	"""reference tests/test_pixell.py:1051-1085 (full-sky geometries): alm2map_adjoint == alm2map^T in the
	zipped real alm basis (alm_bash / map_bash)"""
	shape, wcs = geom.fullsky_geometry(shape=(ny, nx), variant=variant)
	lmax = 5
	ainfo = cs.alm_info(lmax)
	nz = 2*ainfo.nelem - int(ainfo.lm2ind(1, 1))
	spin = [0, 2] if ncomp == 3 else [0]
	mat1 = np.zeros((ncomp, nz, ncomp)+shape)
	for ci in range(ncomp):
		for i in range(nz):
			z = np.zeros((ncomp, nz)); z[ci, i] = 1
			omap = geom.zeros((ncomp,)+shape, wcs)
			cs.alm2map(_unzip_alm(z, ainfo), omap, spin=spin, ainfo=ainfo)
			mat1[ci, i] = omap
	mat2 = np.zeros((ncomp, nz, ncomp)+shape)
	for I in np.ndindex(*((ncomp,)+shape)):
		umap = geom.zeros((ncomp,)+shape, wcs); umap[I] = 1
		oalm = np.zeros((ncomp, ainfo.nelem), np.complex128)
		cs.alm2map_adjoint(umap, alm=oalm, spin=spin, ainfo=ainfo)
		mat2[(slice(None), slice(None))+I] = _zip_alm(oalm, ainfo)
	np.testing.assert_array_almost_equal(mat1, mat2, decimal=12)

def test_cyl_band_and_cut_sky(cs, geom):
	"""method 'cyl': a declination band (full rows) and a cut-sky patch (partial rows), both spins, with the
	reference's weights and Jacobi iterations, against the oracle's restatement of the same host logic"""
	lmax = 40
	rng = np.random.default_rng(3)
	ai = ao.AlmInfo(lmax)
	alm = rng.standard_normal((3, ai.nelem)) + 1j*rng.standard_normal((3, ai.nelem))
	alm[:, :lmax+1] = alm[:, :lmax+1].real; alm[1:, [0, 1, lmax+1]] = 0
	for variant in ("fejer1", "cc"):
		band = geom.band_geometry(np.deg2rad([-30, 50]), res=np.deg2rad(2.0), variant=variant)
		fs, fw = geom.fullsky_geometry(res=np.deg2rad(2.0), variant=variant)
		patch = geom.slice_geometry(fs, fw, 12, 70, 20, 130)
		for shape, wcs in (band, patch):
			assert cs.get_method(shape, wcs) == "cyl"
			m = geom.zeros((3,)+shape, wcs)
			cs.alm2map(alm, m, spin=[0, 2])
			want = np.zeros((3,)+shape); pr.alm2map(alm, want, geo_of(shape, wcs), spin=[0, 2])
			assert relerr(m, want) < 1e-12
			for niter in (0, 2):
				got = cs.map2alm(m, lmax=lmax, spin=[0, 2], niter=niter)
				ref = pr.map2alm(want, geo_of(shape, wcs), lmax=lmax, spin=[0, 2], niter=niter)
				assert relerr(got, ref) < 1e-11
			# derivative maps (reference curvedsky.py:918-920 sign convention, pinned by MM_lensed golden via the oracle)
			d = geom.zeros((2,)+shape, wcs)
			cs.alm2map(alm[0], d, deriv=True)
			dw = np.zeros((2,)+shape); pr.alm2map(alm[:1], dw, geo_of(shape, wcs), deriv=True)
			assert relerr(d, dw) < 1e-12

def test_golden_unlensed_through_curvedsky(cs, geom):
	"""reference tests/test_pixell.py:351-360 end to end: curvedsky.rand_alm(seed=1) (numpy stream on the host,
	transpose + colouring on the GPU) -> curvedsky.alm2map(spin=[0,2]) == MM_unlensed_071123.fits"""
	g = np.load(os.path.join(GOLDEN, "unlensed_071123.npz"))
	ps = np.load(os.path.join(GOLDEN, "lens_ps_400.npy"))
	alm = cs.rand_alm(ps, lmax=400, seed=1)
	assert relerr(alm, ao.rand_alm(ps, 400, 1)) < 1e-14
	wcs = geom.CarWCS(g["crval"], g["cdelt"], g["crpix"])
	m = geom.zeros((3,)+tuple(int(v) for v in g["shape"][1:]), wcs)
	cs.alm2map(alm[1:], m, spin=[0, 2])
	got, want = np.asarray(m)[:, g["rows"]], g["map"]
	assert np.allclose(got, want, rtol=1e-9, atol=1e-10)

def test_golden_pixels_rand_map(cs, geom):
	"""reference tests/test_pixell.py:568-580: rand_map(seed=10, lmax=1500) on CC 1081x2160, 36 reference pixels"""
	g = np.load(os.path.join(GOLDEN, "pixels_041121.npz"))
	lmax = 1500
	shape, wcs = geom.fullsky_geometry(res=np.deg2rad(10/60), variant="cc")
	l = np.arange(lmax+500)
	am = (np.pi/180/60)**2
	dl = np.zeros(l.size); dl[2:] = 3.0**2*am*2*np.pi/(l[2:]*(l[2:]+1.0))
	off = np.array([0, 1, 29, 30, 58, 59]); rows = 510+off; cols = (2130+off) % 2160
	for name, cl in (("white_10", np.full(l.size, 10.0**2*am)), ("constant_dl_1", dl)):
		m = cs.rand_map(shape, wcs, cl, lmax=lmax, seed=10)
		got = np.asarray(m)[np.ix_(rows, cols)]
		assert np.allclose(got, g[name].reshape(got.shape), rtol=1e-9, atol=1e-10), name

def test_almxfl_and_lens_alms(cs):
	"""reference tests/test_pixell.py:850-868, 827-836"""
	lmax = 30
	ainfo = cs.alm_info(lmax)
	rng = np.random.default_rng(1)
	for pre in ((), (3,)):
		alm = rng.standard_normal(pre+(ainfo.nelem,)) + 1j*rng.standard_normal(pre+(ainfo.nelem,))
		np.testing.assert_array_almost_equal(cs.almxfl(alm, np.ones(lmax+1)), alm, decimal=14)
		np.testing.assert_array_almost_equal(cs.almxfl(alm, lambda l: np.ones(l.size)), alm, decimal=14)
		f = lambda l: l*(l+1.)/2
		g = lambda l: np.where(l > 0, 2/(l*(l+1.)+(l == 0)), 0)
		back = cs.almxfl(cs.almxfl(alm, f), g)
		mask = np.ones(ainfo.nelem, bool); mask[0] = False
		np.testing.assert_array_almost_equal(back[..., mask], alm[..., mask], decimal=12)
		assert relerr(cs.almxfl(alm, f), ao.lmul(ao.AlmInfo(lmax), alm, f(np.arange(lmax+1.0)))) < 1e-15

def test_cmisc_vs_oracle(cs):
	from pixell_b200 import cmisc
	rng = np.random.default_rng(2)
	for lmax, mmax in ((37, 37), (50, 20)):
		ai, oi = cs.alm_info(lmax, mmax), ao.AlmInfo(lmax, mmax)
		alm = rng.standard_normal((3, ai.nelem)) + 1j*rng.standard_normal((3, ai.nelem))
		# alm2cl, broadcasting 3x3 with the duplicate-pair cache (reference curvedsky.py:672-712)
		cl = cs.alm2cl(alm[:, None], alm[None, :], ainfo=ai)
		assert cl.shape == (3, 3, lmax+1)
		for i in range(3):
			for j in range(3): assert relerr(cl[i, j], ao.alm2cl(oi, alm[i], alm[j])) < 1e-13
		cl32 = cs.alm2cl(alm[0].astype(np.complex64), ainfo=ai)
		assert cl32.dtype == np.float32 and relerr(cl32, ao.alm2cl(oi, alm[0])) < 1e-5
		assert cs.alm2cl(alm[0].astype(np.complex64), ainfo=ai, dtype=np.float64).dtype == np.float64
		with pytest.raises(TypeError): cs.alm2cl(alm[0], ainfo=ai, dtype=np.float32)
		# lmul matrix path incl. lfmax below / above lmax, in place
		for lfmax in (lmax, lmax-7, lmax+5):
			M = rng.standard_normal((3, 3, lfmax+1))
			assert relerr(ai.lmul(alm, M), ao.lmul(oi, alm, M)) < 1e-14
			a2 = alm.copy(); ai.lmul(a2, M, out=a2)
			assert relerr(a2, ao.lmul(oi, alm, M)) < 1e-14
			f = rng.standard_normal(lfmax+1)
			assert relerr(ai.lmul(alm, f), ao.lmul(oi, alm, f)) < 1e-15
		if mmax == lmax:
			assert np.array_equal(ai.transpose_alm(alm), ao.transpose_alm(oi, alm))
			a3 = alm.copy(); ai.transpose_alm(a3, a3); assert np.array_equal(a3, ao.transpose_alm(oi, alm))
		ai2, oi2 = cs.alm_info(lmax-9, min(mmax, lmax-9)), ao.AlmInfo(lmax-9, min(mmax, lmax-9))
		assert np.array_equal(cs.transfer_alm(ai, alm, ai2), ao.transfer_alm(oi, alm, oi2))

def test_torch_tensors_through_curvedsky(cs, geom):
	import torch
	lmax = 64
	shape, wcs = geom.fullsky_geometry(shape=(96, 192))
	ai = cs.alm_info(lmax)
	rng = np.random.default_rng(4)
	alm = rng.standard_normal((3, ai.nelem)) + 1j*rng.standard_normal((3, ai.nelem))
	alm[:, :lmax+1] = alm[:, :lmax+1].real; alm[1:, [0, 1, lmax+1]] = 0
	talm = torch.from_numpy(alm).cuda()
	tmap = torch.empty((3,)+shape, dtype=torch.float64, device="cuda")
	cs.alm2map(talm, tmap, spin=[0, 2], wcs=wcs)
	want = np.zeros((3,)+shape); pr.alm2map(alm, want, geo_of(shape, wcs), spin=[0, 2])
	assert relerr(tmap.cpu().numpy(), want) < 1e-12
	back = cs.map2alm(tmap, lmax=lmax, spin=[0, 2], wcs=wcs)
	assert back.is_cuda and relerr(back.cpu().numpy(), alm) < 1e-12
	cl = cs.alm2cl(talm[0], ainfo=ai)
	assert cl.is_cuda and relerr(cl.cpu().numpy(), ao.alm2cl(ao.AlmInfo(lmax), alm[0])) < 1e-13

def test_rand_alm_shapes(cs):
	"""reference tests/test_pixell.py:320-337 (test_rand_alm): rand_alm and rand_alm_healpy agree on shapes; same seed, same alm"""
	mypower = np.ones(50)
	for lmax in [50, 100, 150, 300]:
		palm = cs.rand_alm(mypower, lmax=lmax)
		halm = cs.rand_alm_healpy(mypower, lmax=lmax)
		assert palm.shape == halm.shape == ((lmax+1)*(lmax+2)//2,)
	a = cs.rand_alm(mypower, lmax=60, seed=3); b = cs.rand_alm(mypower, lmax=60, seed=3)
	assert np.array_equal(a, b) and np.all(a[:61].imag == 0)

@pytest.mark.parametrize("geom,ny,nx,lmax", [("F1", 4608, 2048, 1000), ("CC", 4611, 2050, 900)])
def test_streamed_host_transforms_match_the_device_path(geom, ny, nx, lmax):
	"""host-memory calls on large grids run the synthesis in chunks of ring pairs and the adjoint Legendre stage in ranges
	of m, with the results leaving for the host chunk by chunk (api.cu): same kernels on the same data, so the numbers
	must be identical to the device-resident call"""
	import torch
	from pixell_b200 import sht
	rng = np.random.default_rng(4)
	nalm = (lmax+1)*(lmax+2)//2
	mstart = sht.default_mstart(lmax, lmax)
	for spin in (0, 2):
		nc = 1 if spin == 0 else 2
		alm = rng.standard_normal((nc, nalm)) + 1j*rng.standard_normal((nc, nalm))
		alm[:, :lmax+1] = alm[:, :lmax+1].real
		for m in range(spin): alm[:, mstart[m]+np.arange(m, spin)] = 0
		kw = dict(spin=spin, lmax=lmax, geometry=geom, phi0=0.1)
		for flip_y in (False, True):
			dev = sht.synthesis_2d(alm=torch.from_numpy(alm).cuda(), ntheta=ny, nphi=nx, flip_y=flip_y, **kw).cpu().numpy()
			host = np.full((nc, ny, nx), np.nan)
			sht.synthesis_2d(alm=alm, map=host, flip_y=flip_y, **kw)
			assert np.array_equal(host, dev)
		back_dev = sht.adjoint_synthesis_2d(map=torch.from_numpy(dev).cuda(), flip_y=True, **kw).cpu().numpy()
		back_host = np.full((nc, nalm), np.nan+0j)
		sht.adjoint_synthesis_2d(map=dev, alm=back_host, flip_y=True, **kw)
		assert np.array_equal(back_host, back_dev)
		a_dev = sht.analysis_2d(map=torch.from_numpy(dev).cuda(), flip_y=True, **kw).cpu().numpy()
		a_host = sht.analysis_2d(map=dev, flip_y=True, **kw)
		assert np.array_equal(a_host, a_dev)
		assert np.abs(a_host-alm).max() < 1e-11*np.abs(alm).max()

def test_grouped_host_pair_matches_the_device_path(cs):
	"""curvedsky.alm2map / map2alm on host arrays run all spin groups in ONE engine call (copies of one group under the
	kernels of the next, the first group's alm gated range by range, the last group's alm leaving while the Legendre
	adjoint still runs): the numbers must be those of the device-resident call"""
	import torch
	from pixell_b200 import geometry
	lmax, ny, nx = 900, 4608, 2048
	shape, wcs = geometry.fullsky_geometry(shape=(ny, nx))
	ai = cs.alm_info(lmax)
	rng = np.random.default_rng(11)
	alm = rng.standard_normal((3, ai.nelem)) + 1j*rng.standard_normal((3, ai.nelem))
	alm[:, :lmax+1] = alm[:, :lmax+1].real
	alm[1:, [0, 1, lmax+1]] = 0
	dmap = torch.empty((3,)+shape, dtype=torch.float64, device="cuda")
	cs.alm2map(torch.from_numpy(alm).cuda(), dmap, spin=[0, 2], wcs=wcs, ainfo=ai)
	hmap = geometry.ndmap(np.full((3,)+shape, np.nan), wcs)
	cs.alm2map(alm, hmap, spin=[0, 2], ainfo=ai)
	assert np.array_equal(np.asarray(hmap), dmap.cpu().numpy())
	dback = torch.zeros((3, ai.nelem), dtype=torch.complex128, device="cuda")
	cs.map2alm(dmap, dback, spin=[0, 2], wcs=wcs, ainfo=ai)
	hback = np.full((3, ai.nelem), np.nan+0j)
	cs.map2alm(hmap, hback, spin=[0, 2], ainfo=ai)
	assert np.array_equal(hback, dback.cpu().numpy())
	assert np.abs(hback-alm).max() < 1e-11*np.abs(alm).max()

def test_calls_on_one_plan_do_not_overlap_in_its_scratch(cs):
	"""a device-memory call (asynchronous, on the caller's stream) followed at once by a host-memory call (on the plan's own
	streams) share the plan's leg / theta-stage scratch: the second must wait for the first (b2_sht_plan::ev_last).  With
	pinned host arrays the copies are short enough that the two would otherwise run at the same time."""
	import torch
	from pixell_b200 import geometry
	lmax, ny, nx = 900, 4608, 2048
	shape, wcs = geometry.fullsky_geometry(shape=(ny, nx))
	ai = cs.alm_info(lmax)
	rng = np.random.default_rng(12)
	alm = rng.standard_normal((3, ai.nelem)) + 1j*rng.standard_normal((3, ai.nelem))
	alm[:, :lmax+1] = alm[:, :lmax+1].real
	alm[1:, [0, 1, lmax+1]] = 0
	dmap = torch.empty((3,)+shape, dtype=torch.float64, device="cuda")
	cs.alm2map(torch.from_numpy(alm).cuda(), dmap, spin=[0, 2], wcs=wcs, ainfo=ai)
	want = torch.zeros((3, ai.nelem), dtype=torch.complex128, device="cuda")
	cs.map2alm(dmap, want, spin=[0, 2], wcs=wcs, ainfo=ai)
	torch.cuda.synchronize()
	want = want.cpu().numpy()
	pmap = torch.empty((3,)+shape, dtype=torch.float64, pin_memory=True); pmap.copy_(dmap)
	palm = torch.empty((3, ai.nelem), dtype=torch.complex128, pin_memory=True)
	hmap = geometry.ndmap(pmap.numpy(), wcs)
	for rep in range(3):
		dback = torch.zeros((3, ai.nelem), dtype=torch.complex128, device="cuda")
		palm.zero_()
		cs.map2alm(dmap, dback, spin=[0, 2], wcs=wcs, ainfo=ai)          # still running when the next call starts
		cs.map2alm(hmap, palm.numpy(), spin=[0, 2], ainfo=ai)
		torch.cuda.synchronize()
		assert np.array_equal(dback.cpu().numpy(), want)
		assert np.array_equal(palm.numpy(), want)
