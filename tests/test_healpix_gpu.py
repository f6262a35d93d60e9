"""GPU parity tests of ring sets with per-ring nphi / phi0 (b2_sht_plan_rings_general: HEALPix, single-pixel rings)
through pixell_b200.sht and the curvedsky mirrors alm2map_healpix / map2alm_healpix / profile2harm / harm2profile
(reference curvedsky.py:312-405, 1192-1234, 1544-1593).  Checker: the oracle's ring synthesis, which handles
arbitrary rings.  Tolerance 1e-11 relative to the largest value."""
import numpy as np, pytest

pytestmark = pytest.mark.gpu

def rel(a, b): return np.abs(a-b).max()/max(np.abs(b).max(), 1e-300)

def rand_alm(ncomp, lmax, seed):
	rng = np.random.default_rng(seed)
	nalm = (lmax+1)*(lmax+2)//2
	alm = rng.standard_normal((ncomp, nalm)) + 1j*rng.standard_normal((ncomp, nalm))
	alm[:, :lmax+1] = alm[:, :lmax+1].real
	return alm

@pytest.mark.parametrize("nside,lmax", [(4, 11), (16, 40), (32, 95)])
def test_healpix_synthesis_and_adjoint(nside, lmax):
	from pixell_b200 import curvedsky as cs, sht
	from oracle import sht_oracle as so
	ri = cs.get_ring_info_healpix(nside)
	assert ri.nrow == 4*nside-1 and int(ri.offsets[-1]+ri.nphi[-1]) == 12*nside**2
	kw = dict(theta=ri.theta, nphi=ri.nphi, phi0=ri.phi0, ringstart=ri.offsets, lmax=lmax)
	rng = np.random.default_rng(nside)
	for spin in (0, 2):
		nca = 1 if spin == 0 else 2
		alm = rand_alm(nca, lmax, 3+spin)
		if spin: alm[:, [0, 1, lmax+1]] = 0
		want = so.synthesis(alm=alm, spin=spin, **kw)
		got = sht.synthesis(alm=alm, spin=spin, **kw)
		assert got.shape == (nca, 12*nside**2) and rel(got, want) < 1e-11
		m = rng.standard_normal(want.shape)
		assert rel(sht.adjoint_synthesis(map=m, spin=spin, **kw), so.adjoint_synthesis(map=m, spin=spin, **kw)) < 1e-11

def test_alm2map_healpix_roundtrip_and_torch():
	"""T,Q,U through the pixell-shaped functions; map2alm_healpix with Jacobi iterations recovers a band-limited alm"""
	import torch
	from pixell_b200 import curvedsky as cs, sht
	from oracle import sht_oracle as so
	nside, lmax = 16, 24
	alm = rand_alm(3, lmax, 9); alm[1:, [0, 1, lmax+1]] = 0
	m = cs.alm2map_healpix(alm, nside=nside, spin=[0, 2])
	assert m.shape == (3, 12*nside**2)
	ri = cs.get_ring_info_healpix(nside)
	kw = dict(theta=ri.theta, nphi=ri.nphi, phi0=ri.phi0, ringstart=ri.offsets, lmax=lmax)
	want = np.concatenate([so.synthesis(alm=alm[:1], spin=0, **kw), so.synthesis(alm=alm[1:], spin=2, **kw)])
	assert rel(m, want) < 1e-11
	# device tensors go through the same plan
	tm = sht.synthesis(alm=torch.from_numpy(alm[1:]).cuda(), spin=2, **kw)
	assert tm.is_cuda and rel(tm.cpu().numpy(), want[1:]) < 1e-11
	back0 = cs.map2alm_healpix(m, lmax=lmax, spin=[0, 2], niter=0)
	back3 = cs.map2alm_healpix(m, lmax=lmax, spin=[0, 2], niter=3)
	e0, e3 = rel(back0, alm), rel(back3, alm)
	assert e3 < 1e-3 and e3 < 0.1*e0
	# the transpose pair
	d = cs.alm2map_healpix(alm[0], nside=nside, deriv=True)
	dw = so.synthesis(alm=alm[:1], spin=1, mode="DERIV1", **kw); dw[0] *= -1
	assert d.shape == (2, 12*nside**2) and rel(d, dw) < 1e-11

def test_theta_limits():
	from pixell_b200 import curvedsky as cs
	nside, lmax = 8, 12
	alm = rand_alm(1, lmax, 2)
	full = cs.alm2map_healpix(alm, nside=nside, spin=[0])
	cut = cs.alm2map_healpix(alm, nside=nside, spin=[0], theta_min=0.6, theta_max=2.0)
	ri = cs.get_ring_info_healpix(nside)
	inside = np.concatenate([np.arange(int(o), int(o+n)) for t, o, n in zip(ri.theta, ri.offsets, ri.nphi) if 0.6 <= t <= 2.0])
	mask = np.zeros(12*nside**2, bool); mask[inside] = True
	assert np.all(cut[0, ~mask] == 0) and rel(cut[0, mask], full[0, mask]) < 1e-12

def test_profile_transforms():
	"""harm2profile against the Legendre series, profile2harm back (reference curvedsky.py:1544-1593)"""
	from pixell_b200 import curvedsky as cs
	lmax = 60
	l = np.arange(lmax+1)
	bl = np.exp(-0.5*l*(l+1)*0.05**2)
	r = np.linspace(0, 0.8, 400)
	br = cs.harm2profile(bl, r)
	want = np.polynomial.legendre.legval(np.cos(r), bl*(2*l+1)/(4*np.pi))
	assert rel(br, want) < 1e-12
	rr = np.linspace(0, np.pi, 2001)
	back = cs.profile2harm(cs.harm2profile(bl, rr), rr, lmax=lmax)
	assert rel(back, bl) < 1e-6

def test_reproject_between_car_and_healpix():
	"""reproject.map2healpix / healpix2map, method "harm" (reference reproject.py:118-361): CAR -> HEALPix is exact for a
	band-limited sky (exact analysis, then synthesis on the HEALPix rings); with a rotation it equals rotate_alm in between"""
	from pixell_b200 import reproject, curvedsky as cs, geometry
	lmax, nside = 40, 32
	alm = rand_alm(3, lmax, 12); alm[1:, [0, 1, lmax+1]] = 0
	shape, wcs = geometry.fullsky_geometry(res=np.deg2rad(2.0))
	m = cs.alm2map(alm, geometry.zeros((3,)+shape, wcs), spin=[0, 2])
	heal = reproject.map2healpix(m, nside=nside, lmax=lmax)
	assert heal.shape == (3, 12*nside**2)
	assert rel(heal, cs.alm2map_healpix(alm, nside=nside, spin=[0, 2])) < 1e-10
	hrot = reproject.map2healpix(m, nside=nside, lmax=lmax, rot="cel,gal")
	ang = reproject.rot2euler("cel,gal")
	assert rel(hrot, cs.alm2map_healpix(cs.rotate_alm(alm, *ang), nside=nside, spin=[0, 2])) < 1e-9
	# and back: the HEALPix analysis is approximate (pixel weights + Jacobi iterations)
	back = reproject.healpix2map(heal, shape, wcs, lmax=lmax, niter=3)
	assert rel(np.asarray(back), np.asarray(m)) < 2e-3
	with pytest.raises(NotImplementedError): reproject.map2healpix(m, nside=nside, method="spline")
	assert reproject.restrict_nside(100, "pow2") == 128 and reproject.restrict_nside(100, "mul32") == 128 and reproject.restrict_nside(100.2, "any") == 101

def test_alm2map_healpix_roundtrip():
	"""mirror of the reference tests/test_pixell.py:967-1026: nside 2, lmax 4, one coefficient set, spin 0 and spin 1,
	1-, 2- and 3-dimensional alm, float64 and float32, with and without a preallocated output, niter = 7"""
	from pixell_b200 import curvedsky
	nside = 2
	lmax = nside*2
	nside = lmax//2
	ainfo = curvedsky.alm_info(lmax)
	npix = 12*nside**2
	niter = 7
	for use_oalm in [False, True]:
		for dtype in [np.float64, np.float32]:
			ctype = np.result_type(dtype, 0j)
			spin = 0
			alm = np.zeros((ainfo.nelem), dtype=ctype)
			i = ainfo.lm2ind(lmax, lmax)
			alm[i] = 1. + 1.j
			omap = np.zeros(npix, dtype)
			curvedsky.alm2map_healpix(alm, omap, spin=spin)
			assert np.any(omap != 0)
			alm_out = np.zeros_like(alm) if use_oalm else None
			alm_out = curvedsky.map2alm_healpix(omap, alm=alm_out, spin=spin, ainfo=ainfo, niter=niter)
			np.testing.assert_array_almost_equal(alm_out, alm)
			spin = 1
			alm = np.zeros((2, ainfo.nelem), dtype=ctype)
			alm[0, i] = 1. + 1.j
			alm[1, i] = 2. - 2.j
			omap = np.zeros((2, npix), dtype)
			curvedsky.alm2map_healpix(alm, omap, spin=spin)
			alm_out = np.zeros_like(alm) if use_oalm else None
			alm_out = curvedsky.map2alm_healpix(omap, alm=alm_out, spin=spin, ainfo=ainfo, niter=niter)
			np.testing.assert_array_almost_equal(alm_out, alm)
			alm = np.zeros((3, 2, ainfo.nelem), dtype=ctype)
			for k in range(3):
				alm[k, 0, i] = (2*k+1)*(1. + 1.j)
				alm[k, 1, i] = (2*k+2)*(1. - 1.j)
			omap = np.zeros((3, 2, npix), dtype)
			curvedsky.alm2map_healpix(alm, omap, spin=spin)
			alm_out = np.zeros_like(alm) if use_oalm else None
			alm_out = curvedsky.map2alm_healpix(omap, alm=alm_out, spin=spin, ainfo=ainfo, niter=niter)
			np.testing.assert_array_almost_equal(alm_out, alm)
