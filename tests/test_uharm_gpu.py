"""GPU tests of the Unified Harmonic Transform mirror (pixell_b200.uharm, reference pixell/uharm.py:8-209): the same
filtering code on the flat-sky FFT path and on the curved-sky SHT path, checked against numpy / the oracle."""
import numpy as np, pytest

pytestmark = pytest.mark.gpu

def rel(a, b): return np.abs(a-b).max()/max(np.abs(b).max(), 1e-300)

def test_flat_mode_smoothing_matches_numpy():
	from pixell_b200 import uharm, geometry, enmap
	ny, nx, res = 64, 96, np.deg2rad(2/60)
	d = np.rad2deg(res)
	wcs = geometry.CarWCS(crval=[0, 0], cdelt=[-d, d], crpix=[nx/2+0.5, ny/2+0.5])
	uht = uharm.UHT((ny, nx), wcs)
	assert uht.mode == "flat"
	rng = np.random.default_rng(0)
	m = geometry.ndmap(rng.standard_normal((3, ny, nx)), wcs)
	bl = np.exp(-0.5*np.arange(20000)**2*np.deg2rad(5/60)**2)
	beam = uht.lprof2hprof(bl)
	assert beam.shape == (ny, nx)
	out = uht.harm2map(uht.hmul(beam, uht.map2harm(m)))
	# spin 0 everywhere: plain 2-D filtering; the "phys" factors cancel in the round trip
	want = np.fft.ifft2(np.fft.fft2(m)*np.interp(enmap.modlmap((ny, nx), wcs), np.arange(20000), bl)).real
	assert rel(np.asarray(out), want) < 1e-12
	# "phys" normalisation (Parseval): sum |map2harm|^2 = pixsize * sum map^2
	h = uht.map2harm(m[0])
	ps = uht.harm2powspec(h)
	assert abs(float(np.sum(np.asarray(ps)))/(enmap.pixsize((ny, nx), wcs)*ny*nx) - np.mean(np.asarray(m[0])**2)) < 1e-10
	assert abs(uht.sum_hprof(ps) - float(np.sum(np.asarray(ps)))*uht.nper) < 1e-9*abs(uht.sum_hprof(ps))
	# adjoint pair
	y = rng.standard_normal((ny, nx)) + 1j*rng.standard_normal((ny, nx))
	lhs = np.vdot(np.asarray(uht.harm2map_adjoint(m[0])), y)
	rhs = np.vdot(np.asarray(m[0]), np.asarray(enmap.harm2map(geometry.ndmap(y, wcs), spin=0, normalize="phys", keep_imag=True)))
	assert abs(lhs-rhs) < 1e-10*abs(lhs)

def test_curved_mode_smoothing_matches_alm_filtering():
	from pixell_b200 import uharm, geometry, curvedsky as cs
	shape, wcs = geometry.fullsky_geometry(res=np.deg2rad(1.0))
	lmax = 100
	uht = uharm.UHT(shape, wcs, mode="curved", lmax=lmax)
	rng = np.random.default_rng(1)
	ai = cs.alm_info(lmax)
	alm = rng.standard_normal((3, ai.nelem)) + 1j*rng.standard_normal((3, ai.nelem)); alm[:, :lmax+1] = alm[:, :lmax+1].real
	alm[1:, [0, 1, lmax+1]] = 0
	m = cs.alm2map(alm, geometry.zeros((3,)+shape, wcs), spin=[0, 2])
	bl = np.exp(-0.5*np.arange(lmax+1)**2*np.deg2rad(3.0)**2)
	hp = uht.lprof2hprof(bl[:50])                                   # shorter than lmax: zero padded
	assert hp.shape == (lmax+1,) and np.all(hp[50:] == 0)
	harm = uht.map2harm(m, spin=[0, 2])
	assert rel(harm, alm) < 1e-10
	out = uht.harm2map(uht.hmul(bl, harm), spin=[0, 2])
	want = cs.alm2map(cs.almxfl(alm.copy(), bl), geometry.zeros((3,)+shape, wcs), spin=[0, 2])
	assert rel(np.asarray(out), np.asarray(want)) < 1e-11
	# profiles: a Gaussian beam in r and l
	r = np.linspace(0, np.pi, 4001)
	br = uht.hprof2rprof(bl, r)
	assert rel(uht.rprof2hprof(br, r), bl) < 1e-5
	# power spectrum with the patch correction and the expanded profile
	ps = uht.harm2powspec(harm[0])
	assert rel(ps, cs.alm2cl(alm[0])) < 1e-10
	full = uht.hprof2harm(bl)
	assert full.shape == (ai.nelem,) and full[ai.lm2ind(7, 3)] == bl[7]
	# quadrature weights tie map2harm to the adjoint of harm2map (plain ring weights on 180 rings are exact to degree 179 < 2 lmax,
	# so only approximately here: that gap is what the exact analysis closes)
	W = uht.quad_weights()
	a2 = uht.harm2map_adjoint(geometry.ndmap(np.asarray(m[0])*W, wcs), spin=0)
	assert rel(a2, alm[0]) < 5e-3
