"""GPU parity tests of synthesis at arbitrary positions (K8: b2_general_* + b2_alm2leg + b2_fft_* through
pixell_b200.sht.synthesis_general / curvedsky.alm2map_pos; reference curvedsky.py:174-207, 993-1016).
The checker is the oracle's ring synthesis with one single-pixel ring per position (direct evaluation of the
series).  Tolerance: 1e-10 relative to the largest value (the reference's own NUFFT epsilon for float64)."""
import numpy as np, pytest

pytestmark = pytest.mark.gpu

def rel(a, b): return np.abs(a-b).max()/max(np.abs(b).max(), 1e-300)

def rand_alm(ncomp, lmax, seed, spin_lmin=0):
	rng = np.random.default_rng(seed)
	nalm = (lmax+1)*(lmax+2)//2
	alm = rng.standard_normal((ncomp, nalm)) + 1j*rng.standard_normal((ncomp, nalm))
	alm[:, :lmax+1] = alm[:, :lmax+1].real
	if spin_lmin:
		from pixell_b200 import curvedsky
		ai = curvedsky.alm_info(lmax)
		for m in range(min(spin_lmin, lmax+1)):
			for l in range(m, spin_lmin): alm[:, ai.lm2ind(l, m)] = 0
	return alm

def positions(n, seed):
	rng = np.random.default_rng(seed)
	theta = np.arccos(rng.uniform(-1, 1, n)); phi = rng.uniform(0, 2*np.pi, n)
	theta[:4] = [0.0, np.pi, 1e-9, np.pi-1e-7]            # poles and their neighbourhood
	phi[4:8] = [0.0, 2*np.pi-1e-12, 1e-15, np.pi]
	return np.stack([theta, phi], 1)

def direct(alm, loc, spin, lmax, mode="STANDARD"):
	from oracle import sht_oracle as so
	n = len(loc)
	return so.synthesis(alm=alm, theta=loc[:, 0], nphi=np.ones(n, int), phi0=loc[:, 1], ringstart=np.arange(n), spin=spin, lmax=lmax, mode=mode)

@pytest.mark.parametrize("spin,lmax", [(0, 47), (2, 64), (1, 30), (3, 33)])
def test_synthesis_general_matches_direct_sum(spin, lmax):
	from pixell_b200 import sht
	nca = 1 if spin == 0 else 2
	alm = rand_alm(nca, lmax, 10+spin, spin_lmin=spin)
	loc = positions(300, spin)
	want = direct(alm, loc, spin, lmax)
	got = sht.synthesis_general(alm=alm, loc=loc, spin=spin, lmax=lmax)
	assert got.shape == want.shape
	assert rel(got, want) < 1e-10

def test_synthesis_general_deriv1_and_torch():
	import torch
	from pixell_b200 import sht
	lmax = 40
	alm = rand_alm(1, lmax, 3)
	loc = positions(200, 9)[8:]                       # derivative components are direction dependent at the poles
	want = direct(alm, loc, 1, lmax, mode="DERIV1")
	got = sht.synthesis_general(alm=alm, loc=loc, spin=1, lmax=lmax, mode="DERIV1")
	assert rel(got, want) < 1e-10
	tgot = sht.synthesis_general(alm=torch.from_numpy(alm).cuda(), loc=torch.from_numpy(loc).cuda(), spin=1, lmax=lmax, mode="DERIV1")
	assert tgot.is_cuda and rel(tgot.cpu().numpy(), want) < 1e-10

def test_alm2map_pos_interface():
	"""pos = [{dec, ra}, ...] with trailing dimensions, negative ra, T/Q/U with spin [0, 2], deriv"""
	from pixell_b200 import curvedsky as cs
	lmax = 50
	alm = rand_alm(3, lmax, 5, spin_lmin=0)
	ai = cs.alm_info(lmax)
	alm[1:, [ai.lm2ind(0, 0), ai.lm2ind(1, 0), ai.lm2ind(1, 1)]] = 0
	rng = np.random.default_rng(6)
	pos = np.stack([rng.uniform(-np.pi/2, np.pi/2, (7, 9)), rng.uniform(-np.pi, np.pi, (7, 9))])
	got = cs.alm2map_pos(alm, pos)
	assert got.shape == (3, 7, 9)
	loc = np.stack([np.pi/2-pos[0].reshape(-1), pos[1].reshape(-1) % (2*np.pi)], 1)
	want = np.concatenate([direct(alm[:1], loc, 0, lmax), direct(alm[1:], loc, 2, lmax)]).reshape(3, 7, 9)
	assert rel(got, want) < 1e-10
	d = cs.alm2map_pos(alm[0], pos, deriv=True)
	assert d.shape == (2, 7, 9)
	dw = direct(alm[:1], loc, 1, lmax, mode="DERIV1").reshape(2, 7, 9); dw[0] *= -1
	assert rel(d, dw) < 1e-10

def test_alm2map_pos_agrees_with_ring_synthesis_at_scale():
	"""lmax 1500 on the pixel centres of a CAR patch: the non-uniform path against the ring path (K1 + K3)"""
	import torch
	from pixell_b200 import curvedsky as cs, geometry
	lmax = 1500
	alm = rand_alm(3, lmax, 7)
	ai = cs.alm_info(lmax)
	alm[1:, [ai.lm2ind(0, 0), ai.lm2ind(1, 0), ai.lm2ind(1, 1)]] = 0
	lval = np.zeros(ai.nelem)
	for mm in range(lmax+1): lval[ai.lm2ind(np.arange(mm, lmax+1), mm)] = np.arange(mm, lmax+1)
	alm *= 1.0/(lval+10.0)                            # red spectrum, like a sky
	shape, wcs = geometry.fullsky_geometry(res=np.deg2rad(6/60))
	m = cs.alm2map(alm, np.zeros((3,)+shape), spin=[0, 2], wcs=wcs)
	ys = np.array([0, 1, 17, 600, 901, shape[0]-2, shape[0]-1]); xs = np.arange(0, shape[1], 37)
	dec = geometry.dec_of(wcs, ys)[:, None] + 0*xs[None, :]; ra = geometry.ra_of(wcs, xs)[None, :] + 0*ys[:, None]
	got = cs.alm2map_pos(alm, np.stack([dec, ra]))
	want = np.asarray(m)[:, ys][:, :, xs]
	assert rel(got, want) < 1e-10

def direct_adjoint(map, loc, spin, lmax, mode="STANDARD"):
	from oracle import sht_oracle as so
	n = len(loc)
	return so.adjoint_synthesis(map=map, theta=loc[:, 0], nphi=np.ones(n, int), phi0=loc[:, 1], ringstart=np.arange(n), spin=spin, lmax=lmax, mode=mode)

@pytest.mark.parametrize("spin,lmax,mode", [(0, 47, "STANDARD"), (2, 64, "STANDARD"), (1, 30, "STANDARD"), (1, 36, "DERIV1")])
def test_adjoint_synthesis_general_matches_direct_sum(spin, lmax, mode):
	from pixell_b200 import sht
	ncm = 1 if spin == 0 else 2
	loc = positions(300, 20+spin)
	if mode == "DERIV1": loc = loc[8:]
	rng = np.random.default_rng(30+spin)
	m = rng.standard_normal((ncm, len(loc)))
	want = direct_adjoint(m, loc, spin, lmax, mode)
	got = sht.adjoint_synthesis_general(map=m, loc=loc, spin=spin, lmax=lmax, mode=mode)
	assert got.shape == want.shape
	assert rel(got, want) < 1e-10

def test_alm2map_pos_adjointness():
	"""<alm2map_pos(a), v> = <a, alm2map_pos(v, adjoint)> in the zipped real alm basis of the reference's adjointness
	test (tests/test_pixell.py:218-230, 1051-1085): m = 0 real parts, sqrt(2) (re, im) of the m > 0 coefficients"""
	from pixell_b200 import curvedsky as cs
	lmax = 40
	ai = cs.alm_info(lmax)
	alm = rand_alm(3, lmax, 40)
	alm[1:, [ai.lm2ind(0, 0), ai.lm2ind(1, 0), ai.lm2ind(1, 1)]] = 0
	rng = np.random.default_rng(41)
	pos = np.stack([rng.uniform(-np.pi/2, np.pi/2, 250), rng.uniform(-np.pi, np.pi, 250)])
	v = rng.standard_normal((3, 250))
	fwd = cs.alm2map_pos(alm, pos)
	back = cs.alm2map_pos(np.zeros_like(alm), pos, map=v.copy(), adjoint=True)
	lhs = np.sum(fwd*v)
	w = np.full(ai.nelem, 2.0); w[:lmax+1] = 1.0
	rhs = np.sum(w*alm.real*back.real) + np.sum((w*alm.imag*back.imag)[:, lmax+1:])
	assert abs(lhs-rhs) < 1e-10*abs(lhs)

def test_rotate_alm():
	"""zyz Euler rotation of alm (reference curvedsky.py:714-742): z rotations are phases, the inverse angles undo a
	rotation, and the reference's own gal -> equ angles carry the galactic pole to RA 192.86, Dec 27.13 (J2000)"""
	from pixell_b200 import curvedsky as cs
	lmax = 48
	ai = cs.alm_info(lmax)
	alm = rand_alm(2, lmax, 50)
	mval = np.zeros(ai.nelem, int)
	for m in range(lmax+1): mval[ai.lm2ind(np.arange(m, lmax+1), m)] = m
	r = cs.rotate_alm(alm, 0.3, 0, 0.4)
	assert rel(r, alm*np.exp(-1j*mval*0.7)) < 1e-10
	ang = np.array([0.5, 1.1, -2.0])
	fwd = cs.rotate_alm(alm, *ang)
	assert rel(fwd, alm) > 0.1
	assert rel(cs.rotate_alm(fwd, *(-ang[::-1])), alm) < 1e-10
	# a narrow beam at the north pole of the "gal" frame, moved to "equ"
	l = np.arange(lmax+1)
	bl = np.exp(-0.5*l*(l+1)*0.08**2)*np.sqrt((2*l+1)/(4*np.pi))
	pole = np.zeros(ai.nelem, complex); pole[:lmax+1] = bl
	eq = cs.rotate_alm(pole, *cs.euler_angs[("gal", "equ")])
	dec, ra = np.deg2rad(27.12825), np.deg2rad(192.85948)
	peak = cs.alm2map_pos(eq[None], np.array([[dec, dec+0.05, dec-0.05, -dec], [ra, ra, ra+0.05, ra]]))[0]
	top = np.sum(bl*np.sqrt((2*l+1)/(4*np.pi)))
	assert abs(peak[0]-top) < 1e-6*top and peak[1] < peak[0] and peak[2] < peak[0] and peak[3] < 0.01*top

@pytest.mark.parametrize("ang", [(0.3, 1.1, -2.0), (-1.3, 0.4, 0.7), (2.5, 2.9, 0.1), (0.0, 0.5, 0.0)])
def test_rotate_alm_against_the_wigner_oracle(ang):
	"""every Euler angle non-zero, non-axisymmetric alm: the engine's rotate_alm (synthesis at the back-rotated nodes +
	exact analysis) against oracle/rotate_oracle.py (Wigner-D by the explicit sum, itself checked against a convention-free
	quadrature and against the reference's published gal -> equ angles in tests/test_oracle_basic.py); also prof2alm"""
	from pixell_b200 import curvedsky as cs
	from oracle import rotate_oracle as ro
	lmax = 24
	alm = rand_alm(1, lmax, 60)[0]
	want = ro.rotate_alm_wigner(alm, lmax, *ang)
	got = cs.rotate_alm(alm, *ang)
	assert rel(got, want) < 1e-10
	got32 = cs.rotate_alm(alm.astype(np.complex64), *ang)
	assert got32.dtype == np.complex64 and rel(got32, want) < 1e-5

def test_prof2alm_places_a_profile_on_the_sky():
	"""reference pixell/curvedsky.py:558-585: a polar profile analysed at m = 0 and rotated to [ra, dec]"""
	from pixell_b200 import curvedsky as cs
	n = 65                                         # CC grid: lmax = 63
	theta = np.arange(n)*np.pi/(n-1)
	prof = np.exp(-0.5*(theta/0.15)**2)
	ra, dec = 0.7, -0.4
	alm = cs.prof2alm(prof, dir=[ra, dec])
	assert alm.shape == (cs.alm_info(63).nelem,)
	pos = np.array([[dec, dec, dec+0.15, np.pi/2-1e-3], [ra, ra+0.15/np.cos(dec), ra, 0.0]])
	val = cs.alm2map_pos(alm[None], pos)[0]
	assert abs(val[0]-1) < 2e-3 and abs(val[2]-np.exp(-0.5)) < 5e-3 and abs(val[1]-np.exp(-0.5)) < 2e-2 and abs(val[3]) < 1e-3
	flat = cs.prof2alm(prof, norot=True)
	assert flat.shape == (64,) and abs(np.sum(flat.real*np.sqrt((2*np.arange(64)+1)/(4*np.pi)))-1) < 2e-3

def test_golden_lensed_map():
	"""The reference's lensing golden MM_lensed_071123.fits (reference tests/test_pixell.py:351-356 through
	lensing.rand_map -> lens_map_curved, lensing.py:468-492): rand_alm(seed=1) -> phi gradient with alm2map(deriv=True)
	-> offset_by_grad (host geometry, tests/lens_helper.py) -> alm2map_pos at the displaced positions -> rotate_pol.
	Pins the arbitrary-position synthesis (K8) and the DERIV1 path to the reference's own output (which ducc produced
	with its NUFFT at epsilon = 1e-10)."""
	import os
	import lens_helper
	from conftest import GOLDEN
	from pixell_b200 import curvedsky as cs, geometry, enmap
	g = np.load(os.path.join(GOLDEN, "lensed_071123.npz")); u = np.load(os.path.join(GOLDEN, "unlensed_071123.npz"))
	ps = np.load(os.path.join(GOLDEN, "lens_ps_400.npy"))
	alm = cs.rand_alm(ps, lmax=400, seed=1)
	phi_alm, cmb_alm = alm[0], alm[1:]
	wcs = geometry.CarWCS(u["crval"], u["cdelt"], u["crpix"])
	shape = tuple(int(v) for v in u["shape"][1:])
	grad = cs.alm2map(phi_alm, geometry.zeros((2,)+shape, wcs), deriv=True)
	dec = geometry.dec_of(wcs, np.arange(shape[0]))[:, None] + np.zeros(shape)
	ra = geometry.ra_of(wcs, np.arange(shape[1]))[None, :] + np.zeros(shape)
	raw = lens_helper.offset_by_grad(np.array([dec, ra]), np.asarray(grad))
	lensed = cs.alm2map_pos(cmb_alm, raw[:2], spin=[0, 2])
	lensed = enmap.rotate_pol(lensed, raw[2])
	got, want = lensed[:, g["rows"]], g["map"]
	# the reference's own criterion
	assert np.all(np.isclose(got, want))
	# and tighter, away from the pole rows (where the reference's positions are degenerate): T to 2e-9 of its range
	inner = slice(1, -1)
	assert np.abs(got[0, inner]-want[0, inner]).max() < 2e-9*np.abs(want[0]).max()
	assert np.abs(got[1:, inner]-want[1:, inner]).max() < 2e-9*np.abs(want[1:]).max()
	# lensing moves the map: the same comparison against the unlensed golden must fail by a wide margin
	assert np.abs(got[0, inner]-u["map"][0, inner]).max() > 1e-2*np.abs(want[0]).max()

def test_method_general_matches_ring_methods():
	"""alm2map / map2alm with method="general" (reference curvedsky.py:796-820, 875-898, 1088-1120): the non-uniform path at the
	pixel centres gives what the ring methods give, on a full-sky grid and on a cut-sky patch, with adjoints and Jacobi steps"""
	from pixell_b200 import curvedsky as cs, geometry
	lmax = 36
	ai = cs.alm_info(lmax)
	alm = rand_alm(3, lmax, 60); alm[1:, [ai.lm2ind(0, 0), ai.lm2ind(1, 0), ai.lm2ind(1, 1)]] = 0
	fshape, fwcs = geometry.fullsky_geometry(res=np.deg2rad(4.0))
	pshape, pwcs = geometry.slice_geometry(fshape, fwcs, 8, 31, 10, 60)
	rng = np.random.default_rng(61)
	for shape, wcs in ((fshape, fwcs), (pshape, pwcs)):
		want = cs.alm2map(alm, geometry.zeros((3,)+shape, wcs), spin=[0, 2])
		got = cs.alm2map(alm, geometry.zeros((3,)+shape, wcs), spin=[0, 2], method="general")
		assert rel(np.asarray(got), np.asarray(want)) < 1e-10
		d = cs.alm2map(alm[0], geometry.zeros((2,)+shape, wcs), deriv=True, method="general")
		assert rel(np.asarray(d), np.asarray(cs.alm2map(alm[0], geometry.zeros((2,)+shape, wcs), deriv=True))) < 1e-10
		m = geometry.ndmap(rng.standard_normal((3,)+shape), wcs)
		a1 = cs.alm2map_adjoint(m, spin=[0, 2], ainfo=ai, method="general")
		a0 = cs.alm2map_adjoint(m, spin=[0, 2], ainfo=ai)
		assert rel(a1, a0) < 1e-10
		# same weights on both sides: per-pixel areas for "general", the same numbers per ring (north first) for "cyl";
		# Jacobi steps only on the full sky (on a patch "cyl" iterates on zero-padded full rings, "general" on the patch)
		wring = geometry.pixsize_rows(shape, wcs)[::-1]
		for niter in ((0, 2) if shape == fshape else (0,)):
			b1 = cs.map2alm(want, lmax=lmax, spin=[0, 2], method="general", niter=niter)
			b0 = cs.map2alm(want, lmax=lmax, spin=[0, 2], method="cyl", niter=niter, weights=wring)
			assert rel(b1, b0) < 1e-9
	assert cs.calc_locinfo(pshape, pwcs).loc.shape == (pshape[0]*pshape[1], 2)
