"""Self-consistency of the CPU oracle: quadrature weights, exact analysis on minimal grids
(reference tests/test_pixell.py:870-965), adjointness (:1051-1085), and alm helpers against the
reference's own cmisc_core.c when oracle/_ref is built.  CPU only."""
import ctypes, os
import numpy as np, pytest
from oracle import sht_oracle as so, alm_oracle as ao, pixell_ref as pr

GRIDS = ["CC", "F1", "MW", "MWflip", "DH", "F2"]

def rand_alm(lmax, ncomp, seed, spin=0, mmax=None):
	rng = np.random.default_rng(seed)
	ai = ao.AlmInfo(lmax, mmax)
	alm = (rng.standard_normal((ncomp, ai.nelem)) + 1j*rng.standard_normal((ncomp, ai.nelem)))/2**0.5
	alm[:, :lmax+1] = alm[:, :lmax+1].real
	if spin > 0:
		for m in range(ai.mmax+1):
			alm[:, ai.mstart[m]+np.arange(m, min(spin, lmax+1))] = 0
	return alm, ai

@pytest.mark.parametrize("name", GRIDS)
@pytest.mark.parametrize("n", [7, 32, 33])
def test_gridweights_moments(name, n):
	w = so.get_gridweights(name, n)
	th = so.grid_theta(name, n)
	assert abs(w.sum() - 4*np.pi) < 1e-13
	nex = {"DH": n//2, "F2": n}.get(name, n)   # degree of exactness differs for DH (half the nodes carry it)
	for j in range(nex):
		want = 0.0 if j % 2 else 2*np.pi*2/(1-j*j)
		assert abs(np.sum(w*np.cos(j*th)) - want) < 5e-14, (name, n, j)

@pytest.mark.parametrize("name", ["CC", "F1", "MW", "MWflip"])
def test_gridweights_vs_bruteforce(name):
	for n in (6, 9):
		assert np.allclose(so.get_gridweights(name, n), so.gridweights_bruteforce(name, n), atol=1e-12)

@pytest.mark.parametrize("spin", [0, 1, 2])
@pytest.mark.parametrize("name,ny,nx,lmax", [("F1", 32, 61, 30), ("CC", 33, 64, 30), ("MW", 31, 62, 30),
	("MWflip", 31, 62, 30), ("DH", 64, 64, 30), ("F2", 63, 64, 30), ("F1", 80, 70, 30)])
def test_analysis_2d_exact(spin, name, ny, nx, lmax):
	nc = 1 if spin == 0 else 2
	alm, ai = rand_alm(lmax, nc, 5, spin)
	kw = dict(spin=spin, lmax=lmax, mmax=lmax, mstart=ai.mstart, geometry=name, phi0=0.3)
	m = so.synthesis_2d(alm=alm, ntheta=ny, nphi=nx, **kw)
	back = so.analysis_2d(map=m, **kw)
	assert np.abs(back-alm).max() < 2e-13

@pytest.mark.parametrize("spin", [0, 1, 2])
def test_adjointness(spin):
	"""<Y a, m> == <a, Y^T m> with the real inner product on alm (m=0 weight 1, m>0 weight 2)."""
	lmax, ny, nx = 12, 9, 16
	nc = 1 if spin == 0 else 2
	alm, ai = rand_alm(lmax, nc, 7, spin)
	rng = np.random.default_rng(8)
	m = rng.standard_normal((nc, ny, nx))
	kw = dict(spin=spin, lmax=lmax, mmax=lmax, mstart=ai.mstart, geometry="CC", phi0=-0.4)
	def dot_alm(a, b):
		w = np.full(ai.nelem, 2.0); w[:lmax+1] = 1
		return np.sum(w*(a.real*b.real + a.imag*b.imag))
	Ya = so.synthesis_2d(alm=alm, ntheta=ny, nphi=nx, **kw)
	YTm = so.adjoint_synthesis_2d(map=m, **kw)
	assert abs(np.sum(Ya*m) - dot_alm(alm, YTm)) < 1e-11
	Am = so.analysis_2d(map=m, **dict(kw, lmax=7, mmax=7, mstart=ao.AlmInfo(7).mstart))
	a7, ai7 = rand_alm(7, nc, 9, spin)
	ATa = so.adjoint_analysis_2d(alm=a7, ntheta=ny, nphi=nx, **dict(kw, lmax=7, mmax=7, mstart=ai7.mstart))
	w = np.full(ai7.nelem, 2.0); w[:8] = 1
	assert abs(np.sum(ATa*m) - np.sum(w*(a7.real*Am.real + a7.imag*Am.imag))) < 1e-11

def test_deriv1_matches_spin1_gradient():
	lmax, ny, nx = 20, 24, 48
	alm, ai = rand_alm(lmax, 1, 3)
	l = np.concatenate([np.arange(m, lmax+1) for m in range(lmax+1)])
	grad = np.zeros((2, ai.nelem), complex); grad[0] = alm[0]*np.sqrt(l*(l+1.0))
	kw = dict(lmax=lmax, mmax=lmax, mstart=ai.mstart, geometry="F1", ntheta=ny, nphi=nx)
	a = so.synthesis_2d(alm=alm, spin=1, mode="DERIV1", **kw)
	b = so.synthesis_2d(alm=grad, spin=1, **kw)
	assert np.abs(a-b).max() < 1e-12
	# finite-difference check of d/dtheta on a fine grid
	f = so.synthesis_2d(alm=alm, spin=0, **dict(kw, ntheta=2001, geometry="CC"))[0]
	d = so.synthesis_2d(alm=alm, spin=1, mode="DERIV1", **dict(kw, ntheta=2001, geometry="CC"))[0]
	h = np.pi/2000
	fd = (f[2:]-f[:-2])/(2*h)
	assert np.abs(fd-d[1:-1]).max() < 1e-3*np.abs(d).max()

def test_pixell_layer_flips_and_cyl():
	"""alm2map through the pixell-shaped layer on a fullsky F1 map (y south-first, x decreasing)
	equals the ring-level synthesis evaluated at the pixel coordinates."""
	lmax = 16
	geo = pr.fullsky_geo(shape=(18, 36))
	alm, ai = rand_alm(lmax, 3, 11)
	alm[1:, ai.mstart[0]+np.arange(0, 2)] = 0; alm[1:, ai.mstart[1]+1] = 0
	m = np.zeros((3,)+geo.shape)
	pr.alm2map(alm, m, geo, spin=[0, 2])
	theta = np.pi/2 - geo.dec(np.arange(18)); phi = geo.ra(np.arange(36))
	# direct evaluation, ring by ring, one pixel per ring call
	for comp_sl, spin in ((slice(0, 1), 0), (slice(1, 3), 2)):
		leg = so.alm2leg(alm[comp_sl], theta, spin, lmax, lmax, ai.mstart)
		mm = np.arange(lmax+1); wm = np.where(mm == 0, 1.0, 2.0)
		direct = np.einsum("crm,xm->crx", leg*wm, np.exp(1j*np.outer(phi, mm))).real
		assert np.abs(direct - m[comp_sl]).max() < 1e-12
	back = pr.map2alm(m, geo, lmax=lmax, spin=[0, 2])
	assert np.abs(back-alm).max() < 1e-12

# ------------------------------------------------------------- alm helpers vs the reference's C

def _ref():
	lib = ao.ref_cmisc()
	if lib is None: pytest.skip("oracle/_ref/libcmisc_ref.so not built (reference tree absent)")
	return lib

def test_alm2cl_vs_reference_c():
	lib = _ref()
	lmax = 37
	a, ai = rand_alm(lmax, 2, 21)
	a[:, :lmax+1] += 0.3j   # the C code ignores imag(a_l0); make sure we do too
	cl = np.zeros(lmax+1)
	dp = ctypes.POINTER(ctypes.c_double); ip = ctypes.POINTER(ctypes.c_int64)
	lib.alm2cl_dp(ctypes.c_int(lmax), ctypes.c_int(lmax), ai.mstart.ctypes.data_as(ip),
		a[0].view(np.float64).ctypes.data_as(dp), a[1].view(np.float64).ctypes.data_as(dp), cl.ctypes.data_as(dp))
	assert np.allclose(cl, ao.alm2cl(ai, a[0], a[1]), rtol=1e-13, atol=1e-15)

def test_lmul_vs_reference_c():
	lib = _ref()
	lmax = 29
	a, ai = rand_alm(lmax, 3, 22)
	dp = ctypes.POINTER(ctypes.c_double); ip = ctypes.POINTER(ctypes.c_int64)
	for lfmax in (lmax, lmax-5, lmax+4):
		f = np.random.default_rng(3).standard_normal(lfmax+1)
		b = a[0].copy()
		lib.lmul_dp(ctypes.c_int(lmax), ctypes.c_int(lmax), ai.mstart.ctypes.data_as(ip),
			b.view(np.float64).ctypes.data_as(dp), ctypes.c_int(lfmax), f.ctypes.data_as(dp))
		assert np.allclose(b, ao.lmul(ai, a[0], f), rtol=1e-14, atol=0)
		M = np.random.default_rng(4).standard_normal((3, 3, lfmax+1))
		src = a.copy(); dst = np.zeros_like(a)
		pp = (dp*3)(*[src[i].view(np.float64).ctypes.data_as(dp) for i in range(3)])
		po = (dp*3)(*[dst[i].view(np.float64).ctypes.data_as(dp) for i in range(3)])
		pm = (dp*9)(*[M[r, c].ctypes.data_as(dp) for r in range(3) for c in range(3)])
		lib.lmatmul_dp(ctypes.c_int(3), ctypes.c_int(3), ctypes.c_int(lmax), ctypes.c_int(lmax),
			ai.mstart.ctypes.data_as(ip), pp, ctypes.c_int(lfmax), pm, po)
		assert np.allclose(dst, ao.lmul(ai, a, M), rtol=1e-13, atol=1e-15)

def test_transpose_alm_vs_reference_c():
	lib = _ref()
	lmax = 23
	a, ai = rand_alm(lmax, 1, 23)
	out = np.zeros_like(a[0])
	dp = ctypes.POINTER(ctypes.c_double); ip = ctypes.POINTER(ctypes.c_int64)
	lib.transpose_alm_dp(ctypes.c_int(lmax), ctypes.c_int(lmax), ai.mstart.ctypes.data_as(ip),
		a[0].view(np.float64).ctypes.data_as(dp), out.view(np.float64).ctypes.data_as(dp))
	assert np.array_equal(out, ao.transpose_alm(ai, a[0]))

def test_philox_known_answer():
	"""Philox4x32-10 of the numpy restatement against the published known-answer vector (Random123 kat_vectors:
	counter 0, key 0); the device kernel b2_rand_alm is checked against this restatement on the GPU"""
	x = ao.philox4x32_10(np.array([0], np.uint64), 0)
	assert [int(v[0]) for v in x] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
	z = ao.philox_normal_pairs(np.arange(200000, dtype=np.uint64), 42)
	assert abs(z.real.mean()) < 0.01 and abs(z.real.var()-1) < 0.02 and abs(z.imag.var()-1) < 0.02 and abs(np.mean(z.real*z.imag)) < 0.01

@pytest.mark.parametrize("geom,ny,lmax,spin", [("F1", 64, 60, 0), ("CC", 65, 63, 2), ("MW", 181, 150, 1), ("F1", 1024, 1000, 2), ("CC", 1026, 1024, 0)])
def test_tuned_cpu_legendre_matches_the_checker(geom, ny, lmax, spin):
	"""oracle/sht_fast.c (the CPU arm of bench.py: SIMD over ring blocks, ring skipping, scaled prologue) against
	oracle/sht_oracle.c (the checker the CUDA kernels are pinned to), on a subset of m"""
	theta = so.grid_theta(geom, ny)
	mstart = so.default_mstart(lmax, lmax); nalm = (lmax+1)*(lmax+2)//2
	rng = np.random.default_rng(1)
	nc = 1 if spin == 0 else 2
	ms = np.unique(np.concatenate([np.arange(0, lmax+1, max(1, lmax//7)), [lmax]])).astype(np.int32)
	alm = rng.standard_normal((nc, nalm)) + 1j*rng.standard_normal((nc, nalm))
	want = so.alm2leg(alm, theta, spin, lmax, lmax, mstart)[:, :, ms].transpose(0, 2, 1)
	got = so.fast_alm2leg(alm, theta, spin, lmax, lmax, mstart, ms)
	assert np.abs(got-want).max() < 1e-12*np.abs(want).max()
	leg = rng.standard_normal((nc, len(ms), ny)) + 1j*rng.standard_normal((nc, len(ms), ny))
	full = np.zeros((nc, ny, lmax+1), complex); full[:, :, ms] = leg.transpose(0, 2, 1)
	want = so.leg2alm(full, theta, spin, lmax, lmax, mstart, nalm)
	got = so.fast_leg2alm(leg, theta, spin, lmax, lmax, mstart, nalm, ms)
	idx = np.concatenate([mstart[m] + np.arange(max(m, spin), lmax+1) for m in ms])
	assert np.abs(got[:, idx]-want[:, idx]).max() < 1e-12*np.abs(want).max()

def test_rotate_alm_oracles_agree_and_follow_the_published_euler_angles():
	"""oracle/rotate_oracle.py: the Wigner-D rotation against the convention-free brute force (quadrature of the rotated field),
	all three Euler angles non-zero, non-axisymmetric alm; and the convention against pixell/curvedsky.py:714-716: with the
	gal -> equ angles the galactic pole must land on (ra, dec) of the NGP and the celestial pole's galactic direction on z"""
	from oracle import rotate_oracle as ro
	lmax = 9
	rng = np.random.default_rng(8)
	nalm = (lmax+1)*(lmax+2)//2
	alm = rng.standard_normal(nalm) + 1j*rng.standard_normal(nalm); alm[:lmax+1] = alm[:lmax+1].real
	for ang in [(0.3, 1.1, -2.0), (-1.3, 0.4, 0.7), (0.0, 0.5, 0.0), (1.0, 0.0, 0.0)]:
		a = ro.rotate_alm_wigner(alm, lmax, *ang); b = ro.rotate_alm_bruteforce(alm, lmax, *ang)
		assert np.abs(a-b).max() < 1e-12*np.abs(alm).max(), ang
	deg = np.pi/180
	psi, theta, phi = 57.06793215*deg, 62.87115487*deg, -167.14056929*deg
	R = ro.rotmat(psi, theta, phi)
	ngp = R @ np.array([0, 0, 1.0])                      # galactic pole in equatorial coordinates: ra 192.859, dec 27.128
	assert abs(np.arctan2(ngp[1], ngp[0])/deg % 360 - 192.85948) < 1e-3 and abs(np.arcsin(ngp[2])/deg - 27.12825) < 1e-3
	l, b = 122.93192*deg, 27.12825*deg                   # celestial pole in galactic coordinates
	ncp = R @ np.array([np.cos(b)*np.cos(l), np.cos(b)*np.sin(l), np.sin(b)])
	assert np.abs(ncp - np.array([0, 0, 1.0])).max() < 1e-4
	# a pure dipole along z, rotated: a'_1m follows the moved axis
	d = np.zeros(3, complex); d[1] = 1.0                 # lmax = 1: (00, 10, 11)
	out = ro.rotate_alm_wigner(d, 1, 0.0, np.pi/2, 0.0)  # z -> x: f' = Y_10(R^-1 x) ~ x/r = -(Y_11 - Y_1-1)/sqrt 2 -> a_11 = -1/sqrt 2
	assert abs(out[1]) < 1e-14 and abs(out[2] + 2**-0.5) < 1e-14
