"""Pins the CPU oracle (oracle/) against the reference's own known-answer fixtures
(SURVEY.md 8c): tests/golden/*.npz were extracted from /root/reference/tests/data by
tests/golden/make_golden.py.  CPU only."""
import os
import numpy as np
from conftest import GOLDEN
from oracle import sht_oracle as so, alm_oracle as ao, pixell_ref as pr

def test_unlensed_fits_alm2map_spin02():
	"""reference tests/test_pixell.py:351-360 (test_lensing, unlensed map): rand_alm(seed=1) ->
	alm2map(spin=[0,2]) on the CC 181x360 grid at lmax=400."""
	g = np.load(os.path.join(GOLDEN, "unlensed_071123.npz"))
	ps = np.load(os.path.join(GOLDEN, "lens_ps_400.npy"))
	alm = ao.rand_alm(ps, 400, 1)
	geo = pr.Geo(tuple(g["shape"]), g["crval"], g["cdelt"], g["crpix"])
	m = np.zeros((3,)+geo.shape)
	pr.alm2map(alm[1:], m, geo, spin=[0,2])
	got, want = m[:, g["rows"]], g["map"]
	# the reference test uses np.isclose defaults (rtol 1e-5, atol 1e-8); we are far tighter
	assert np.abs(got[0]-want[0]).max() < 5e-10          # |T| <= 408
	assert np.abs(got[1:]-want[1:]).max() < 5e-11        # |Q,U| <= 7.7, pole rows included
	assert np.allclose(got, want, rtol=1e-9, atol=1e-10)

def test_pixels_pkl_rand_map_scalar():
	"""reference tests/test_pixell.py:568-580 (test_pixels, fullsky_10arc_car): rand_map(seed=10,
	lmax=1500) through healpy.synalm's scalar stream + alm2map on CC 1081x2160; 36 pixels."""
	g = np.load(os.path.join(GOLDEN, "pixels_041121.npz"))
	lmax = 1500
	geo = pr.fullsky_geo(res=10/60*np.pi/180, variant="cc")
	assert geo.shape == (1081, 2160)
	l = np.arange(lmax+1)
	# spectra exactly as reference tests/test_pixell.py:120-127 builds them from tests/tests.yml
	am = (np.pi/180/60)**2
	spectra = {"white_10": np.full(lmax+1, 10.0**2*am)}
	dl = np.zeros(lmax+1); dl[2:] = 3.0**2*am*2*np.pi/(l[2:]*(l[2:]+1.0))
	spectra["constant_dl_1"] = dl
	off = np.array([0, 1, 29, 30, 58, 59])
	rows = 510 + off; cols = (2130 + off) % 2160
	for name, cl in spectra.items():
		want = g[name]
		alm = ao.rand_alm_healpy_scalar(cl, lmax, 10)
		m = np.zeros((1,)+geo.shape)
		pr.alm2map(alm[None], m, geo, spin=[0])
		got = m[0][np.ix_(rows, cols)]
		if want.shape != got.shape: want = want.reshape(got.shape)
		assert np.allclose(got, want, rtol=1e-9, atol=1e-10), name
		cut = m[0][510:570][:, (2130+np.arange(60)) % 2160]   # the cut_span_180_2 extract
		# upstream's "meansquare" is literally np.mean(arr*2.) (tests/test_pixell.py:94-95)
		assert np.isclose(np.mean(cut*2.), g[name+"_ms"], rtol=1e-9, atol=1e-12), name

def test_offset_by_grad_golden():
	"""the numpy restatement of lensing.offset_by_grad (tests/lens_helper.py) against the reference's own offset
	goldens MM_offset_{obs_pos,grad,raw_pos}_071123.fits (reference tests/test_pixell.py:333-349)"""
	import lens_helper
	g = np.load(os.path.join(os.path.dirname(__file__), "golden", "offset_071123.npz"))
	got = lens_helper.offset_by_grad(g["obs_pos"], g["grad"])
	want = g["raw_pos"]
	assert np.all(np.isclose(got[:2], want[:2], rtol=1e-12, atol=1e-13))
	ok = np.isfinite(want[2])
	assert np.all(np.isclose(got[2][ok], want[2][ok], rtol=1e-9, atol=1e-11))

def test_lensed_golden_through_the_oracle():
	"""MM_lensed_071123.fits (reference tests/test_pixell.py:351-356; lensing.py:468-492) replayed on the CPU: the oracle's
	DERIV1 synthesis for the gradient of phi, the pinned offset_by_grad helper, and the oracle's direct evaluation at the
	displaced positions.  Every 8th row of the grid keeps this to a few seconds."""
	import lens_helper
	g = np.load(os.path.join(GOLDEN, "lensed_071123.npz")); u = np.load(os.path.join(GOLDEN, "unlensed_071123.npz"))
	ps = np.load(os.path.join(GOLDEN, "lens_ps_400.npy"))
	alm = ao.rand_alm(ps, 400, 1)
	ny, nx = (int(v) for v in u["shape"][1:])
	sel = np.arange(2, len(g["rows"])-1, 4)                    # rows of the fixture (which holds every 2nd map row), poles excluded
	rows = g["rows"][sel]
	dec = np.deg2rad(u["crval"][1] + (rows+1-u["crpix"][1])*u["cdelt"][1])
	ra = np.deg2rad(u["crval"][0] + (np.arange(nx)+1-u["crpix"][0])*u["cdelt"][0])
	pos = np.array([dec[:, None]+0*ra[None, :], ra[None, :]+0*dec[:, None]])
	loc = np.stack([np.pi/2-pos[0].reshape(-1), pos[1].reshape(-1)], 1)
	grad = so.synthesis_general(alm=alm[:1], loc=loc, spin=1, lmax=400, mode="DERIV1").reshape(2, len(rows), nx)
	grad[0] *= -1                                              # theta derivative -> dec derivative (curvedsky.py:918-920)
	raw = lens_helper.offset_by_grad(pos, grad)
	rloc = np.stack([np.pi/2-raw[0].reshape(-1), raw[1].reshape(-1)], 1)
	t = so.synthesis_general(alm=alm[1:2], loc=rloc, spin=0, lmax=400)
	qu = so.synthesis_general(alm=alm[2:4], loc=rloc, spin=2, lmax=400)
	c, s2 = np.cos(2*raw[2].reshape(-1)), np.sin(2*raw[2].reshape(-1))
	got = np.array([t[0], c*qu[0]-s2*qu[1], s2*qu[0]+c*qu[1]]).reshape(3, len(rows), nx)
	want = g["map"][:, sel]
	assert np.all(np.isclose(got, want))
	assert np.abs(got-want).max() < 2e-9*np.abs(want[0]).max()
