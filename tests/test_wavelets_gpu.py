"""GPU tests of the curved-sky wavelet transform mirror (pixell_b200.wavelets, reference pixell/wavelets.py:206-417)."""
import numpy as np, pytest

pytestmark = pytest.mark.gpu

def rel(a, b): return np.abs(a-b).max()/max(np.abs(b).max(), 1e-300)

def _sky(lmin, lmax, shape, wcs, seed, ncomp=None, lcut=None):
	from pixell_b200 import curvedsky as cs, geometry
	rng = np.random.default_rng(seed)
	ai = cs.alm_info(lmax)
	pre = () if ncomp is None else (ncomp,)
	alm = rng.standard_normal(pre+(ai.nelem,)) + 1j*rng.standard_normal(pre+(ai.nelem,))
	alm[..., :lmax+1] = alm[..., :lmax+1].real
	for m in range(lmax+1):                                  # only lmin <= l <= lmax
		idx = ai.lm2ind(np.arange(m, lmax+1), m)
		alm[..., idx[np.arange(m, lmax+1) < lmin]] = 0
		if lcut is not None: alm[..., idx[np.arange(m, lmax+1) > lcut]] = 0
	return alm, cs.alm2map(alm, geometry.zeros(pre+shape, wcs), spin=[0])

@pytest.mark.parametrize("basis", ["cosine", "butter"])
def test_wavelet_roundtrip_and_scale_content(basis):
	"""perfect reconstruction of a sky band-limited to the basis' range (sum of squared filters = 1), every scale map
	equal to the direct synthesis of the filtered alm on that scale's grid"""
	from pixell_b200 import wavelets, uharm, curvedsky as cs, geometry
	shape, wcs = geometry.fullsky_geometry(res=np.deg2rad(1.0))
	lmax = 160
	if basis == "cosine": b = wavelets.CosineNeedlet([10, 20, 40, 80, 160]); lmin = 10
	else: b = wavelets.ButterTrim(lmin=10, lmax=lmax); lmin = 0
	uht = uharm.UHT(shape, wcs, mode="curved", lmax=lmax)
	wt = wavelets.WaveletTransform(uht, basis=b)
	assert wt.nlevel == b.n and all(g[0][-2] <= shape[0] for g in wt.geometries)
	# the finest scale's grid has lmax rows and carries l <= lmax - 1 exactly (as in the reference): keep the sky below that
	alm, m = _sky(lmin, lmax, shape, wcs, 3, ncomp=2, lcut=lmax-10)
	ls = np.arange(lmax+1.0)
	tot = sum(b(i, ls)**2 for i in range(b.n))
	assert np.allclose(tot[lmin:lmax] if basis == "cosine" else tot, 1.0)
	wave = wt.map2wave(m)
	assert wave.nmap == wt.nlevel and wave.pre == (2,)
	for i, (gs, gw) in enumerate(wt.geometries):
		small = cs.alm_info(int(b.lmaxs[i]))
		a = cs.transfer_alm(cs.alm_info(lmax), alm, small)
		a = cs.almxfl(a, b(i, np.arange(small.lmax+1.0))/wt.norms[i], ainfo=small)
		want = cs.alm2map(a, geometry.zeros((2,)+tuple(gs[-2:]), gw), spin=[0], ainfo=small)
		assert rel(np.asarray(wave.maps[i]), np.asarray(want)) < 1e-10
	back = wt.wave2map(wave)
	assert rel(np.asarray(back), np.asarray(m)) < 1e-9

def test_wavelet_on_band_and_device_tensors():
	import torch
	from pixell_b200 import wavelets, uharm, geometry
	shape, wcs = geometry.band_geometry(np.deg2rad([-40, 30]), res=np.deg2rad(1.0))
	b = wavelets.CosineNeedlet([8, 16, 32, 64])
	uht = uharm.UHT(shape, wcs, mode="curved", lmax=64)
	wt = wavelets.WaveletTransform(uht, basis=b)
	for (gs, gw) in wt.geometries:
		dec = geometry.dec_of(gw, np.array([-0.5, gs[-2]-0.5]))
		assert np.rad2deg(dec.min()) <= -40+1e-9 and np.rad2deg(dec.max()) >= 30-1e-9 and gs[-2] < 181
	rng = np.random.default_rng(4)
	m = rng.standard_normal(shape)
	w_np = wt.map2wave(geometry.ndmap(m, wcs))
	w_t = wt.map2wave(torch.from_numpy(m).cuda())
	for a, b_ in zip(w_np.maps, w_t.maps):
		assert b_.is_cuda and rel(b_.cpu().numpy(), np.asarray(a)) < 1e-12
