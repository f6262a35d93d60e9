"""pixell_b200.wavelets (host-side parts) against the reference's OWN pixell/wavelets.py, imported unmodified through the
scaffolding of tests/refshim.py: wavelet bases (reference wavelets.py:15-161), the variance basis (:168-206), the flat-sky
scale geometries (:463-470) and the pixel-space HaarTransform (:419-456).  The harmonic transforms are compared on the GPU
(tests/test_wavelets_gpu.py)."""
import sys
import numpy as np, pytest
import refshim

pytestmark = pytest.mark.skipif(refshim.reference_root() is None, reason="reference files not staged (scripts/stage_reference.py)")

@pytest.fixture(scope="module")
def ref():
	import importlib
	from oracle import sht_oracle as so
	mods = refshim.install(so, refshim.oracle_cmisc())
	mods["fft"].set_engine("numpy")
	for name in ("multimap", "uharm", "wavelets"): mods[name] = importlib.import_module("pixell."+name)
	return mods

def test_bases_match_the_reference(ref):
	from pixell_b200 import wavelets as W
	R = ref["wavelets"]
	l = np.arange(0, 3000, dtype=np.float64)
	for mine, theirs in [(W.ButterTrim(lmin=7, lmax=2500), R.ButterTrim(lmin=7, lmax=2500)), (W.Butterworth(lmin=7, lmax=2500), R.Butterworth(lmin=7, lmax=2500)),
			(W.ButterTrim(step=1.5, shape=5, trim=3e-2, lmin=20, lmax=900), R.ButterTrim(step=1.5, shape=5, trim=3e-2, lmin=20, lmax=900)),
			(W.DigitalButterTrim(lmin=7, lmax=1200), R.DigitalButterTrim(lmin=7, lmax=1200))]:
		assert mine.n == theirs.n and np.array_equal(mine.lmaxs, theirs.lmaxs)
		for i in range(mine.n):
			ll = l[l < 1190] if isinstance(mine, W.DigitalButterTrim) else l
			with np.errstate(invalid="ignore"): a, b = mine(i, ll), theirs(i, ll)
			assert np.allclose(a, b, rtol=1e-13, atol=1e-13, equal_nan=True), (type(mine).__name__, i)
	peaks = np.array([10, 40, 100, 300, 800])
	a, b = W.CosineNeedlet(peaks), R.CosineNeedlet(peaks)
	for i in range(a.n): assert np.array_equal(a(i, l), b(i, l))
	assert np.array_equal(W.digitize(np.linspace(0, 1, 50)**2), R.digitize(np.linspace(0, 1, 50)**2))

def test_variance_basis_matches_the_reference(ref):
	from pixell_b200 import wavelets as W
	R = ref["wavelets"]
	mine, theirs = W.Butterworth(lmin=10, lmax=2000).get_variance_basis(), R.Butterworth(lmin=10, lmax=2000).get_variance_basis()
	l = np.geomspace(1, 5000, 300)
	assert mine.n == theirs.n
	for i in range(mine.n): assert np.allclose(mine(i, l), theirs(i, l), rtol=1e-12, atol=1e-14)

def test_flat_scale_geometries_match_the_reference(ref):
	from pixell_b200 import wavelets as W, geometry
	R, enmap = ref["wavelets"], ref["enmap"]
	shape, wcs = enmap.geometry(pos=(0, 0), shape=(200, 300), res=np.deg2rad(0.05))
	mywcs = geometry.CarWCS(wcs.wcs.crval, wcs.wcs.cdelt, wcs.wcs.crpix)
	for ores in np.deg2rad([0.05, 0.11, 0.4, 1.3]):
		s1, w1 = W.make_wavelet_geometry_flat(shape, mywcs, np.deg2rad(0.05), ores)
		s2, w2 = R.make_wavelet_geometry_flat(shape, wcs, np.deg2rad(0.05), ores)
		assert tuple(s1) == tuple(int(v) for v in s2)
		assert np.allclose(w1.wcs.cdelt, w2.wcs.cdelt, rtol=1e-14) and np.allclose(w1.wcs.crpix, w2.wcs.crpix, rtol=1e-13) and np.allclose(w1.wcs.crval, w2.wcs.crval)

@pytest.mark.parametrize("shape,ref_point", [((64, 80), [0, 0]), ((61, 83), [0.01, -0.02]), ((3, 50, 47), None)])
def test_haar_transform_matches_the_reference(ref, shape, ref_point):
	from pixell_b200 import wavelets as W, geometry
	R, enmap = ref["wavelets"], ref["enmap"]
	gshape, wcs = enmap.geometry(pos=(0.02, -0.03), shape=shape[-2:], res=np.deg2rad(0.1))
	rng = np.random.default_rng(1)
	data = rng.standard_normal(shape)
	mywcs = geometry.CarWCS(wcs.wcs.crval, wcs.wcs.cdelt, wcs.wcs.crpix)
	a = W.HaarTransform(3, ref=ref_point).map2wave(geometry.ndmap(data, mywcs))
	b = R.HaarTransform(3, ref=ref_point).map2wave(enmap.enmap(data, wcs))
	assert a.nmap == b.nmap
	for x, y in zip(a.maps, b.maps):
		assert x.shape == y.shape and np.allclose(np.asarray(x), np.asarray(y), rtol=0, atol=1e-14)
		assert np.allclose(x.wcs.wcs.crpix, y.wcs.wcs.crpix, atol=1e-9) and np.allclose(x.wcs.wcs.cdelt, y.wcs.wcs.cdelt)
	assert np.allclose(np.asarray(W.HaarTransform(3, ref=ref_point).wave2map(a)), data, atol=1e-13)
