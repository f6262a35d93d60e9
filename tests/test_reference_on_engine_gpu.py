"""SURVEY.md 8b, last row: the reference's UNMODIFIED curvedsky.py / enmap.py / fft.py (staged byte for byte by
scripts/stage_reference.py) imported with pixell_b200.sht registered as `ducc0`, pixell_b200.cmisc as `pixell.cmisc` and
pixell_b200.fft registered in pixell.fft.engines -- the integration INTEGRATION.md describes -- and the reference's own test
functions (tests/test_pixell.py) run on top of the CUDA engine, as they are."""
import numpy as np, pytest
import refshim

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(refshim.reference_root() is None, reason="reference files not staged (scripts/stage_reference.py)")]

@pytest.fixture(scope="module")
def ref():
	from pixell_b200 import sht, cmisc, fft as b2fft
	mods = refshim.install(sht, cmisc)
	b2fft.register(mods["fft"])                    # pixell.fft.engines["b200"] = engine; set_engine("b200")
	return mods

# reference tests/test_pixell.py:760-824, 850-868, 870-965, 967-1026, 1028-1046, 1051-1085
@pytest.mark.parametrize("name", ["test_prepare_alm_mmax", "test_almxfl", "test_alm2map_2d_roundtrip", "test_alm2map_healpix_roundtrip",
	"test_alm_conversion", "test_adjointness"])
def test_reference_sht_tests_on_the_engine(ref, name):
	r = refshim.run_reference_tests([name])
	assert r.testsRun == 1 and r.wasSuccessful(), (r.failures + r.errors)[0][1][-2000:]

# reference tests/test_pixell.py:373-540 through pixell.fft with the b200 engine selected
@pytest.mark.parametrize("name", ["test_fft", "test_fft_input_shape", "test_ifft_input_shape", "test_queb_rotmat_complex", "test_queb_rotmat_real"])
def test_reference_fft_tests_on_the_engine(ref, name):
	assert ref["fft"].engine == "b200"
	r = refshim.run_reference_tests([name])
	assert r.testsRun == 1 and r.wasSuccessful(), (r.failures + r.errors)[0][1][-2000:]

def test_reference_curvedsky_matches_the_mirror(ref):
	"""the reference's curvedsky on the engine and this repository's mirror of it give the same numbers (full-sky and cut-sky
	CAR, T,Q,U): what a user switching from `pixell.curvedsky` to `pixell_b200.curvedsky` would see"""
	from pixell_b200 import curvedsky as mine, geometry
	cs, enmap = ref["curvedsky"], ref["enmap"]
	lmax = 60
	rng = np.random.default_rng(0)
	for variant in ("fejer1", "cc"):
		shape, wcs = enmap.fullsky_geometry(res=np.deg2rad(2.0), variant=variant)
		ainfo = cs.alm_info(lmax)
		alm = rng.standard_normal((3, ainfo.nelem)) + 1j*rng.standard_normal((3, ainfo.nelem))
		alm[:, :lmax+1] = alm[:, :lmax+1].real; alm[1:, [0, 1, lmax+1]] = 0
		want = cs.alm2map(alm, enmap.zeros((3,)+shape, wcs), spin=[0, 2])
		myshape, mywcs = geometry.fullsky_geometry(res=np.deg2rad(2.0), variant=variant)
		assert tuple(myshape) == tuple(shape)
		got = mine.alm2map(alm, geometry.zeros((3,)+myshape, mywcs), spin=[0, 2])
		assert np.abs(np.asarray(got)-np.asarray(want)).max() < 1e-11*np.abs(want).max()
		back_ref = cs.map2alm(want, lmax=lmax, spin=[0, 2]); back = mine.map2alm(got, lmax=lmax, spin=[0, 2])
		assert np.abs(back-back_ref).max() < 1e-11*np.abs(back_ref).max()
		# a declination band (method "cyl")
		sub = want[:, 20:-25]
		myw = geometry.slice_geometry(myshape, mywcs, 20, myshape[-2]-25)[1]
		a1 = cs.map2alm(sub, lmax=lmax, spin=[0, 2]); a2 = mine.map2alm(geometry.ndmap(np.ascontiguousarray(got[:, 20:-25]), myw), lmax=lmax, spin=[0, 2])
		assert np.abs(a1-a2).max() < 1e-11*np.abs(a1).max()
