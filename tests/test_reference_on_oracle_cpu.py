"""The reference's OWN tests (tests/test_pixell.py, unmodified) run against the reference's OWN curvedsky.py / enmap.py
(unmodified, staged by scripts/stage_reference.py) with the CPU oracle standing in for ducc0 and pixell.cmisc: this pins the
oracle to the reference's round-trip, adjointness (all five geometries, the gnomonic patch through the arbitrary-position
synthesis included) and error-contract tests, and proves the scaffolding the GPU run (test_reference_on_engine_gpu.py) uses.
Skipped when the reference files are not staged (nothing here reads /root/reference on the GPU box)."""
import pytest
import refshim

pytestmark = pytest.mark.skipif(refshim.reference_root() is None, reason="reference files not staged (scripts/stage_reference.py)")

@pytest.fixture(scope="module")
def ref():
	from oracle import sht_oracle as so
	mods = refshim.install(so, refshim.oracle_cmisc())
	mods["fft"].set_engine("numpy")      # the stand-in ducc0 has no ducc0.fft; the GPU run registers the b200 engine here
	return mods

# reference tests/test_pixell.py:760-824, 850-868, 870-965, 967-1026, 1028-1046, 1051-1085, 320-337
@pytest.mark.parametrize("name", ["test_prepare_alm_mmax", "test_almxfl", "test_alm2map_2d_roundtrip", "test_alm2map_healpix_roundtrip",
	"test_alm_conversion", "test_adjointness"])
def test_reference_sht_tests_on_the_oracle(ref, name):
	r = refshim.run_reference_tests([name])
	assert r.testsRun == 1 and r.wasSuccessful(), (r.failures + r.errors)[0][1][-2000:]

# reference tests/test_pixell.py:373-540 on the reference's numpy FFT engine: checks the WCS stand-in under enmap.fft / map2harm
@pytest.mark.parametrize("name", ["test_fft", "test_fft_input_shape", "test_ifft_input_shape", "test_queb_rotmat_complex", "test_queb_rotmat_real"])
def test_reference_fft_tests_with_the_stand_in_wcs(ref, name):
	r = refshim.run_reference_tests([name])
	assert r.testsRun == 1 and r.wasSuccessful(), (r.failures + r.errors)[0][1][-2000:]
