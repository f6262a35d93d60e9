"""Parity of the CUDA path against the CPU oracle AT THE BASELINE SIZES (BASELINE.json configs C2 and C3), not through
the round trip: the oracle runs every k-th m (all l, all rings), the engine runs everything (oracle/parity.py).
Tolerance: 1e-10 relative (north_star), float32 1e-5.  Restates the call sequences of the reference's
pixell/curvedsky.py:900-962 (alm2map_raw_2d) and :1018-1046 (map2alm_raw_2d)."""
import numpy as np, pytest
pytestmark = pytest.mark.gpu
from oracle import parity

TOL = 1e-10

@pytest.mark.parametrize("name,ny,nx,lmax,spin,mstride", [
	("C2", 4608, 9216, 4096, 0, 256),
	("C3", 8192, 16384, 8000, 0, 500),
	("C3", 8192, 16384, 8000, 2, 500),
])
def test_legendre_comb_at_baseline_size(name, ny, nx, lmax, spin, mstride):
	r = parity.comb_legendre("F1", ny, nx, lmax, spin, mstride, seed=spin)
	assert r["m_sampled"] >= 16
	assert r["alm2leg"] < TOL and r["leg2alm"] < TOL, r

@pytest.mark.parametrize("name,ny,nx,lmax,spin,mstride", [
	("C2", 4608, 9216, 4096, 0, 256),
	("C3", 8192, 16384, 8000, 0, 500),
	("C3", 8192, 16384, 8000, 2, 500),
])
def test_transform_comb_at_baseline_size(name, ny, nx, lmax, spin, mstride):
	"""synthesis_2d on every pixel and the exact analysis_2d on every l,m, alm supported on the comb"""
	r = parity.comb_transform("F1", ny, nx, lmax, spin, mstride, seed=3+spin, phi0=0.3)
	assert r["synthesis"] < TOL and r["analysis"] < TOL, r

def test_transform_comb_float32():
	r = parity.comb_transform("F1", 4608, 9216, 4096, 0, 256, seed=7, dtype=np.float32)
	assert r["synthesis"] < 1e-5 and r["analysis"] < 1e-5, r

def test_cc_grid_comb_lmax_2700():
	"""a Clenshaw-Curtis grid (pixell's res=4' full sky: 2701 x 5400) at its maximal lmax"""
	r = parity.comb_transform("CC", 2701, 5400, 2699, 2, 180, seed=9)
	assert r["synthesis"] < TOL and r["analysis"] < TOL, r
