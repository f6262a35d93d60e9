"""CPU-only checks: the C-ABI library loads and exports every symbol include/b200sht.h declares,
and the host-side mirror of pixell's curvedsky interface (geometry analysis, alm_info, spin groups,
error contract) agrees with the oracle's restatement of the reference.  No kernel is launched."""
import os, re, ctypes
import numpy as np, pytest
from conftest import ROOT

def header_symbols():
	txt = open(os.path.join(ROOT, "include", "b200sht.h")).read()
	txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
	return sorted(set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", txt)))

def test_library_exports_header_symbols():
	from pixell_b200 import _lib
	lib = ctypes.CDLL(_lib.LIBPATH)
	syms = header_symbols()
	assert len(syms) >= 25
	for s in syms: assert hasattr(lib, s), "missing symbol %s" % s
	# the ctypes prototypes cover the header too
	assert set(syms) <= set(_lib._PROTOS), set(syms) - set(_lib._PROTOS)

def test_no_oracle_in_product():
	"""the product package must never import the checker"""
	for root, _, files in os.walk(os.path.join(ROOT, "pixell_b200")):
		for f in files:
			if f.endswith((".py", ".cu", ".cuh")):
				txt = open(os.path.join(root, f)).read()
				assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt and "libshtoracle" not in txt, f

def test_missing_library_fails_loudly(monkeypatch, tmp_path):
	from pixell_b200 import _lib
	monkeypatch.setattr(_lib, "_lib", None)
	monkeypatch.setattr(_lib, "LIBPATH", str(tmp_path/"nope.so"))
	with pytest.raises(ImportError): _lib.lib()

def test_alm_info():
	from pixell_b200 import curvedsky as cs
	from oracle import alm_oracle as ao
	for kw in (dict(lmax=10), dict(lmax=10, mmax=4), dict(nalm=66), dict(lmax=7, layout="rect"), dict(lmax=5, stride=2)):
		a, b = cs.alm_info(**kw), ao.AlmInfo(**kw)
		assert (a.lmax, a.mmax, a.nelem) == (b.lmax, b.mmax, b.nelem)
		assert np.array_equal(a.mstart.astype(np.int64), b.mstart)
	with pytest.raises(AssertionError): cs.alm_info(lmax=10, mmax=4, nalm=66)   # reference tests/test_pixell.py:760-824
	assert cs.alm_info(lmax=10).lm2ind(3, 2) == 2*(21-2)//2 + 3

def test_spin_helper():
	from pixell_b200.curvedsky import spin_helper
	assert list(spin_helper([0, 2], 3)) == [(0, 0, 1), (2, 1, 3)]
	assert list(spin_helper([0, 1, 2], 5)) == [(0, 0, 1), (1, 1, 3), (2, 3, 5)]
	assert list(spin_helper(0, 2)) == [(0, 0, 1), (0, 1, 2)]
	with pytest.raises(IndexError): list(spin_helper([0, 2], 2))

def test_analyse_geometry_matches_reference_restatement():
	from pixell_b200 import curvedsky as cs, geometry as g
	from oracle import pixell_ref as pr
	cases = []
	for variant in ("fejer1", "cc"):
		shape, wcs = g.fullsky_geometry(res=np.deg2rad(1.0), variant=variant); cases.append((shape, wcs))
		shape, wcs = g.fullsky_geometry(shape=(36+(variant == "cc"), 72), variant=variant); cases.append((shape, wcs))
		cases.append(g.band_geometry(np.deg2rad([-20, 35]), res=np.deg2rad(0.5), variant=variant))
		s, w = g.fullsky_geometry(res=np.deg2rad(1.0), variant=variant); cases.append(g.slice_geometry(s, w, 20, 100, 10, 200))
	cases.append(((10, 20), g.CarWCS([0.3, 0], [-0.7, 0.7], [10, 5])))            # 360/0.7 not an integer -> general
	cases.append(((30, 360), g.CarWCS([0.5, 0], [1.0, -1.0], [180.5, 15.2])))      # increasing ra, north-first, odd offset -> cyl
	for shape, wcs in cases:
		mine = cs.analyse_geometry(shape, wcs)
		ref = pr.analyse_geometry(pr.Geo(shape, wcs.wcs.crval, wcs.wcs.cdelt, wcs.wcs.crpix))
		assert mine.case == ref["case"], (shape, wcs)
		if mine.case == "general": continue
		assert list(mine.flip) == list(ref["flip"]) and tuple(mine.ypad) == tuple(ref["ypad"]) and tuple(mine.xpad) == tuple(ref["xpad"])
		assert abs(mine.phi0-ref["phi0"]) < 1e-14
		if ref["ducc_geo"] is None: assert mine.ducc_geo is None
		else:
			for k in ("name", "nx", "ny", "yoff", "lmax"): assert mine.ducc_geo[k] == ref["ducc_geo"][k]
		assert cs.get_method(shape, wcs) == pr.get_method(ref)

def test_fullsky_geometry_rules():
	"""enmap.fullsky_geometry (pixell/enmap.py:1713-1740)"""
	from pixell_b200 import geometry as g, curvedsky as cs
	shape, wcs = g.fullsky_geometry(res=np.deg2rad(10/60), variant="cc")
	assert shape == (1081, 2160)
	assert cs.analyse_geometry(shape, wcs).ducc_geo.name == "CC"
	shape, wcs = g.fullsky_geometry(shape=(8192, 16384))
	mi = cs.analyse_geometry(shape, wcs)
	assert mi.case == "2d" and mi.ducc_geo.name == "F1" and mi.ducc_geo.lmax == 8191 and list(mi.flip) == [True, True]
	# golden FITS header of the reference's unlensed map: CC 181 x 360
	mi = cs.analyse_geometry((181, 360), g.CarWCS([0.5, 0], [-1, 1], [180.5, 91]))
	assert mi.case == "2d" and mi.ducc_geo.name == "CC"

def test_prepare_alm_contract():
	from pixell_b200 import curvedsky as cs
	with pytest.raises(ValueError): cs.prepare_alm()
	alm, ai = cs.prepare_alm(lmax=5, pre=(3,))
	assert alm.shape == (3, 21) and alm.dtype == np.complex128
	with pytest.raises(ValueError): cs.prepare_alm(alm=np.zeros(21, np.complex64), dtype=np.float64)    # tests/test_pixell.py:1028-1046
	alm, ai = cs.prepare_alm(alm=np.zeros(21, np.complex64), dtype=np.float64, convert=True)
	assert alm.dtype == np.complex128

def test_sym_expand_and_multi_pow():
	from pixell_b200 import curvedsky as cs
	from oracle import alm_oracle as ao
	ps = np.random.default_rng(0).standard_normal((6, 9))
	full = cs.sym_expand(ps)
	assert full.shape == (3, 3, 9) and np.array_equal(full[0, 1], ps[3]) and np.array_equal(full[1, 0], ps[3]) and np.array_equal(full[0, 2], ps[5])
	cov = np.einsum("ikl,jkl->ijl", full, full)
	assert np.allclose(cs.multi_pow_half(cov), ao.eigpow_half(cov))

def test_prepare_alm_mmax():
	"""reference tests/test_pixell.py:760-826 (test_prepare_alm_mmax), same six cases"""
	from pixell_b200 import curvedsky
	lmax, nalm = 3, 10
	alm_in = np.arange(nalm, dtype=np.complex128)
	ainfo_in = curvedsky.alm_info(lmax=3, mmax=3, nalm=nalm, stride=1, layout="triangular")
	def same(a, b): return (a.lmax, a.mmax, a.nelem) == (b.lmax, b.mmax, b.nelem)
	alm_out, ainfo_out = curvedsky.prepare_alm(alm=alm_in, ainfo=None)                      # 1: only alm
	np.testing.assert_array_almost_equal(alm_out, alm_in); assert same(ainfo_out, ainfo_in)
	alm_out, ainfo_out = curvedsky.prepare_alm(alm=None, ainfo=ainfo_in)                    # 2: only alm_info -> zeros
	np.testing.assert_array_almost_equal(alm_out, alm_in*0); assert same(ainfo_out, ainfo_in)
	alm_out, ainfo_out = curvedsky.prepare_alm(alm=alm_in, ainfo=ainfo_in)                  # 3: both
	np.testing.assert_array_almost_equal(alm_out, alm_in); assert same(ainfo_out, ainfo_in)
	with pytest.raises(AssertionError):                                                     # 4: lmax=3, mmax=1 alm alone
		curvedsky.prepare_alm(alm=np.arange(7, dtype=np.complex128), ainfo=None, lmax=lmax)
	ainfo1 = curvedsky.alm_info(lmax=3, mmax=1, nalm=7, stride=1, layout="triangular")
	alm_out, ainfo_out = curvedsky.prepare_alm(alm=None, ainfo=ainfo1)                      # 5: only alm_info, mmax < lmax
	np.testing.assert_array_almost_equal(alm_out, np.zeros(7, np.complex128)); assert same(ainfo_out, ainfo1)
	alm7 = np.arange(7, dtype=np.complex128)
	alm_out, ainfo_out = curvedsky.prepare_alm(alm=alm7, ainfo=ainfo1)                      # 6: both, mmax < lmax
	np.testing.assert_array_almost_equal(alm_out, alm7); assert same(ainfo_out, ainfo1)
	with pytest.raises(ValueError): curvedsky.prepare_alm()
	with pytest.raises(ValueError): curvedsky.prepare_alm(alm=alm7.astype(np.complex64), ainfo=ainfo1)      # dtype contract (:1424)
