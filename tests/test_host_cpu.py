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

def test_healpix_ring_info():
	"""get_ring_info_healpix (reference curvedsky.py:1192-1222): ring counts, pixel offsets, HEALPix ring formulas"""
	from pixell_b200 import curvedsky as cs
	for nside in (1, 2, 8, 64):
		ri = cs.get_ring_info_healpix(nside)
		assert ri.nrow == 4*nside-1 and ri.npix == 12*nside**2
		assert int(ri.offsets[-1]+ri.nphi[-1]) == 12*nside**2 and np.all(np.diff(ri.offsets.astype(np.int64)) == ri.nphi[:-1].astype(np.int64))
		assert np.all(np.diff(ri.theta) > 0) and np.allclose(ri.theta + ri.theta[::-1], np.pi)
		assert np.all(ri.nphi == ri.nphi[::-1]) and ri.nphi.max() == 4*nside and ri.nphi[0] == 4
	ri = cs.get_ring_info_healpix(4)
	# z = 1 - i^2/(3 nside^2) in the caps, z = 4/3 - 2i/(3 nside) in the belt; first pixel at pi/(4i) or 0 / pi/(4 nside)
	assert np.allclose(np.cos(ri.theta[:3]), 1-np.arange(1, 4)**2/48.0)
	assert np.allclose(np.cos(ri.theta[3:8]), 4/3.0-2*np.arange(4, 9)/12.0)
	assert np.allclose(ri.phi0[:3], np.pi/(4*np.arange(1, 4))) and np.allclose(ri.phi0[3:7], [np.pi/16, 0, np.pi/16, 0])
	sub = cs.apply_minfo_theta_lim(ri, 1.0, 2.0)
	assert np.all((sub.theta >= 1.0) & (sub.theta <= 2.0)) and len(sub.theta) < ri.nrow
	rr = cs.get_ring_info_radial([0.1, 0.2, 0.5])
	assert np.all(rr.nphi == 1) and list(rr.offsets) == [0, 1, 2]

def test_general_synthesis_host_helpers():
	"""grid sizes and kernel deconvolution factors of the arbitrary-position synthesis (sht.synthesis_general)"""
	from pixell_b200 import sht
	for n in (2, 17, 100, 4098, 16002, 32004):
		f = sht._fast_len(n)
		assert f >= n and f % 2 == 0
		k = f
		for p in (2, 3, 5):
			while k % p == 0: k //= p
		assert k == 1
	lmax, M = 64, sht._fast_len(2*(2*64+2))
	corr = sht._kernel_corr(lmax, M)
	assert corr.shape == (lmax+1,) and np.all(corr > 0) and np.all(np.diff(corr) > 0)      # the kernel transform decays with k
	# P_0 is the integral of the kernel over its support: check against a plain Riemann sum
	z = np.linspace(-1, 1, 200001)
	P0 = 0.5*sht.GENERAL_W*np.trapezoid(np.exp(sht.GENERAL_BETA*(np.sqrt(1-z*z)-1)), z)
	assert abs(1/corr[0]-P0) < 1e-8*P0

def test_flat_sky_host_helpers():
	"""enmap mirrors that need no device: spin_helper, queb_rotmat / map_mul, calc_window, reproject helpers"""
	from pixell_b200 import enmap, reproject, geometry
	assert list(enmap.spin_helper([0, 2], 3)) == [(0, 0, 1), (2, 1, 3)]
	assert list(enmap.spin_helper([0, 2], 6)) == [(0, 0, 1), (2, 1, 3), (0, 3, 4), (2, 4, 6)]
	with pytest.raises(IndexError): list(enmap.spin_helper([0, 2], 2))
	ny, nx = 6, 8
	wcs = geometry.CarWCS(crval=[0, 0], cdelt=[-0.5, 0.5], crpix=[nx/2+0.5, ny/2+0.5])
	lm = enmap.lmap((ny, nx), wcs)
	rot = enmap.queb_rotmat(lm)
	inv = enmap.queb_rotmat(lm, inverse=True)
	assert rot.shape == (2, 2, ny, nx)
	v = np.random.default_rng(0).standard_normal((2, ny, nx))
	assert np.allclose(enmap.map_mul(inv, enmap.map_mul(rot, v)), v)
	assert np.allclose(enmap.queb_rotmat(lm, iau=True), inv)
	wy, wx = enmap.calc_window((ny, nx))
	assert wy[0] == 1 and wx[0] == 1 and np.allclose(wy, np.sinc(np.fft.fftfreq(ny)))
	e = reproject.rot2euler("gal,cel")
	assert np.allclose(e, np.deg2rad([57.06793215, 62.87115487, -167.14056929]))
	from scipy.spatial.transform import Rotation
	R = Rotation.from_euler("zyz", reproject.rot2euler("cel,gal"))*Rotation.from_euler("zyz", e)
	assert np.allclose(R.as_matrix(), np.eye(3), atol=1e-12)
	assert np.allclose(reproject.rot2euler([0.1, 0.2, 0.3]), [0.1, 0.2, 0.3])
	with pytest.raises(ValueError): reproject.rot2euler("gal")

def test_uht_flat_host_side():
	from pixell_b200 import uharm, geometry, enmap
	ny, nx = 32, 48
	d = 2/60
	wcs = geometry.CarWCS(crval=[0, 0], cdelt=[-d, d], crpix=[nx/2+0.5, ny/2+0.5])
	uht = uharm.UHT((3, ny, nx), wcs)
	assert uht.mode == "flat" and uht.shape == (ny, nx) and uht.npix == ny*nx
	bl = np.linspace(1, 0, 30000)
	hp = uht.lprof2hprof(bl)
	assert np.allclose(np.asarray(hp), np.interp(np.asarray(enmap.modlmap((ny, nx), wcs)), np.arange(30000), bl))
	assert np.isclose(uht.sum_hprof(np.ones((ny, nx))), uht.ntot)
	shape, fw = geometry.fullsky_geometry(res=np.deg2rad(2.0))
	assert uharm.UHT(shape, fw).mode == "curved"
	assert uharm.res2lmax(np.deg2rad(1.0)) == 180

def test_wavelet_bases_and_scale_geometries():
	"""host side of pixell_b200.wavelets (reference wavelets.py:48-75, 131-161, 402-417, 472-495)"""
	from pixell_b200 import wavelets, geometry
	l = np.arange(400.0)
	cn = wavelets.CosineNeedlet([10, 30, 90, 270])
	assert cn.n == 4 and list(cn.lmaxs) == [30, 90, 270, 270] and list(cn.lmins) == [10, 10, 30, 90]
	tot = sum(cn(i, l)**2 for i in range(cn.n))
	assert np.allclose(tot[10:270], 1) and np.all(tot[:10] == 0) and np.all(tot[271:] == 0)
	assert cn(1, l)[30] == 1 and abs(cn(1, l)[10]) < 1e-15 and cn(1, l)[90] == 0
	bt = wavelets.ButterTrim(lmin=10, lmax=320)
	assert bt.n == 4 and bt.lmaxs[-1] == 320 and np.all(np.diff(bt.lmaxs) > 0)
	assert np.allclose(sum(bt(i, l)**2 for i in range(bt.n)), 1)
	for i in range(bt.n-1): assert np.all(bt(i, l)[l > bt.lmaxs[i]] == 0)        # harmonically compact
	assert np.all(wavelets.trim_kernel(np.array([0.0, 0.005, 0.5, 1.0]), 1e-2) == np.clip(np.array([0.0, 0.005, 0.5, 1.0])*1.02-0.01, 0, 1))
	# scale grids: full sky at the scale's resolution (never coarser than 2 degrees), cropped to the map's rows
	shape, wcs = geometry.fullsky_geometry(res=np.deg2rad(0.5))
	gs, gw = wavelets.make_wavelet_geometry_curved(shape, wcs, np.pi/100)
	assert gs == (100, 200)
	assert wavelets.make_wavelet_geometry_curved(shape, wcs, np.deg2rad(10))[0] == (90, 180)
	bs, bw = geometry.band_geometry(np.deg2rad([-30, 10]), res=np.deg2rad(0.5))
	cs_, cw = wavelets.make_wavelet_geometry_curved(bs, bw, np.pi/60)
	dec = np.rad2deg(geometry.dec_of(cw, np.array([-0.5, cs_[-2]-0.5])))
	assert cs_[-1] == 180 and dec.min() <= -30+1e-9 and dec.max() >= 10 and cs_[-2] <= 22      # 2 degree grid (minres), 21 rows
	cut, cutw = geometry.slice_geometry(shape, wcs, 0, shape[0], 0, shape[1]//2)
	with pytest.raises(NotImplementedError): wavelets.make_wavelet_geometry_curved(cut, cutw, np.pi/60)

def test_fft_host_helpers():
	"""pixell.fft's pure host helpers (reference fft.py:319-345, 389-433) as the reference behaves: values below were
	produced by the reference's functions (fft_len keeps the last product it meets, not the largest: fft_len(7) == 6)"""
	from pixell_b200 import fft as F
	want = {(7, "below"): 6, (7, "above"): 7, (100, "below"): 100, (100, "above"): 100, (1000, "below"): 1000, (1000, "above"): 1000,
		(4099, "below"): 4096, (4099, "above"): 4116, (10000.5, "below"): 10000, (10000.5, "above"): 10010}
	for (n, d), v in want.items(): assert F.fft_len(n, d) == v, (n, d)
	assert F.fft_len(1, "below") is None and F.fft_len(1, "above") == 2
	i = np.arange(10)
	assert np.allclose(F.ind2freq(10, i, 0.5), np.fft.fftfreq(10, 0.5)) and np.allclose(F.freq2ind(10, np.fft.fftfreq(10, 0.5), 0.5), i)
	assert np.allclose(F.int2rfreq(10, i[:6], 0.5), np.fft.rfftfreq(10, 0.5)) and np.allclose(F.rfreq2ind(10, np.fft.rfftfreq(10, 0.5), 0.5), i[:6])
	# resample_fft: low frequencies of both signs survive, the rest is zero padding / dropped
	fa = np.arange(8.0)+0j
	up = F.resample_fft(fa, 12)
	assert np.array_equal(up, np.array([0, 1, 2, 3, 0, 0, 0, 0, 4, 5, 6, 7])+0j)
	down = F.resample_fft(fa, 5, norm=2)
	assert np.array_equal(down, 2*np.array([0, 1, 5, 6, 7])+0j)
	acc = np.ones(12, complex)
	F.resample_fft(fa, 12, out=acc, op=lambda a, b: a+b)
	assert np.array_equal(acc, up+1)
	try:
		sys_path_added = False
		import sys
		if "/root/reference" not in sys.path: sys.path.insert(0, "/root/reference"); sys_path_added = True
		from pixell import fft as RF
	except Exception: RF = None
	if RF is not None:
		rng = np.random.default_rng(0); big = rng.standard_normal((3, 8, 9))+0j
		for n, axes in ((12, -1), ((5, 20), (-2, -1)), (4, 0)):
			assert np.array_equal(F.resample_fft(big, n, axes=axes, norm=0.5), RF.resample_fft(big, n, axes=axes, norm=0.5))
		for n in (2, 17.2, 360, 8191): assert F.fft_len(n) == RF.fft_len(n) and F.fft_len(n, "above", [3, 7]) == RF.fft_len(n, "above", [3, 7])
