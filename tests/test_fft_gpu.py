"""GPU parity tests of the FFT engine (K7; C ABI b2_fft_* through pixell_b200.fft / pixell_b200.enmap)
against numpy's pocketfft (the same algorithm family ducc0.fft ships; reference pixell/fft.py:33-60).
Tolerance: 1e-12 relative to the largest output element for float64, 2e-6 for float32."""
import numpy as np, pytest

pytestmark = pytest.mark.gpu

@pytest.fixture(scope="module")
def F():
	from pixell_b200 import fft
	return fft

def rel(a, b): return np.abs(a-b).max()/max(np.abs(b).max(), 1e-300)

def cdata(shape, seed=0, dtype=np.complex128):
	rng = np.random.default_rng(seed)
	return (rng.standard_normal(shape) + 1j*rng.standard_normal(shape)).astype(dtype)

# last-axis lengths: powers of two, smooth, primes (Bluestein), longer than one CTA's shared memory (split lines)
@pytest.mark.parametrize("shape", [(7, 64), (3, 1000), (5, 61), (2, 4099), (2, 2160), (2, 20000), (1, 32768), (1, 43200)])
def test_c2c_last_axis(F, shape):
	a = cdata(shape)
	got = F.fft(a, axes=[-1])
	assert rel(got, np.fft.fft(a, axis=-1)) < 1e-12
	back = F.ifft(got, axes=[-1], normalize=True)
	assert rel(back, a) < 1e-12

@pytest.mark.parametrize("shape", [(64, 33), (1000, 7), (61, 5), (20000, 3), (16384, 2)])
def test_c2c_strided_axis(F, shape):
	a = cdata(shape, 1)
	got = F.fft(a, axes=[0])
	assert rel(got, np.fft.fft(a, axis=0)) < 1e-12
	back = F.ifft(got, axes=[0])
	assert rel(back, a*shape[0]) < 1e-12

# (2, 2048, 24), (1, 4096, 6000): both axes run as thread-block clusters (lines split over distributed shared memory)
@pytest.mark.parametrize("shape", [(3, 48, 80), (2, 33, 45), (1, 128, 4099), (2, 4, 6, 10), (2, 2048, 24), (1, 4096, 6000), (1, 1280, 5000)])
def test_c2c_2d(F, shape):
	a = cdata(shape, 2)
	got = F.fft(a, axes=[-2, -1])
	assert rel(got, np.fft.fft2(a)) < 1e-12
	# in place
	b = a.copy()
	F.fft(b, b, axes=[-2, -1])
	assert rel(b, np.fft.fft2(a)) < 1e-12
	back = F.ifft(got, axes=[-2, -1], normalize=True)
	assert rel(back, a) < 1e-12

def test_strided_views(F):
	"""non-contiguous input and output views (reference tests/test_pixell.py:412-422)"""
	big = cdata((4, 40, 70), 3)
	a = big[1:3, 4:36, 3:67]
	obig = np.full((2, 50, 80), 7+7j)
	o = obig[:, 10:42, 8:72]
	F.fft(a, o, axes=[-2, -1])
	assert rel(o, np.fft.fft2(a)) < 1e-12
	assert np.all(obig[:, :10] == 7+7j) and np.all(obig[:, :, 72:] == 7+7j)      # untouched outside the view
	t = a.transpose(0, 2, 1)              # transform axes with swapped strides
	assert rel(F.fft(t, axes=[-2, -1]), np.fft.fft2(t)) < 1e-12

@pytest.mark.parametrize("shape", [(4, 100), (4, 101), (3, 61), (2, 20000), (1, 32768)])
def test_r2c_c2r_1d(F, shape):
	rng = np.random.default_rng(4)
	a = rng.standard_normal(shape)
	ft = F.rfft(a)
	assert ft.shape == shape[:-1]+(shape[-1]//2+1,)
	assert rel(ft, np.fft.rfft(a, axis=-1)) < 1e-12
	back = F.irfft(ft, n=shape[-1], normalize=True)
	assert rel(back, a) < 1e-12

@pytest.mark.parametrize("shape", [(3, 50, 64), (2, 33, 45), (1, 64, 4100), (2, 1000, 18), (1, 2048, 10000), (2, 1536, 16384), (1, 6, 65536)])
def test_r2c_c2r_2d(F, shape):
	rng = np.random.default_rng(5)
	a = rng.standard_normal(shape)
	ft = F.rfft(a, axes=[-2, -1])
	assert rel(ft, np.fft.rfft2(a)) < 1e-12
	back = F.irfft(ft, n=shape[-1], axes=[-2, -1])
	assert rel(back, a*shape[-1]*shape[-2]) < 1e-12

def test_float32(F):
	a = cdata((3, 32, 96), 6, np.complex64)
	got = F.fft(a, axes=[-2, -1])
	assert got.dtype == np.complex64
	assert rel(got, np.fft.fft2(a.astype(np.complex128))) < 2e-6
	r = np.random.default_rng(7).standard_normal((2, 40, 50)).astype(np.float32)
	ft = F.rfft(r, axes=[-2, -1])
	assert ft.dtype == np.complex64 and rel(ft, np.fft.rfft2(r.astype(np.float64))) < 2e-6
	assert rel(F.irfft(ft, n=50, axes=[-2, -1], normalize=True), r) < 2e-6

def test_float32_long_lines(F):
	"""float32 data on lines long enough for the split / cluster passes"""
	rng = np.random.default_rng(16)
	a = rng.standard_normal((2, 2048, 10000)).astype(np.float32)
	ft = F.rfft(a, axes=[-2, -1])
	assert ft.dtype == np.complex64
	assert rel(ft, np.fft.rfft2(a.astype(np.float64))) < 5e-6
	back = F.irfft(ft, n=a.shape[-1], axes=[-2, -1], normalize=True)
	assert rel(back, a) < 5e-6
	c = cdata((1, 2048, 6000), 17, np.complex64)
	assert rel(F.fft(c, axes=[-2, -1]), np.fft.fft2(c.astype(np.complex128))) < 5e-6

def test_cluster_split(F, monkeypatch):
	"""B2_FFT_CLUSTER=1: long lines split over thread-block clusters (distributed shared memory) give the same results"""
	monkeypatch.setenv("B2_FFT_CLUSTER", "1")
	F.clear_plans()
	try:
		a = cdata((1, 4096, 6000), 20)
		want = np.fft.fft2(a)
		assert rel(F.fft(a, axes=[-2, -1]), want) < 1e-12
		b = a.copy(); F.fft(b, b, axes=[-2, -1])
		assert rel(b, want) < 1e-12
		assert rel(F.ifft(want, axes=[-2, -1], normalize=True), a) < 1e-12
		rng = np.random.default_rng(21)
		for shape, dt, tol in [((2, 1536, 16384), np.float64, 1e-12), ((1, 6, 65536), np.float64, 1e-12), ((1, 2048, 10000), np.float32, 5e-6)]:
			m = rng.standard_normal(shape).astype(dt)
			ft = F.rfft(m, axes=[-2, -1])
			assert rel(ft, np.fft.rfft2(m.astype(np.float64))) < tol
			assert rel(F.irfft(ft, n=shape[-1], axes=[-2, -1], normalize=True), m) < tol
		c = cdata((20000, 3), 22)
		assert rel(F.fft(c, axes=[0]), np.fft.fft(c, axis=0)) < 1e-12
	finally:
		monkeypatch.delenv("B2_FFT_CLUSTER")
		F.clear_plans()

def test_engine_object(F):
	"""the plug-in shape pixell.fft.engines expects (reference pixell/fft.py:8-60, 126-131)"""
	a = cdata((2, 30, 24), 8); b = np.empty_like(a)
	plan = F.engine.FFTW(a, b, axes=(-2, -1), direction="FFTW_FORWARD", threads=4, flags=["FFTW_ESTIMATE"])
	plan()
	assert rel(b, np.fft.fft2(a)) < 1e-12
	c = np.empty_like(a)
	F.engine.FFTW(b, c, axes=(-2, -1), direction="FFTW_BACKWARD")(normalise_idft=True)
	assert rel(c, a) < 1e-12
	r = np.random.default_rng(9).standard_normal((3, 90)); fr = np.empty((3, 46), complex)
	F.engine.FFTW(r, fr, axes=(-1,), direction="FFTW_FORWARD")()
	assert rel(fr, np.fft.rfft(r)) < 1e-12
	rr = F.engine.empty_aligned((3, 90), np.float64)
	F.engine.FFTW(fr, rr, axes=(-1,), direction="FFTW_BACKWARD")()
	assert rel(rr, r*90) < 1e-12
	with pytest.raises(ValueError): F.engine.FFTW(r, rr, direction=["FFTW_NOSUCHKIND"])      # r2r kinds: see test_dct_dst_all_kinds

def test_errors(F):
	a = cdata((4, 8));
	with pytest.raises(ValueError): F.transform(a, np.empty((4, 7), complex), [-1], True)
	with pytest.raises(ValueError): F.transform(a, np.empty((4, 8), np.complex64), [-1], True)
	with pytest.raises(ValueError): F.transform(a[:, ::-1], np.empty((4, 8), complex), [-1], True)

def test_torch_device(F):
	import torch
	a = cdata((3, 64, 80), 10)
	ta = torch.from_numpy(a).cuda()
	got = F.fft(ta, axes=[-2, -1])
	assert got.is_cuda and rel(got.cpu().numpy(), np.fft.fft2(a)) < 1e-12
	r = torch.randn((2, 100, 128), dtype=torch.float64, device="cuda")
	ft = F.rfft(r, axes=[-2, -1])
	assert rel(ft.cpu().numpy(), np.fft.rfft2(r.cpu().numpy())) < 1e-12
	back = F.irfft(ft, n=128, axes=[-2, -1], normalize=True)
	assert rel(back.cpu().numpy(), r.cpu().numpy()) < 1e-12

def test_enmap_fft_roundtrip():
	"""enmap.fft / ifft normalisation conventions (reference pixell/enmap.py:1307-1337; tests/test_pixell.py test_fft*)"""
	from pixell_b200 import enmap, geometry
	shape, wcs = geometry.slice_geometry(*geometry.fullsky_geometry(res=np.deg2rad(0.5)), 100, 164, 200, 296)
	rng = np.random.default_rng(11)
	m = geometry.ndmap(rng.standard_normal((3,)+shape), wcs)
	f = enmap.fft(m)
	assert rel(f, np.fft.fft2(m)/np.sqrt(m.shape[-1]*m.shape[-2])) < 1e-12
	assert rel(enmap.ifft(f).real, m) < 1e-12
	fp = enmap.fft(m, normalize="phys")
	assert rel(fp, np.asarray(f)*enmap.pixsize(shape, wcs)**0.5) < 1e-12
	assert rel(enmap.ifft(fp, normalize="phys").real, m) < 1e-12
	# Parseval with the symmetric normalisation
	assert abs(np.sum(np.abs(f)**2)/np.sum(np.asarray(m)**2) - 1) < 1e-12
	# harmonic Gaussian smoothing against its numpy restatement
	sigma = np.deg2rad(1.0)
	ly, lx = enmap.laxes(shape, wcs)
	want = np.fft.ifft2(np.fft.fft2(m)*np.exp(-0.5*sigma**2*(ly[:, None]**2+lx[None, :]**2))).real
	assert rel(enmap.smooth_gauss(m, sigma), want) < 1e-12

def test_fft_input_shape(F):
	"""reference tests/test_pixell.py:380-432 (test_fft_input_shape), same five cases"""
	signal = np.ones((1, 2, 5)); signal[0, 1, :] = 10.
	out_exp = np.zeros((1, 2, 5), dtype=np.complex128); out_exp[0, 0, 0] = 5; out_exp[0, 1, 0] = 50
	out = F.fft(signal)
	np.testing.assert_allclose(out, out_exp, atol=1e-12); assert out.flags["C_CONTIGUOUS"]
	signal = np.ones((1, 5, 2)); signal[0, :, 1] = 10.
	out_exp = np.zeros((1, 5, 2), dtype=np.complex128); out_exp[0, 0, 0] = 5; out_exp[0, 0, 1] = 50
	out = F.fft(signal, axes=[-2])
	np.testing.assert_allclose(out, out_exp, atol=1e-12); assert out.flags["C_CONTIGUOUS"]
	signal = np.ones((1, 2, 5, 10)); signal[0, 1, :] = 10.
	out_exp = np.zeros((1, 2, 5, 10), dtype=np.complex128); out_exp[0, 0, 0, 0] = 50; out_exp[0, 1, 0, 0] = 500
	out = F.fft(signal, axes=[-2, -1])
	np.testing.assert_allclose(out, out_exp, atol=1e-12); assert out.flags["C_CONTIGUOUS"]
	signal = np.ones((1, 2, 5, 10), dtype=np.complex128); signal[0, 1, :] = 10
	ft = np.zeros((5, 10, 1, 2), dtype=np.complex128).transpose(2, 3, 0, 1)          # non-contiguous output
	out = F.fft(signal, ft=ft, axes=[-2, -1])
	np.testing.assert_allclose(out, out_exp, atol=1e-12)
	assert np.shares_memory(ft, out) and not out.flags["C_CONTIGUOUS"]
	signal = np.ones((1, 5, 10, 2)); signal[0, :, :, 1] = 10.
	out_exp = np.zeros((1, 5, 10, 2), dtype=np.complex128); out_exp[0, 0, 0, 0] = 50; out_exp[0, 0, 0, 1] = 500
	out = F.fft(signal, axes=[-3, -2])
	np.testing.assert_allclose(out, out_exp, atol=1e-12); assert out.flags["C_CONTIGUOUS"]

def test_ifft_input_shape(F):
	"""reference tests/test_pixell.py:434-486 (test_ifft_input_shape)"""
	fsignal = np.ones((1, 2, 5), dtype=np.complex128); fsignal[0, 1, :] = 10.
	out_exp = np.zeros((1, 2, 5)); out_exp[0, 0, 0] = 5; out_exp[0, 1, 0] = 50
	out = F.ifft(fsignal)
	np.testing.assert_allclose(out, out_exp, atol=1e-12); assert out.flags["C_CONTIGUOUS"]
	fsignal = np.ones((1, 5, 2), dtype=np.complex128); fsignal[0, :, 1] = 10.
	out_exp = np.zeros((1, 5, 2)); out_exp[0, 0, 0] = 5; out_exp[0, 0, 1] = 50
	out = F.ifft(fsignal, axes=[-2])
	np.testing.assert_allclose(out, out_exp, atol=1e-12)
	fsignal = np.ones((1, 2, 5, 10), dtype=np.complex128); fsignal[0, 1, :] = 10.
	out_exp = np.zeros((1, 2, 5, 10)); out_exp[0, 0, 0, 0] = 50; out_exp[0, 1, 0, 0] = 500
	out = F.ifft(fsignal, axes=[-2, -1])
	np.testing.assert_allclose(out, out_exp, atol=1e-12)
	tod = np.zeros((5, 10, 1, 2), dtype=np.complex128).transpose(2, 3, 0, 1)
	out = F.ifft(fsignal, tod=tod, axes=[-2, -1])
	assert np.shares_memory(tod, out) and not out.flags["C_CONTIGUOUS"]
	np.testing.assert_allclose(out, out_exp, atol=1e-12)
	fsignal = np.ones((1, 5, 10, 2), dtype=np.complex128); fsignal[0, :, :, 1] = 10.
	out_exp = np.zeros((1, 5, 10, 2)); out_exp[0, 0, 0, 0] = 50; out_exp[0, 0, 0, 1] = 500
	out = F.ifft(fsignal, axes=[-3, -2])
	np.testing.assert_allclose(out, out_exp, atol=1e-12)

def test_engine_behind_reference_fft_functions(F):
	"""what pixell/fft.py:133-187 does with a registered engine (engines[name].FFTW(tod, ft, flags=, threads=, axes=,
	direction=) then plan()): restated here because /root/reference does not travel to the GPU box"""
	def ref_fft(tod, ft=None, axes=[-1]):
		tod = np.asarray(tod, np.result_type(tod, 0.0))
		if ft is None:
			otype = np.result_type(tod.dtype, 0j); ft = F.engine.empty_aligned(tod.shape, otype, n=32); tod = tod.astype(otype, copy=False)
		F.engine.FFTW(tod, ft, flags=["FFTW_ESTIMATE"], threads=8, axes=tuple(axes), direction="FFTW_FORWARD")()
		return ft
	def ref_ifft(ft, tod=None, axes=[-1], normalize=False):
		if tod is None: tod = F.engine.empty_aligned(ft.shape, ft.dtype, n=32)
		F.engine.FFTW(ft, tod, flags=["FFTW_ESTIMATE"], direction="FFTW_BACKWARD", threads=8, axes=tuple(axes))(normalise_idft=False)
		if normalize: tod /= np.prod([tod.shape[i] for i in axes])
		return tod
	rng = np.random.default_rng(12)
	a = rng.standard_normal((3, 40, 56))
	f = ref_fft(a, axes=[-2, -1])
	assert rel(f, np.fft.fft2(a)) < 1e-12
	assert rel(ref_ifft(f, axes=[-2, -1], normalize=True).real, a) < 1e-12
	hf = np.empty((3, 40, 29), complex)
	ref_fft(a, hf, axes=[-2, -1])
	assert rel(hf, np.fft.rfft2(a)) < 1e-12
	back = np.empty_like(a)
	ref_ifft(hf, back, axes=[-2, -1], normalize=True)
	assert rel(back, a) < 1e-12

def _patch(shape, res=0.01):
	"""small CAR patch centred on (0, 0) with `res` radians per pixel (enmap.geometry(pos=(0,0), shape=, res=))"""
	from pixell_b200 import geometry
	ny, nx = shape
	d = np.rad2deg(res)
	return geometry.CarWCS(crval=[0, 0], cdelt=[-d, d], crpix=[nx/2+0.5, ny/2+0.5])

@pytest.mark.parametrize("ishape", [(10, 10), (10, 11), (11, 10), (11, 11)])
@pytest.mark.parametrize("dtype", [np.complex64, np.complex128])
def test_queb_rotmat_complex(ishape, dtype):
	"""mirror of the reference tests/test_pixell.py:488-518: harm2map -> map2harm round trip of a map that is complex in real space"""
	from pixell_b200 import enmap, geometry
	wcs = _patch(ishape)
	atol = 1e-10 if dtype == np.complex128 else 1e-5
	for comp in (1, 2):
		inp = geometry.ndmap(np.zeros((3,)+ishape, dtype), wcs)
		inp[comp] += 1. + 1.j
		out = enmap.map2harm(enmap.harm2map(inp, keep_imag=True))
		assert out.dtype == dtype
		assert np.allclose(out, inp, rtol=0, atol=atol), (ishape, dtype, comp)

@pytest.mark.parametrize("ishape", [(10, 10), (10, 11), (11, 10), (11, 11)])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_queb_rotmat_real(ishape, dtype):
	"""mirror of the reference tests/test_pixell.py:520-541: map2harm(harm2map(h)) = h for the Fourier map of a real map"""
	from pixell_b200 import enmap, geometry
	wcs = _patch(ishape)
	rng = np.random.default_rng(0)
	m = geometry.ndmap(rng.standard_normal((3,)+ishape).astype(dtype), wcs)
	h = enmap.map2harm(m)
	atol = 1e-10 if dtype == np.float64 else 1e-5
	assert np.allclose(enmap.map2harm(enmap.harm2map(h)), h, rtol=0, atol=atol)

@pytest.mark.parametrize("iau", [False, True])
@pytest.mark.parametrize("spin", [[0, 2], [0, 1], [2, 0]])
def test_map2harm_against_numpy(iau, spin):
	"""map2harm / harm2map against the reference formulas evaluated with numpy: fft2/sqrt(npix), then queb_rotmat and
	map_mul per spin pair (pixell/enmap.py:1358-1400); torch CUDA input gives the same numbers as numpy input"""
	import torch
	from pixell_b200 import enmap, geometry
	shape = (3, 36, 50)
	wcs = _patch(shape[-2:], res=np.deg2rad(0.5))
	rng = np.random.default_rng(7)
	m = rng.standard_normal(shape)
	want = np.fft.fft2(m)/np.sqrt(36*50)
	lm = enmap.lmap(shape, wcs)
	for s, i1, i2 in enmap.spin_helper(spin, 3):
		if s == 0: continue
		want[i1:i2] = enmap.map_mul(enmap.queb_rotmat(lm, iau=iau, spin=s), want[i1:i2])
	got = enmap.map2harm(geometry.ndmap(m, wcs), iau=iau, spin=spin)
	assert rel(np.asarray(got), want) < 1e-12
	tgot = enmap.map2harm(torch.from_numpy(m).cuda(), iau=iau, spin=spin, wcs=wcs)
	assert rel(tgot.cpu().numpy(), want) < 1e-12
	back = enmap.harm2map(got, iau=iau, spin=spin)
	assert back.dtype == np.float64 and rel(np.asarray(back), m) < 1e-12
	# "phys" normalisation and the adjoint pair: <harm2map_adjoint(x), y> = <x, harm2map(y)>
	h = enmap.map2harm(geometry.ndmap(m, wcs), normalize="phys", spin=spin)
	assert rel(np.asarray(h), np.asarray(got if not iau else enmap.map2harm(geometry.ndmap(m, wcs), spin=spin))*enmap.pixsize(shape, wcs)**0.5) < 1e-12
	y = geometry.ndmap(rng.standard_normal(shape) + 1j*rng.standard_normal(shape), wcs)
	lhs = np.vdot(np.asarray(enmap.harm2map_adjoint(geometry.ndmap(m, wcs), spin=spin)), np.asarray(y))
	rhs = np.vdot(m, np.asarray(enmap.harm2map(y, spin=spin, keep_imag=True)))
	assert abs(lhs-rhs) < 1e-10*abs(lhs)

def test_rotate_pol():
	from pixell_b200 import enmap
	rng = np.random.default_rng(8)
	m = rng.standard_normal((3, 5, 6))
	r = enmap.rotate_pol(m, 0.3)
	c, s = np.cos(0.6), np.sin(0.6)
	assert np.allclose(r[0], m[0]) and np.allclose(r[1], c*m[1]-s*m[2]) and np.allclose(r[2], s*m[1]+c*m[2])

@pytest.mark.parametrize("order", [0, 1])
def test_apply_window(order):
	"""pixel window in Fourier space (reference enmap.py:1470-1500) against numpy; unapply undoes it"""
	from pixell_b200 import enmap, geometry
	wcs = _patch((40, 54))
	rng = np.random.default_rng(11)
	m = geometry.ndmap(rng.standard_normal((2, 40, 54)), wcs)
	wy, wx = enmap.calc_window(m.shape, order=order)
	want = np.fft.ifft2(np.fft.fft2(m)*wy[:, None]*wx[None, :]).real
	got = enmap.apply_window(m, order=order)
	assert rel(np.asarray(got), want) < 1e-12
	assert rel(np.asarray(enmap.unapply_window(got, order=order)), np.asarray(m)) < 1e-10
	f = enmap.fft(m)
	assert rel(np.asarray(enmap.apply_window(f, order=order, nofft=True)), np.asarray(f)*wy[:, None]*wx[None, :]) < 1e-13
	# device tensors take the one-pass filter kernel
	import torch
	tm = torch.from_numpy(np.asarray(m)).cuda()
	assert rel(enmap.apply_window(tm, order=order, wcs=wcs).cpu().numpy(), want) < 1e-12
	assert rel(enmap.smooth_gauss(tm, 0.01, wcs=wcs).cpu().numpy(), np.asarray(enmap.smooth_gauss(m, 0.01))) < 1e-12

# ------------------------------------------------------------------ r2r (DCT / DST), reference pixell/fft.py:211-317

_SCIPY = {"DCT-I": ("dct", 1), "DCT-II": ("dct", 2), "DCT-III": ("dct", 3), "DCT-IV": ("dct", 4),
	"DST-I": ("dst", 1), "DST-II": ("dst", 2), "DST-III": ("dst", 3), "DST-IV": ("dst", 4)}

@pytest.mark.parametrize("type", sorted(_SCIPY))
@pytest.mark.parametrize("shape,axes", [((3, 17), [-1]), ((2, 64), [-1]), ((6, 9, 20), [-2, -1]), ((33, 4), [0])])
def test_dct_dst_all_kinds(F, type, shape, axes):
	"""every FFTW r2r kind against scipy.fft (pocketfft), unnormalised as FFTW / the reference define them, and the
	matching inverse with the reference's normalisation 2 (N + d) per axis"""
	import scipy.fft as sf
	a = np.random.default_rng(7).standard_normal(shape)
	fn, t = _SCIPY[type]
	want = a
	for ax in axes: want = getattr(sf, fn)(want, type=t, axis=ax)
	got = F.dct(a, axes=axes, type=type)
	assert rel(got, want) < 1e-12
	back = F.idct(got, axes=axes, type=type, normalize=True)
	assert rel(back, a) < 1e-12

def test_dct_torch_and_chebyshev(F):
	import torch, scipy.fft as sf
	a = np.random.default_rng(8).standard_normal((4, 33))
	got = F.dct(torch.from_numpy(a).cuda(), type="DCT-II")
	assert got.is_cuda and rel(got.cpu().numpy(), sf.dct(a, type=2, axis=-1)) < 1e-12
	assert rel(F.redft00(a), sf.dct(a, type=1, axis=-1)) < 1e-12
	# chebt / ichebt are inverses of each other (one-dimensional input, reference fft.py:307-317)
	c = np.random.default_rng(9).standard_normal(40)
	assert rel(F.ichebt(F.chebt(c)), c) < 1e-12
	# the engine plug-in takes the r2r kinds as a list of directions, one per axis (reference fft.py:66-71)
	b = np.empty_like(a)
	F.engine.FFTW(a, b, axes=(-1,), direction=["FFTW_REDFT10"])()
	assert rel(b, sf.dct(a, type=2, axis=-1)) < 1e-12

def test_enmap_dct_roundtrip():
	"""enmap.fft(dct=True) / ifft(dct=True) (reference enmap.py:1314-1333): DCT-I over both axes with the reference's
	prod(2 n - 1)^(1/2) normalisation on each side"""
	import scipy.fft as sf
	from pixell_b200 import enmap
	m = np.random.default_rng(10).standard_normal((2, 24, 36))
	got = enmap.dct(m)
	want = sf.dct(sf.dct(m, type=1, axis=-1), type=1, axis=-2)/np.sqrt((2*24-1)*(2*36-1))
	assert rel(np.asarray(got), want) < 1e-12
	back = enmap.idct(got)
	assert rel(np.asarray(back), m*(4*23*35)/((2*24-1)*(2*36-1))) < 1e-12

def test_shift_and_resample(F):
	"""fft.shift / fft.resample (reference fft.py:350-387) on the engine: integer shifts are rolls, a band-limited signal
	resampled to a finer grid is the signal evaluated there, and back"""
	x = np.random.default_rng(11).standard_normal((3, 40))
	assert rel(F.shift(x, 3), np.roll(x, 3, -1)) < 1e-12
	assert rel(F.shift(x+0j, [1, -2], axes=(0, 1)), np.roll(np.roll(x, 1, 0), -2, 1)+0j) < 1e-12
	t = np.arange(40)/40.0
	sig = np.cos(2*np.pi*3*t) + 0.5*np.sin(2*np.pi*7*t)
	dsig = -2*np.pi*3/40*np.sin(2*np.pi*3*t) + 0.5*2*np.pi*7/40*np.cos(2*np.pi*7*t)
	assert rel(F.shift(sig, 0.0, deriv=0), -dsig) < 1e-11      # the derivative with respect to the SHIFT, d/ds a(x - s) = -a'(x), as the reference defines it
	t2 = np.arange(100)/100.0
	fine = F.resample(sig, 100)
	assert rel(fine, np.cos(2*np.pi*3*t2) + 0.5*np.sin(2*np.pi*7*t2)) < 1e-12
	assert rel(F.resample(fine, 40), sig) < 1e-12
	# two axes at once, band-limited input (the reference keeps the first c//2 and the last c - c//2 bins of each axis, so
	# only content below c//2 survives an up-and-down trip unchanged; mirrored as it behaves)
	y, x = np.meshgrid(np.arange(13)/13.0, np.arange(19)/19.0, indexing="ij")
	img = np.stack([np.cos(2*np.pi*(2*y+3*x)) + 0.3*np.sin(2*np.pi*(4*y-5*x)), np.cos(2*np.pi*x) + 1.0])
	y2, x2 = np.meshgrid(np.arange(26)/26.0, np.arange(38)/38.0, indexing="ij")
	want = np.stack([np.cos(2*np.pi*(2*y2+3*x2)) + 0.3*np.sin(2*np.pi*(4*y2-5*x2)), np.cos(2*np.pi*x2) + 1.0])
	up = F.resample(img, (26, 38))
	assert up.shape == (2, 26, 38) and rel(up, want) < 1e-12
	assert rel(F.resample(up, (13, 19)), img) < 1e-12

def test_fourier_filter(F):
	"""b2_fourier_filter: the separable and the full 2-D filter against numpy broadcasting, float64 and float32, batches"""
	import torch
	rng = np.random.default_rng(13)
	for dt, tol in ((np.complex128, 1e-15), (np.complex64, 1e-6)):
		a = (rng.standard_normal((3, 17, 40)) + 1j*rng.standard_normal((3, 17, 40))).astype(dt)
		fy, fx, f2 = rng.standard_normal(17), rng.standard_normal(40), rng.standard_normal((17, 40))
		t = torch.from_numpy(a.copy()).cuda()
		assert rel(F.fourier_filter(t, fy=fy, fx=fx).cpu().numpy(), a*fy[:, None]*fx[None, :]) < tol
		t = torch.from_numpy(a.copy()).cuda()
		assert rel(F.fourier_filter(t, f2=f2).cpu().numpy(), a*f2) < tol
		t = torch.from_numpy(a[0].copy()).cuda()
		assert rel(F.fourier_filter(t, fy=fy, fx=fx).cpu().numpy(), a[0]*fy[:, None]*fx[None, :]) < tol
	with pytest.raises(ValueError): F.fourier_filter(torch.zeros((4, 5), dtype=torch.complex128, device="cuda"), fy=np.ones(3), fx=np.ones(5))
