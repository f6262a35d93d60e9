"""pixell_b200.sht -- drop-in for the ducc0.sht.experimental functions pixell's curvedsky calls
(reference pixell/curvedsky.py:328-329, 398-399, 501, 531-555, 855, 907-924, 936-960, 1032-1046,
1068-1084), backed by the CUDA engine in libb200sht.so.  Same keyword-only signatures, same
in-place-and-return behaviour.  Inputs may be numpy arrays (host memory: staged through the GPU)
or torch CUDA tensors (zero-copy).  `nthreads` is accepted and ignored.

Rings may share nphi and phi0 (CAR maps: the fast path) or carry their own (HEALPix, single-pixel rings).
Limits: lstride/pixstride other than what pixell uses are refused.
synthesis_general / adjoint_synthesis_general (arbitrary positions, call site curvedsky.py:993-1016) are provided.
"""
import collections, ctypes
import numpy as np
from . import _lib as L

_MODES = {"STANDARD": L.MODE_STANDARD, "DERIV1": L.MODE_DERIV1}

# ------------------------------------------------------------------ plans (cached)

class Plan:
	def __init__(self, handle): self.handle = handle
	def __del__(self):
		try:
			if self.handle: L.lib().b2_sht_plan_destroy(self.handle); self.handle = None
		except Exception: pass
	@property
	def nbytes(self): return L.lib().b2_sht_plan_bytes(self.handle)
	def last_timing(self):
		"""device milliseconds of the last transform: dict(legendre, ringfft, resample, copies)"""
		out = (ctypes.c_double*4)()
		L.check(L.lib().b2_sht_last_timing(self.handle, out))
		return dict(legendre=out[0], ringfft=out[1], resample=out[2], copies=out[3])

_plans = collections.OrderedDict()
PLAN_CACHE_SIZE = 4

def _cached(key, make):
	dev = L.init()
	key = (dev,)+key
	p = _plans.get(key)
	if p is None:
		while len(_plans) >= PLAN_CACHE_SIZE: _plans.popitem(last=False)
		p = make(); _plans[key] = p
	else: _plans.move_to_end(key)
	return p

def clear_plans(): _plans.clear()

def default_mstart(lmax, mmax):
	m = np.arange(mmax+1, dtype=np.int64)
	return m*(2*lmax+1-m)//2

def _layout(lmax, mmax, mstart, lstride):
	if mmax is None: mmax = lmax
	mstart = default_mstart(lmax, mmax) if mstart is None else L.as_i64(mstart)[:mmax+1]
	if len(mstart) != mmax+1: raise ValueError("mstart must have mmax+1 entries")
	return int(lmax), int(mmax), mstart, int(lstride)

def plan_rings(theta, nphi, phi0, ringstart, lmax, mmax=None, mstart=None, lstride=1, weight=None, xdir=1, npix=None):
	theta = np.ascontiguousarray(theta, dtype=np.float64)
	nphi_a = np.atleast_1d(np.asarray(nphi)).astype(np.int64); phi0_a = np.atleast_1d(np.asarray(phi0, dtype=np.float64))
	if np.any(nphi_a != nphi_a[0]) or np.any(phi0_a != phi0_a[0]):
		# HEALPix-like ring sets: per-ring nphi / phi0 (one FFT group per distinct nphi inside the engine)
		if xdir != 1 or npix is not None: raise NotImplementedError("pixell_b200: rings with individual nphi/phi0 cannot be flipped or cut")
		nphi_a = np.ascontiguousarray(np.broadcast_to(nphi_a, theta.shape)); phi0_a = np.ascontiguousarray(np.broadcast_to(phi0_a, theta.shape))
		ringstart = L.as_i64(ringstart)
		lmax, mmax, mstart, lstride = _layout(lmax, mmax, mstart, lstride)
		w = None if weight is None else np.ascontiguousarray(weight, dtype=np.float64)
		key = ("general", theta.tobytes(), nphi_a.tobytes(), phi0_a.tobytes(), ringstart.tobytes(),
			None if w is None else w.tobytes(), lmax, mmax, mstart.tobytes(), lstride)
		def make():
			h = ctypes.c_void_p()
			L.check(L.lib().b2_sht_plan_rings_general(ctypes.byref(h), len(theta), L.p_dbl(theta), L.p_i64(nphi_a), L.p_dbl(phi0_a),
				L.p_i64(ringstart), None if w is None else L.p_dbl(w), lmax, mmax, L.p_i64(mstart), lstride), ValueError)
			return Plan(h)
		return _cached(key, make)
	nphi0, phi00 = int(nphi_a[0]), float(phi0_a[0])
	ringstart = L.as_i64(ringstart)
	lmax, mmax, mstart, lstride = _layout(lmax, mmax, mstart, lstride)
	npix = nphi0 if npix is None else int(npix)
	w = None if weight is None else np.ascontiguousarray(weight, dtype=np.float64)
	key = ("rings", theta.tobytes(), nphi0, phi00, int(xdir), npix, ringstart.tobytes(),
		None if w is None else w.tobytes(), lmax, mmax, mstart.tobytes(), lstride)
	def make():
		h = ctypes.c_void_p()
		L.check(L.lib().b2_sht_plan_rings(ctypes.byref(h), len(theta), L.p_dbl(theta), nphi0, phi00, int(xdir), npix,
			L.p_i64(ringstart), None if w is None else L.p_dbl(w), lmax, mmax, L.p_i64(mstart), lstride), ValueError)
		return Plan(h)
	return _cached(key, make)

def plan_2d(geometry, ntheta, nphi, phi0, lmax, mmax=None, mstart=None, lstride=1, flip_y=False, flip_x=False):
	lmax, mmax, mstart, lstride = _layout(lmax, mmax, mstart, lstride)
	key = ("2d", geometry, int(ntheta), int(nphi), float(phi0), bool(flip_y), bool(flip_x), lmax, mmax, mstart.tobytes(), lstride)
	def make():
		h = ctypes.c_void_p()
		L.check(L.lib().b2_sht_plan_2d(ctypes.byref(h), geometry.encode(), int(ntheta), int(nphi), float(phi0),
			int(bool(flip_y)), int(bool(flip_x)), lmax, mmax, L.p_i64(mstart), lstride), ValueError)
		return Plan(h)
	return _cached(key, make)

# ------------------------------------------------------------------ execution helpers

def _empty_like(ref, shape, dtype):
	if L.is_torch(ref):
		import torch
		td = {np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32,
			np.dtype(np.complex128): torch.complex128, np.dtype(np.complex64): torch.complex64}[np.dtype(dtype)]
		return torch.zeros(shape, dtype=td, device=ref.device)
	return np.zeros(shape, dtype)

def _check_last_contig(a, name, nlast=1):
	st = L.strides_elems(a)
	want = 1
	for i in range(1, nlast+1):
		if a.shape[-i] != 1 and st[-i] != want: raise ValueError("%s must be C-contiguous in its last %d axes" % (name, nlast))
		want *= a.shape[-i]

def _run(fn, plan, spin, mode, alm, map, nmapdim, has_mode=True):
	"""alm [nca, nalm], map [ncm, npix] or [ncm, ny, nx]"""
	pa, mema, dta = L.buffer_info(alm); pm, memm, dtm = L.buffer_info(map)
	if mema != memm: raise ValueError("alm and map must both be host arrays or both be CUDA tensors")
	if dta == np.complex128 and dtm == np.float64: dtype = L.F64
	elif dta == np.complex64 and dtm == np.float32: dtype = L.F32
	else: raise ValueError("alm/map dtypes must be (complex128, float64) or (complex64, float32), got (%s, %s)" % (dta, dtm))
	ncm = 1 if spin == 0 else 2
	nca = 1 if (spin == 0 or mode == L.MODE_DERIV1) else 2
	if alm.ndim != 2 or alm.shape[0] != nca: raise ValueError("alm must have shape [%d, nalm] for spin %d" % (nca, spin))
	if map.ndim != 1+nmapdim or map.shape[0] != ncm: raise ValueError("map must have %d components for spin %d" % (ncm, spin))
	_check_last_contig(alm, "alm", 1); _check_last_contig(map, "map", nmapdim)
	acs = L.strides_elems(alm)[0] if nca > 1 else 0
	mcs = L.strides_elems(map)[0] if ncm > 1 else 0
	stream = L.current_stream(map)
	args = [plan.handle, int(spin)] + ([mode] if has_mode else []) + [dtype, 1, pa, acs, 0, pm, mcs, 0, mema, stream]
	L.check(fn(*args))

OPS = {"synthesis": 0, "adjoint_synthesis": 1, "analysis_2d": 2, "adjoint_analysis_2d": 3}

def run_groups(plan, op, groups, alm, map, mode=L.MODE_STANDARD):
	"""Several spin groups of one component-stacked alm [ncomp, nalm] / map [ncomp, ...] pair in one engine call
	(b2_sht_execute_groups): groups = [(spin, j1, j2), ...] as enmap.spin_helper yields them.  Host arrays are
	pipelined (copies of one group overlap the kernels of the next).  The caller guarantees contiguous trailing axes."""
	pa, mema, dta = L.buffer_info(alm); pm, memm, dtm = L.buffer_info(map)
	if mema != memm: raise ValueError("alm and map must both be host arrays or both be CUDA tensors")
	if dta == np.complex128 and dtm == np.float64: dtype = L.F64
	elif dta == np.complex64 and dtm == np.float32: dtype = L.F32
	else: raise ValueError("alm/map dtypes must be (complex128, float64) or (complex64, float32), got (%s, %s)" % (dta, dtm))
	acs, mcs = L.strides_elems(alm)[0], L.strides_elems(map)[0]
	n = len(groups)
	spins = (ctypes.c_int*n)(*[int(g[0]) for g in groups])
	alms = (ctypes.c_void_p*n)(*[int(pa) + int(g[1])*int(acs)*int(dta.itemsize) for g in groups])
	maps = (ctypes.c_void_p*n)(*[int(pm) + int(g[1])*int(mcs)*int(dtm.itemsize) for g in groups])
	a_cs = (ctypes.c_int64*n)(*[int(acs)]*n); m_cs = (ctypes.c_int64*n)(*[int(mcs)]*n)
	L.check(L.lib().b2_sht_execute_groups(plan.handle, OPS[op], n, spins, mode, dtype, alms, a_cs, maps, m_cs, mema, L.current_stream(map)))

def run_batch(plan, spin, alm, map):
	"""Synthesis of a batch of alm sets of one spin in one engine call (b2_synthesis with nbatch > 1): alm [nb, nca, nalm],
	map [nb, ncm, ny, nx] (or [nb, ncm, npix]), device tensors or host arrays, trailing axes contiguous, uniform batch and
	component strides.  Device-resident float64 batches go through the batched Legendre kernels (the members share the
	recurrence); the results are bit-identical to member-by-member calls."""
	pa, mema, dta = L.buffer_info(alm); pm, memm, dtm = L.buffer_info(map)
	if mema != memm: raise ValueError("alm and map must both be host arrays or both be CUDA tensors")
	if dta == np.complex128 and dtm == np.float64: dtype = L.F64
	elif dta == np.complex64 and dtm == np.float32: dtype = L.F32
	else: raise ValueError("alm/map dtypes must be (complex128, float64) or (complex64, float32), got (%s, %s)" % (dta, dtm))
	ncm = 1 if spin == 0 else 2
	if alm.ndim != 3 or alm.shape[1] != ncm or map.ndim < 3 or map.shape[1] != ncm or map.shape[0] != alm.shape[0]:
		raise ValueError("run_batch: alm must be [nb, %d, nalm] and map [nb, %d, ...] for spin %d" % (ncm, ncm, spin))
	_check_last_contig(alm, "alm", 1); _check_last_contig(map, "map", map.ndim-2)
	sa, sm = L.strides_elems(alm), L.strides_elems(map)
	L.check(L.lib().b2_synthesis(plan.handle, int(spin), L.MODE_STANDARD, dtype, int(alm.shape[0]), pa, sa[1] if ncm > 1 else 0, sa[0],
		pm, sm[1] if ncm > 1 else 0, sm[0], mema, L.current_stream(map)))

def _alm_len(mstart, lmax, lstride): return int(np.max(mstart) + lmax*lstride + 1)

# ------------------------------------------------------------------ ducc0.sht.experimental look-alikes

def synthesis(*, alm, theta, nphi, phi0, ringstart, spin, lmax, mmax=None, mstart=None, lstride=1, pixstride=1,
		map=None, mode="STANDARD", nthreads=0, weight=None, xdir=1, npix=None, **kw):
	"""ducc0.sht.experimental.synthesis: alm[nca, nalm] -> map[ncm, npix_total] (filled in place and returned)."""
	if pixstride != 1: raise NotImplementedError("pixstride != 1")
	plan = plan_rings(theta, nphi, phi0, ringstart, lmax, mmax, mstart, lstride, weight, xdir, npix)
	md = _MODES[mode]
	if map is None:
		ncm = 1 if spin == 0 else 2
		n = int(np.max(np.asarray(ringstart).astype(np.int64) + (np.broadcast_to(np.asarray(nphi).astype(np.int64), np.shape(ringstart)) if npix is None else npix)))
		map = _empty_like(alm, (ncm, n), np.float64 if L.buffer_info(alm)[2] == np.complex128 else np.float32)
	_run(L.lib().b2_synthesis, plan, spin, md, alm, map, 1)
	return map

def adjoint_synthesis(*, map, theta, nphi, phi0, ringstart, spin, lmax, mmax=None, mstart=None, lstride=1, pixstride=1,
		alm=None, mode="STANDARD", nthreads=0, weight=None, xdir=1, npix=None, **kw):
	"""ducc0.sht.experimental.adjoint_synthesis: map -> alm = Y^T (w map)."""
	if pixstride != 1: raise NotImplementedError("pixstride != 1")
	plan = plan_rings(theta, nphi, phi0, ringstart, lmax, mmax, mstart, lstride, weight, xdir, npix)
	md = _MODES[mode]
	if alm is None:
		lm, mm, ms, ls = _layout(lmax, mmax, mstart, lstride)
		nca = 1 if (spin == 0 or md == L.MODE_DERIV1) else 2
		alm = _empty_like(map, (nca, _alm_len(ms, lm, ls)), np.complex128 if L.buffer_info(map)[2] == np.float64 else np.complex64)
	_run(L.lib().b2_adjoint_synthesis, plan, spin, md, alm, map, 1)
	return alm

def _plan_for_2d(map, ntheta, nphi, geometry, phi0, lmax, mmax, mstart, lstride, flip_y, flip_x):
	if map is not None: ntheta, nphi = map.shape[-2:]
	if ntheta is None or nphi is None: raise ValueError("need map or ntheta/nphi")
	return plan_2d(geometry, ntheta, nphi, phi0, lmax, mmax, mstart, lstride, flip_y, flip_x), int(ntheta), int(nphi)

def synthesis_2d(*, alm, spin, lmax, geometry, ntheta=None, nphi=None, mmax=None, mstart=None, lstride=1,
		phi0=0.0, map=None, mode="STANDARD", nthreads=0, flip_y=False, flip_x=False, **kw):
	"""ducc0.sht.experimental.synthesis_2d (pixell/curvedsky.py:908)."""
	plan, ntheta, nphi = _plan_for_2d(map, ntheta, nphi, geometry, phi0, lmax, mmax, mstart, lstride, flip_y, flip_x)
	if map is None:
		map = _empty_like(alm, (1 if spin == 0 else 2, ntheta, nphi), np.float64 if L.buffer_info(alm)[2] == np.complex128 else np.float32)
	_run(L.lib().b2_synthesis, plan, spin, _MODES[mode], alm, map, 2)
	return map

def adjoint_synthesis_2d(*, map, spin, lmax, geometry, mmax=None, mstart=None, lstride=1, phi0=0.0, alm=None,
		mode="STANDARD", nthreads=0, flip_y=False, flip_x=False, **kw):
	"""ducc0.sht.experimental.adjoint_synthesis_2d (pixell/curvedsky.py:907)."""
	plan, ntheta, nphi = _plan_for_2d(map, None, None, geometry, phi0, lmax, mmax, mstart, lstride, flip_y, flip_x)
	md = _MODES[mode]
	if alm is None:
		lm, mm, ms, ls = _layout(lmax, mmax, mstart, lstride)
		nca = 1 if (spin == 0 or md == L.MODE_DERIV1) else 2
		alm = _empty_like(map, (nca, _alm_len(ms, lm, ls)), np.complex128 if L.buffer_info(map)[2] == np.float64 else np.complex64)
	_run(L.lib().b2_adjoint_synthesis, plan, spin, md, alm, map, 2)
	return alm

def analysis_2d(*, map, spin, lmax, geometry, mmax=None, mstart=None, lstride=1, phi0=0.0, alm=None,
		nthreads=0, flip_y=False, flip_x=False, **kw):
	"""ducc0.sht.experimental.analysis_2d (pixell/curvedsky.py:1033): exact inverse of synthesis_2d
	for band-limited maps."""
	plan, ntheta, nphi = _plan_for_2d(map, None, None, geometry, phi0, lmax, mmax, mstart, lstride, flip_y, flip_x)
	if alm is None:
		lm, mm, ms, ls = _layout(lmax, mmax, mstart, lstride)
		alm = _empty_like(map, (1 if spin == 0 else 2, _alm_len(ms, lm, ls)), np.complex128 if L.buffer_info(map)[2] == np.float64 else np.complex64)
	_run(L.lib().b2_analysis_2d, plan, spin, L.MODE_STANDARD, alm, map, 2, has_mode=False)
	return alm

def adjoint_analysis_2d(*, alm, spin, lmax, geometry, ntheta=None, nphi=None, mmax=None, mstart=None, lstride=1,
		phi0=0.0, map=None, nthreads=0, flip_y=False, flip_x=False, **kw):
	"""ducc0.sht.experimental.adjoint_analysis_2d (pixell/curvedsky.py:1032)."""
	plan, ntheta, nphi = _plan_for_2d(map, ntheta, nphi, geometry, phi0, lmax, mmax, mstart, lstride, flip_y, flip_x)
	if map is None:
		map = _empty_like(alm, (1 if spin == 0 else 2, ntheta, nphi), np.float64 if L.buffer_info(alm)[2] == np.complex128 else np.float32)
	_run(L.lib().b2_adjoint_analysis_2d, plan, spin, L.MODE_STANDARD, alm, map, 2, has_mode=False)
	return map

def get_gridweights(geometry, ntheta):
	"""ducc0.sht.experimental.get_gridweights (pixell/curvedsky.py:501, 531, 855)."""
	L.init()
	out = np.zeros(int(ntheta), np.float64)
	L.check(L.lib().b2_gridweights(geometry.encode(), int(ntheta), L.p_dbl(out)), ValueError)
	return out

def maxlmax(geometry, ny):
	"""pixell/curvedsky.py:1349-1353 (get_ducc_maxlmax)"""
	if   geometry == "CC": return ny-2
	elif geometry == "DH": return (ny-2)//2
	elif geometry == "F2": return (ny-1)//2
	else:                  return ny-1

# ------------------------------------------------------------------ arbitrary positions (ducc0 synthesis_general)

GENERAL_W, GENERAL_BETA = 13, 2.30*13        # interpolation kernel exp(beta (sqrt(1 - z^2) - 1)), |z| <= 1, over W grid points

def _fast_len(n):
	"""smallest even length >= n with prime factors 2, 3, 5 only (the FFT engine's register-butterfly path)"""
	n = int(n) + (int(n) & 1)
	while True:
		k = n
		for p in (2, 3, 5):
			while k % p == 0: k //= p
		if k == 1: return n
		n += 2

def _kernel_corr(lmax, M, W=GENERAL_W, beta=GENERAL_BETA):
	"""1/P_k, k = 0..lmax: P_k = (W/2) int_{-1}^{1} exp(beta (sqrt(1-z^2) - 1)) cos(k W pi z / M) dz is what the
	periodic sum of the interpolation kernel multiplies mode k with (Gauss-Legendre quadrature, 240 nodes)"""
	z, w = np.polynomial.legendre.leggauss(240)
	phi = np.exp(beta*(np.sqrt(1-z*z)-1))
	k = np.arange(lmax+1)[:, None]
	P = 0.5*W*np.sum(w*phi*np.cos(k*(W*np.pi/M)*z), 1)
	return 1.0/P

def synthesis_general(*, alm, loc, spin, lmax, mmax=None, mstart=None, lstride=1, epsilon=1e-10, map=None, mode="STANDARD",
		nthreads=0, **kw):
	"""ducc0.sht.experimental.synthesis_general: alm[nca, nalm] -> map[ncm, npos] at loc[npos, 2] = (theta, phi) radians.
	alm / loc / map: numpy arrays or torch CUDA tensors (complex128 / float64).  The kernel width is fixed (W = 13,
	about 1e-12), so any epsilon >= 1e-12 is honoured."""
	import torch
	from . import fft as enfft
	if epsilon < 1e-12: raise ValueError("synthesis_general: epsilon below 1e-12 is not supported")
	md = _MODES[mode]
	lmax, mmax, mstart, lstride = _layout(lmax, mmax, mstart, lstride)
	ncm = 1 if spin == 0 else 2
	nca = 1 if (spin == 0 or md == L.MODE_DERIV1) else 2
	if alm.ndim != 2 or alm.shape[0] != nca: raise ValueError("alm must have shape [%d, nalm] for spin %d" % (nca, spin))
	adt = L.buffer_info(alm)[2]
	if adt not in (np.complex128, np.complex64): raise ValueError("synthesis_general: alm must be complex128 or complex64")
	dev = torch.device("cuda", L.init())
	host = not L.is_torch(alm)
	# single precision rides on the double-precision pipeline (ducc accepts complex64 alm with float32 maps here)
	talm = torch.from_numpy(np.ascontiguousarray(alm, dtype=np.complex128)).to(dev) if host else alm.to(torch.complex128).contiguous()
	tloc = (torch.from_numpy(np.ascontiguousarray(loc, dtype=np.float64)) if not L.is_torch(loc) else loc).to(dev).contiguous()
	if tloc.ndim != 2 or tloc.shape[1] != 2 or tloc.dtype != torch.float64: raise ValueError("loc must be float64 [npos, 2] = (theta, phi)")
	npos = tloc.shape[0]
	# Legendre stage on a Clenshaw-Curtis ring set whose doubled circle has a fast FFT length
	N = _fast_len(2*lmax+2); nt = N//2+1
	M = _fast_len(max(2*(2*lmax+2), 4*GENERAL_W))      # oversampled grid; never smaller than the interpolation kernel needs
	plan = plan_2d("CC", nt, N, 0.0, lmax, mmax, mstart, lstride)
	nring_pad = (nt+31)//32*32
	nm = mmax+1
	lib = L.lib(); st = torch.cuda.current_stream(dev).cuda_stream
	leg = torch.empty((ncm, nm, nring_pad), dtype=torch.complex128, device=dev)
	L.check(lib.b2_alm2leg(plan.handle, int(spin), md, talm.data_ptr(), int(talm.stride(0)) if nca > 1 else 0, leg.data_ptr(), st))
	ext = torch.empty((ncm, nm, N), dtype=torch.complex128, device=dev)
	L.check(lib.b2_general_extend(leg.data_ptr(), ext.data_ptr(), ncm, nm, nt, nring_pad, int(spin), st))
	del leg
	coef = torch.empty_like(ext)                                   # (lines too long for one CTA are not transformed in place)
	enfft.transform(ext, coef, (-1,), True, 1.0/N)                 # theta Fourier coefficients c[c][m][k]
	del ext
	corr = torch.from_numpy(_kernel_corr(lmax, M)).to(dev)
	grid = torch.empty((ncm, M, M//2+1), dtype=torch.complex128, device=dev)
	L.check(lib.b2_general_scatter(coef.data_ptr(), grid.data_ptr(), ncm, lmax, nm, N, M, corr.data_ptr(), st))
	del coef
	fine = torch.empty((ncm, M, M), dtype=torch.float64, device=dev)
	enfft.transform(grid, fine, (-2, -1), False, 1.0)            # unnormalised inverse: the Fourier series on the M x M grid
	del grid
	direct = not (host or map is None or not L.is_torch(map)) and map.dtype == torch.float64
	tout = map if direct else torch.empty((ncm, npos), dtype=torch.float64, device=dev)
	if tout.shape != (ncm, npos) or tout.stride(-1) != 1: raise ValueError("map must have shape [%d, npos] with a contiguous last axis" % ncm)
	L.check(lib.b2_general_interp(fine.data_ptr(), ncm, M, tloc.data_ptr(), npos, GENERAL_W, GENERAL_BETA, tout.data_ptr(), int(tout.stride(0)), st))
	if map is None:
		if adt == np.complex64: tout = tout.to(torch.float32)
		return tout.cpu().numpy() if host else tout
	if not L.is_torch(map): map[...] = tout.cpu().numpy().astype(map.dtype, copy=False)
	elif not direct: map.copy_(tout)
	return map

def adjoint_synthesis_general(*, map, loc, spin, lmax, mmax=None, mstart=None, lstride=1, epsilon=1e-10, alm=None, mode="STANDARD",
		nthreads=0, **kw):
	"""ducc0.sht.experimental.adjoint_synthesis_general: map[ncm, npos] at loc[npos, 2] -> alm[nca, nalm] = Y^T map, the
	exact transpose of synthesis_general step by step (spread, forward FFT, gather, inverse theta-FFT, fold, K2)."""
	import torch
	from . import fft as enfft
	if epsilon < 1e-12: raise ValueError("adjoint_synthesis_general: epsilon below 1e-12 is not supported")
	md = _MODES[mode]
	lmax, mmax, mstart, lstride = _layout(lmax, mmax, mstart, lstride)
	ncm = 1 if spin == 0 else 2
	nca = 1 if (spin == 0 or md == L.MODE_DERIV1) else 2
	if map.ndim != 2 or map.shape[0] != ncm: raise ValueError("map must have shape [%d, npos] for spin %d" % (ncm, spin))
	dev = torch.device("cuda", L.init())
	host = not L.is_torch(map)
	tmap = (torch.from_numpy(np.ascontiguousarray(map, dtype=np.float64)).to(dev) if host else map.to(torch.float64)).contiguous()
	tloc = (torch.from_numpy(np.ascontiguousarray(loc, dtype=np.float64)) if not L.is_torch(loc) else loc).to(dev).contiguous()
	if tloc.ndim != 2 or tloc.shape[1] != 2 or tloc.shape[0] != tmap.shape[1]: raise ValueError("loc must be float64 [npos, 2] = (theta, phi)")
	npos = tloc.shape[0]
	N = _fast_len(2*lmax+2); nt = N//2+1
	M = _fast_len(max(2*(2*lmax+2), 4*GENERAL_W))      # oversampled grid; never smaller than the interpolation kernel needs
	plan = plan_2d("CC", nt, N, 0.0, lmax, mmax, mstart, lstride)
	nring_pad = (nt+31)//32*32
	nm = mmax+1
	lib = L.lib(); st = torch.cuda.current_stream(dev).cuda_stream
	fine = torch.empty((ncm, M, M), dtype=torch.float64, device=dev)
	L.check(lib.b2_general_spread(fine.data_ptr(), ncm, M, tloc.data_ptr(), npos, GENERAL_W, GENERAL_BETA, tmap.data_ptr(), int(tmap.stride(0)), st))
	grid = torch.empty((ncm, M, M//2+1), dtype=torch.complex128, device=dev)
	enfft.transform(fine, grid, (-2, -1), True, 1.0)
	del fine
	corr = torch.from_numpy(_kernel_corr(lmax, M)).to(dev)
	coef = torch.empty((ncm, nm, N), dtype=torch.complex128, device=dev)
	L.check(lib.b2_general_gather(coef.data_ptr(), grid.data_ptr(), ncm, lmax, nm, N, M, corr.data_ptr(), st))
	del grid
	ext = torch.empty_like(coef)
	enfft.transform(coef, ext, (-1,), False, 1.0/N)
	del coef
	leg = torch.empty((ncm, nm, nring_pad), dtype=torch.complex128, device=dev)
	L.check(lib.b2_general_fold(leg.data_ptr(), ext.data_ptr(), ncm, nm, nt, nring_pad, int(spin), st))
	del ext
	nalm = _alm_len(mstart, lmax, lstride)
	talm = torch.zeros((nca, nalm), dtype=torch.complex128, device=dev)
	L.check(lib.b2_leg2alm(plan.handle, int(spin), md, talm.data_ptr(), nalm if nca > 1 else 0, leg.data_ptr(), st))
	if alm is None:
		if L.buffer_info(map)[2] == np.float32: talm = talm.to(torch.complex64)
		return talm.cpu().numpy() if host else talm
	if L.is_torch(alm): alm.copy_(talm)
	else: alm[...] = talm.cpu().numpy().astype(alm.dtype, copy=False)
	return alm
