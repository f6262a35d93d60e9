"""pixell_b200.fft -- the pixell.fft interface (reference pixell/fft.py) on the B200 FFT engine.

Two layers, both backed by b2_fft_plan_create / b2_fft_execute of libb200sht.so:
  * `engine`: an object with the plug-in shape the reference's `fft.engines` registry expects
    (pixell/fft.py:8-113): `engine.FFTW(a, b, axes=(-1,), direction='FFTW_FORWARD', threads=1, flags=...)`
    returns a plan object; calling it (`plan(normalise_idft=False)`) fills `b` in place;
    `engine.empty_aligned(shape, dtype, n=None)`.  Register it with
        pixell.fft.engines["b200"] = pixell_b200.fft.engine; pixell.fft.set_engine("b200")
    (INTEGRATION.md).  The transform kind is inferred from shapes and dtypes exactly as the reference's
    numpy / ducc engines do: equal shapes -> c2c, else r2c (forward) / c2r (backward).
  * `fft, ifft, rfft, irfft` with the reference's signatures and conventions (:133-209): forward
    unnormalised, backward unnormalised unless normalize=True, output allocated when not given.
Arrays may be numpy arrays (host; staged through the GPU, strided views allowed) or torch CUDA tensors
(zero-copy).  nthread / flags are accepted and ignored.  There is no CPU fallback.
"""
import collections, ctypes
import numpy as np
from . import _lib as L

_plans = collections.OrderedDict()
PLAN_CACHE_SIZE = 8

class _Plan:
	def __init__(self, handle): self.handle = handle
	def __del__(self):
		try:
			if self.handle: L.lib().b2_fft_plan_destroy(self.handle); self.handle = None
		except Exception: pass

def clear_plans(): _plans.clear()

def _get_plan(shape, istride, ostride, axes, kind, dtype):
	dev = L.init()
	key = (dev, tuple(shape), tuple(istride), tuple(ostride), tuple(axes), kind, dtype)
	p = _plans.get(key)
	if p is None:
		while len(_plans) >= PLAN_CACHE_SIZE: _plans.popitem(last=False)
		h = ctypes.c_void_p()
		sh, ist, ost = L.as_i64(shape), L.as_i64(istride), L.as_i64(ostride)
		ax = np.ascontiguousarray(np.asarray(axes, dtype=np.int32))
		L.check(L.lib().b2_fft_plan_create(ctypes.byref(h), len(shape), L.p_i64(sh), L.p_i64(ist), L.p_i64(ost),
			len(axes), ax.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), kind, dtype), ValueError)
		p = _Plan(h); _plans[key] = p
	else: _plans.move_to_end(key)
	return p

def _collapse(shape, ist, ost, axes):
	"""Merge leading non-transform dimensions until at most 4 remain (the engine's limit) when their strides allow it."""
	shape, ist, ost, axes = list(shape), list(ist), list(ost), sorted(axes)
	d = 0
	while len(shape) > 4 and d+1 < len(shape):
		if d not in axes and d+1 not in axes and ist[d] == ist[d+1]*shape[d+1] and ost[d] == ost[d+1]*shape[d+1]:
			shape[d+1] *= shape[d]; del shape[d], ist[d], ost[d]
			axes = [a-1 if a > d else a for a in axes]
		else: d += 1
	return shape, ist, ost, axes

def transform(a, b, axes, forward, scale=1.0):
	"""b = DFT(a) over `axes` (kind inferred like the reference engines).  a, b: numpy arrays or torch CUDA tensors."""
	pa, mema, dta = L.buffer_info(a); pb, memb, dtb = L.buffer_info(b)
	if mema != memb: raise ValueError("fft: input and output must both be host arrays or both be CUDA tensors")
	nd = a.ndim
	if nd == 0: raise ValueError("fft: zero-dimensional input")
	axes = [ax+nd if ax < 0 else ax for ax in (list(axes) if np.ndim(axes) else [axes])]
	if len(axes) > 2: raise NotImplementedError("pixell_b200.fft transforms at most two axes at a time")
	ca, cb = dta.kind == "c", dtb.kind == "c"
	if tuple(a.shape) == tuple(b.shape) and ca and cb: kind, full = L.FFT_C2C, a.shape
	elif not ca and cb and forward: kind, full = L.FFT_R2C, a.shape
	elif ca and not cb and not forward: kind, full = L.FFT_C2R, b.shape
	else: raise ValueError("fft: cannot infer the transform from shapes %s -> %s and dtypes %s -> %s" % (a.shape, b.shape, dta, dtb))
	half = list(full); half[axes[-1]] = full[axes[-1]]//2+1
	if kind == L.FFT_R2C and tuple(b.shape) != tuple(half): raise ValueError("fft: r2c output must have shape %s" % (tuple(half),))
	if kind == L.FFT_C2R and tuple(a.shape) != tuple(half): raise ValueError("fft: c2r input must have shape %s" % (tuple(half),))
	prec = {8: L.F64, 4: L.F32}[dta.itemsize//(2 if ca else 1)]
	if dtb.itemsize//(2 if cb else 1) != dta.itemsize//(2 if ca else 1): raise ValueError("fft: input and output precision differ")
	ist, ost = L.strides_elems(a), L.strides_elems(b)
	if any(s < 0 for s in ist) or any(s < 0 for s in ost): raise ValueError("fft: negative strides are not supported")
	shape, ist, ost, axes2 = _collapse(full, ist, ost, axes)
	if len(shape) > 4: raise NotImplementedError("fft: more than 4 non-mergeable dimensions")
	# keep the caller's axis order (the last listed axis is the real one)
	order = [sorted(axes).index(ax) for ax in axes]
	axes2 = [axes2[i] for i in order]
	plan = _get_plan(shape, ist, ost, axes2, kind, prec)
	stream = L.current_stream(b)
	L.check(L.lib().b2_fft_execute(plan.handle, pa, pb, 1 if forward else 0, float(scale), mema, stream))
	return b

# ------------------------------------------------------------------ engine plug-in (pixell/fft.py:8-113)

class FFTW:
	"""Plan object with the call shape of pyfftw.FFTW / the reference's numpy_FFTW and ducc_FFTW classes."""
	def __init__(self, a, b, axes=(-1,), direction="FFTW_FORWARD", threads=1, flags=None, *args, **kwargs):
		self.a, self.b = a, b
		self.axes = tuple(axes) if np.ndim(axes) else (axes,)
		if not isinstance(direction, str):
			# r2r: one FFTW kind per axis (pixell/fft.py:66-71, 211-259)
			direction = list(direction)
			if len(direction) != len(self.axes) or any(d not in _R2R for d in direction): raise ValueError("unknown r2r direction %s" % (direction,))
		elif direction not in ("FFTW_FORWARD", "FFTW_BACKWARD"): raise ValueError("unknown direction %s" % direction)
		self.direction = direction
	def __call__(self, normalise_idft=False):
		if not isinstance(self.direction, str): return r2r(self.a, self.b, self.axes, self.direction)
		fwd = self.direction == "FFTW_FORWARD"
		scale = 1.0
		if not fwd and normalise_idft:
			out_shape = self.b.shape
			scale = 1.0/np.prod([out_shape[ax] for ax in self.axes])
		transform(self.a, self.b, self.axes, fwd, scale)
		return self.b

def empty_aligned(shape, dtype, n=None):
	return np.empty(shape, dtype)

class _Engine: pass
engine = _Engine()
engine.FFTW = FFTW
engine.empty_aligned = empty_aligned

def register(pixell_fft_module, name="b200", select=True):
	"""pixell.fft.engines[name] = engine (and select it): the whole integration on the reference side."""
	pixell_fft_module.engines[name] = engine
	if select: pixell_fft_module.set_engine(name)

# ------------------------------------------------------------------ pixell.fft functions (:133-209)

def _astuple(x): return tuple(x) if np.ndim(x) else (x,)

def _asfc(a):
	if L.is_torch(a): return a
	a = np.asarray(a)
	return np.asarray(a, np.result_type(a, 0.0))

def _empty_like(ref, shape, dtype):
	if L.is_torch(ref):
		import torch
		td = {np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32,
			np.dtype(np.complex128): torch.complex128, np.dtype(np.complex64): torch.complex64}[np.dtype(dtype)]
		return torch.empty(tuple(shape), dtype=td, device=ref.device)
	return np.empty(tuple(shape), dtype)

def _dtype(a): return L.buffer_info(a)[2]
def _size(a): return int(np.prod(a.shape))

def fft(tod, ft=None, nthread=0, axes=[-1], flags=None, _direction="FFTW_FORWARD", engine="auto"):
	"""pixell/fft.py:133-160"""
	tod = _asfc(tod)
	axes = _astuple(-1 if axes is None else axes)
	if _size(tod) == 0: return
	if not isinstance(_direction, str):
		if ft is None: ft = _empty_like(tod, tod.shape, _dtype(tod))
		return r2r(tod, ft, axes, list(_direction))
	if ft is None:
		otype = np.result_type(_dtype(tod), 0j)
		ft = _empty_like(tod, tod.shape, otype)
		tod = tod.to(ft.dtype) if L.is_torch(tod) else tod.astype(otype, copy=False)
	return transform(tod, ft, axes, _direction == "FFTW_FORWARD")

def ifft(ft, tod=None, nthread=0, normalize=False, axes=[-1], flags=None, engine="auto"):
	"""pixell/fft.py:162-187 (normalisation is fused into the last pass instead of a separate division)"""
	ft = _asfc(ft)
	axes = _astuple(-1 if axes is None else axes)
	if _size(ft) == 0: return
	if tod is None: tod = _empty_like(ft, ft.shape, _dtype(ft))
	scale = 1.0/np.prod([tod.shape[i] for i in axes]) if normalize else 1.0
	return transform(ft, tod, axes, False, scale)

def rfft_shape(ishape, axes=[-1]):
	oshape = list(ishape); oshape[axes[-1]] = oshape[axes[-1]]//2+1
	return oshape

def irfft_shape(ishape, n=None, axes=[-1]):
	oshape = list(ishape); oshape[axes[-1]] = (oshape[axes[-1]]-1)*2 if n is None else n
	return oshape

def rfft(tod, ft=None, nthread=0, axes=[-1], flags=None, engine="auto"):
	"""pixell/fft.py:189-198"""
	tod = _asfc(tod)
	axes = _astuple(-1 if axes is None else axes)
	if ft is None: ft = _empty_like(tod, rfft_shape(tod.shape, axes=axes), np.result_type(_dtype(tod), 0j))
	return fft(tod, ft, nthread, axes, flags=flags)

def irfft(ft, tod=None, n=None, nthread=0, normalize=False, axes=[-1], flags=None, engine="auto"):
	"""pixell/fft.py:200-213"""
	ft = _asfc(ft)
	axes = _astuple(-1 if axes is None else axes)
	if tod is None: tod = _empty_like(ft, irfft_shape(ft.shape, axes=axes, n=n), np.zeros([], _dtype(ft)).real.dtype)
	return ifft(ft, tod, nthread, normalize, axes, flags=flags)

# ------------------------------------------------------------------ r2r: DCT / DST (pixell/fft.py:211-317)
# FFTW's eight real-to-real kinds, unnormalised as FFTW defines them.  Every kind is a cosine (sine) sum
#   Y_k = sum_j c_j X_j cos|sin(2 pi P_j Q_k / Lc)   with integer P_j, Q_k,
# i.e. entries Q_k of the DFT of a real sequence of length Lc that holds X_j at P_j and its even (odd) image at Lc - P_j.
# So one r2c transform of the engine per axis does the work (length 2(n-1) ... 8n: these transforms are not on the hot
# path, pixell uses them for enmap.fft(dct=True) and the Chebyshev helpers only).
#             name:          (sine, Lc(n),              P(j),          Q(k))
_R2R = {
	"FFTW_REDFT00": (False, lambda n: 2*(n-1), lambda j: j,     lambda k: k),
	"FFTW_REDFT10": (False, lambda n: 4*n,     lambda j: 2*j+1, lambda k: k),
	"FFTW_REDFT01": (False, lambda n: 4*n,     lambda j: j,     lambda k: 2*k+1),
	"FFTW_REDFT11": (False, lambda n: 8*n,     lambda j: 2*j+1, lambda k: 2*k+1),
	"FFTW_RODFT00": (True,  lambda n: 2*(n+1), lambda j: j+1,   lambda k: k+1),
	"FFTW_RODFT10": (True,  lambda n: 4*n,     lambda j: 2*j+1, lambda k: k+1),
	"FFTW_RODFT01": (True,  lambda n: 4*n,     lambda j: j+1,   lambda k: 2*k+1),
	"FFTW_RODFT11": (True,  lambda n: 8*n,     lambda j: 2*j+1, lambda k: 2*k+1),
}
_dct_names = {"DCT-I": "FFTW_REDFT00", "DCT-II": "FFTW_REDFT10", "DCT-III": "FFTW_REDFT01", "DCT-IV": "FFTW_REDFT11",
	"DST-I": "FFTW_RODFT00", "DST-II": "FFTW_RODFT10", "DST-III": "FFTW_RODFT01", "DST-IV": "FFTW_RODFT11"}
_dct_names.update({v: v for v in list(_dct_names.values())})
_dct_inverses = {"FFTW_REDFT00": "FFTW_REDFT00", "FFTW_REDFT10": "FFTW_REDFT01", "FFTW_REDFT01": "FFTW_REDFT10", "FFTW_REDFT11": "FFTW_REDFT11",
	"FFTW_RODFT00": "FFTW_RODFT00", "FFTW_RODFT10": "FFTW_RODFT01", "FFTW_RODFT01": "FFTW_RODFT10", "FFTW_RODFT11": "FFTW_RODFT11"}
_dct_sizes = {"FFTW_REDFT00": -1, "FFTW_RODFT00": +1}

def _r2r_last(x, kind):
	"""one r2r kind along the last axis of a contiguous real torch CUDA tensor [batch, n]"""
	import torch
	sine, Lc, P, Q = _R2R[kind]
	nb, n = x.shape
	if kind == "FFTW_REDFT00" and n < 2: raise ValueError("DCT-I needs at least two points")
	L_ = Lc(n)
	j = torch.arange(n, device=x.device)
	p = P(j); mir = (L_ - p) % L_
	z = torch.zeros((nb, L_), dtype=x.dtype, device=x.device)
	if kind == "FFTW_RODFT01":
		x = x.clone(); x[:, -1] *= 0.5          # FFTW counts the last input once: (-1)^k X_{n-1}
	z[:, p] = x
	img = mir != p                               # points that are their own image (DCT-I / DCT-III end points) enter once
	z[:, mir[img]] = -x[:, img] if sine else x[:, img]
	Z = torch.empty((nb, L_//2+1), dtype=torch.complex128 if x.dtype == torch.float64 else torch.complex64, device=x.device)
	transform(z, Z, [-1], True)
	q = Q(j)
	return -Z[:, q].imag if sine else Z[:, q].real

def r2r(a, b, axes, kinds):
	"""b = the r2r transform of kind kinds[i] along axes[i] of a (FFTW_REDFTxx / FFTW_RODFTxx), unnormalised"""
	import torch
	if tuple(a.shape) != tuple(b.shape): raise ValueError("r2r: input and output shapes differ")
	dt = _dtype(a)
	if dt.kind != "f" or _dtype(b) != dt: raise ValueError("r2r transforms map real arrays to real arrays of the same precision")
	dev = L.init()
	x = a if L.is_torch(a) else torch.from_numpy(np.ascontiguousarray(a)).to("cuda:%d" % dev)
	nd = x.ndim
	for ax, kind in zip(axes, kinds):
		ax = ax+nd if ax < 0 else ax
		xm = x.movedim(ax, -1)
		sh = xm.shape
		y = _r2r_last(xm.reshape(-1, sh[-1]).contiguous(), kind)
		x = y.reshape(sh).movedim(-1, ax)
	if L.is_torch(b): b.copy_(x)
	else: b[...] = x.cpu().numpy()
	return b

def dct(tod, dt=None, nthread=0, normalize=False, axes=[-1], flags=None, type="DCT-I", engine="auto"):
	"""pixell/fft.py:211-231"""
	tod = _asfc(tod)
	kind = _dct_names[type]
	axes = _astuple(-1 if axes is None else axes)
	if dt is None: dt = _empty_like(tod, tod.shape, _dtype(tod))
	return fft(tod, dt, nthread=nthread, axes=axes, flags=flags, _direction=[kind]*len(axes))

def idct(dt, tod=None, nthread=0, normalize=False, axes=[-1], flags=None, type="DCT-I", engine="auto"):
	"""pixell/fft.py:233-265: the matching inverse kind; normalize=True divides by prod 2 (N + d)"""
	dt = _asfc(dt)
	kind = _dct_inverses[_dct_names[type]]
	off = _dct_sizes.get(kind, 0)
	axes = _astuple(-1 if axes is None else axes)
	if tod is None: tod = _empty_like(dt, dt.shape, _dtype(dt))
	fft(dt, tod, nthread=nthread, axes=axes, flags=flags, _direction=[kind]*len(axes))
	if normalize: tod /= float(np.prod([2*(tod.shape[i]+off) for i in axes]))
	return tod

def redft00(a, b=None, nthread=0, normalize=False, flags=None, engine="auto"):
	"""pixell/fft.py:290-305 (DCT-I along the last axis)"""
	a = _asfc(a)
	if b is None: b = _empty_like(a, a.shape, _dtype(a))
	r2r(a, b, [-1], ["FFTW_REDFT00"])
	if normalize: b /= 2*(a.shape[-1]-1)
	return b

def chebt(a, b=None, nthread=0, flags=None, engine="auto"):
	"""pixell/fft.py:307-311: the Chebyshev transform of a along its last dimension (the reference scales b[1:-1] along the
	FIRST axis of b: for one-dimensional input that is the same thing; mirrored as written)"""
	b = redft00(a, b, nthread, normalize=True, flags=flags)
	b[1:-1] *= 2
	return b

def ichebt(a, b=None, nthread=0, engine="auto"):
	"""pixell/fft.py:313-317"""
	a = _asfc(a)
	a = a.clone() if L.is_torch(a) else a.copy()
	a[1:-1] *= 0.5
	return redft00(a, b, nthread)

def fourier_filter(ft, fy=None, fx=None, f2=None):
	"""ft[..., y, x] *= fy[y]*fx[x] (or *= f2[y, x]) in place on the device, one pass (b2_fourier_filter): the harmonic
	filter between a forward and a backward 2-D transform.  ft: complex torch CUDA tensor, contiguous in x, uniform row
	and batch strides; the filters: real arrays / tensors of the matching precision"""
	import torch
	if not L.is_torch(ft): raise ValueError("fourier_filter works on device tensors (use numpy broadcasting for host arrays)")
	ny, nx = ft.shape[-2:]
	nb = int(np.prod(ft.shape[:-2])) if ft.ndim > 2 else 1
	st = L.strides_elems(ft)
	if st[-1] != 1: raise ValueError("fourier_filter: the last axis must be contiguous")
	bs = st[-3] if ft.ndim > 2 else 0
	if ft.ndim > 3 and not ft.reshape(nb, ny, nx).data_ptr() == ft.data_ptr(): raise ValueError("fourier_filter: leading axes must be mergeable")
	if ft.ndim > 3:
		ft3 = ft.view(nb, ny, nx); bs = L.strides_elems(ft3)[0]
	rdt = torch.float64 if ft.dtype == torch.complex128 else torch.float32
	dev = ft.device
	def prep(a, shape):
		if a is None: return None
		t = torch.as_tensor(a, device=dev).to(rdt).contiguous()
		if tuple(t.shape) != shape: raise ValueError("fourier_filter: filter has shape %s, expected %s" % (tuple(t.shape), shape))
		return t
	ty, tx, t2 = prep(fy, (ny,)), prep(fx, (nx,)), prep(f2, (ny, nx))
	L.check(L.lib().b2_fourier_filter(ft.data_ptr(), nb, bs, st[-2], ny, nx, None if ty is None else ty.data_ptr(), None if tx is None else tx.data_ptr(),
		None if t2 is None else t2.data_ptr(), L.F64 if rdt == torch.float64 else L.F32, L.current_stream(ft)))
	return ft

# ------------------------------------------------------------------ host helpers of pixell.fft (:319-433)

def fft_len(n, direction="below", factors=None):
	"""pixell/fft.py:319-321: the largest product of powers of `factors` not above n ("below"), or the smallest one
	not below n ("above")"""
	factors = [2, 3, 5, 7, 11, 13] if factors is None else [int(f) for f in factors]
	below = direction == "below"
	target = int(np.floor(n)) if below else int(np.ceil(n))
	if 1 in factors: return target
	# The reference walks the smooth numbers i = 1, 2, ... in ascending order and, for "below", keeps the LAST product i*f <= n
	# it meets (factors in the order given) -- not the largest: fft_len(7) is 6, because 2*3 is met after 1*7.  Mirrored as
	# it behaves (callers size their arrays with it); "above" keeps the smallest product >= n.
	top = target if below else max(target, 1)*min(factors)
	smooth = np.zeros(top+2, bool); smooth[1] = True
	best = None
	for i in range(1, target+1):
		if not smooth[i]: continue
		for f in factors:
			m = i*f
			if below:
				if m <= n: best = m
			elif m >= n and (best is None or m < best): best = m
			if m <= top: smooth[m] = True
	return best

def asfcarray(a):
	"""pixell/fft.py:323-325"""
	return _asfc(a)

def empty(shape, dtype):
	"""pixell/fft.py:327-328"""
	return empty_aligned(shape, dtype)

def ind2freq(n, i, d=1.0): return np.where(np.asarray(i) < n/2, i, -n+np.asarray(i))/(d*n)
def int2rfreq(n, i, d=1.0): return np.asarray(i)/(n*d)
def freq2ind(n, f, d=1.0):
	j = np.asarray(f)*(d*n)
	return np.where(j >= 0, j, n+j)
def rfreq2ind(n, f, d=1.0): return np.asarray(f)*(n*d)

def shift(a, shift, axes=None, nofft=False, deriv=None, engine="auto"):
	"""pixell/fft.py:350-369: shift a by a (fractional) number of samples along the given axes (Fourier shift theorem);
	deriv = i also differentiates along axis i; nofft: a is already the transform and the transform is returned"""
	a = np.asanyarray(a)
	work = a.astype(np.result_type(a.dtype, np.complex64), copy=True)
	shift = np.atleast_1d(shift)
	axes = tuple(range(-len(shift), 0)) if axes is None else _astuple(axes)
	fa = work if nofft else fft(work, axes=list(axes))
	for i, ax in enumerate(axes):
		ax %= work.ndim
		fr = fftfreq(work.shape[ax])
		ph = np.exp(-2j*np.pi*fr*shift[i])
		if deriv == i: ph = ph*(-2j*np.pi*fr)
		fa *= ph.reshape((1,)*ax + (-1,) + (1,)*(work.ndim-ax-1))
	res = fa if nofft else ifft(fa, work, axes=list(axes), normalize=True)
	return res if np.iscomplexobj(a) else res.real

def resample_fft(fa, n, out=None, axes=-1, norm=1, op=lambda a, b: b):
	"""pixell/fft.py:389-433: pad or truncate the transform fa along `axes` to n samples (low frequencies of both signs are
	kept: the first c//2 and the last c - c//2 entries, c = min(old, new) length), times norm; out = op(out, ...)"""
	fa = np.asanyarray(fa)
	axes = _astuple(axes)
	n = [int(v) for v in (np.zeros(len(axes), int) + n)]
	oshape = list(fa.shape)
	for ax, m in zip(axes, n): oshape[ax] = m
	oshape = tuple(oshape)
	if out is None: out = np.zeros(oshape, fa.dtype)
	elif tuple(out.shape) != oshape: raise ValueError("out argument has wrong shape in resample. Expected %s but got %s" % (str(oshape), str(out.shape)))
	for corner in np.ndindex(*([2]*len(axes))):
		sel = [slice(None)]*len(oshape)
		for which, ax in zip(corner, axes):
			c = min(fa.shape[ax], oshape[ax])
			sel[ax] = slice(0, c//2) if which == 0 else slice(-(c-c//2), None)
		sel = tuple(sel)
		src = fa[sel] if norm == 1 else fa[sel]*norm
		out[sel] = op(out[sel], src)
	return out

def resample(a, n, axes=None, nthread=0, engine="auto"):
	"""pixell/fft.py:371-387: Fourier resampling of the given axes (default: the last ones) to length n"""
	a = np.asarray(a)
	n = _astuple(n)
	if axes is None: axes = [-len(n)+i for i in range(len(n))]
	axes = list(_astuple(axes))
	if len(n) != len(axes): raise ValueError("Resize size n = %s does not match axes = %s" % (str(n), str(axes)))
	fa = fft(a.astype(np.result_type(a.dtype, np.complex64)), axes=axes)
	fa = resample_fft(fa, n, axes=axes, norm=1/np.prod([a.shape[ax] for ax in axes]))
	out = ifft(fa, axes=axes, normalize=False)
	return out if np.iscomplexobj(a) else out.real

def fftfreq(n, d=1.0, dtype=np.float64): return np.fft.fftfreq(n, d=d).astype(dtype, copy=False)
def rfftfreq(n, d=1.0, dtype=np.float64): return np.arange(n//2+1, dtype=dtype)/(n*d)
