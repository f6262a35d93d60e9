"""pixell_b200.enmap -- the flat-sky harmonic functions of pixell.enmap that sit on the FFT engine
(reference pixell/enmap.py): fft :1307-1322, ifft :1323-1337, laxes :1273-1294, lmap :1242-1250,
modlmap :1252-1258, extent (cylindrical) :998-1014, area :1032-1036, pixsize :1097-1099,
smooth_gauss :1429-1439, map2harm / harm2map and their adjoints :1358-1389, queb_rotmat :1391-1400,
rotate_pol :1402-1416, map_mul :1418-1427, calc_window / apply_window :1470-1500, spin_helper :3378-3388.  Maps are geometry.ndmap (numpy + wcs) or torch CUDA tensors with wcs=.
Only separable cylindrical (CAR) geometries are handled, as everywhere in this package.

The normalisation factor of fft/ifft is folded into the last FFT pass (no extra sweep over the map).
"""
import numpy as np
from . import _lib as L, fft as enfft, geometry
from .geometry import DEG

def _wcs(emap, wcs): return geometry.wcs_of(emap, wcs)

def area(shape, wcs):
	"""enmap.area for cylindrical projections (pixell/enmap.py:1032-1036)"""
	d = np.sort(geometry.dec_of(wcs, np.array([-0.5, shape[-2]-0.5])))
	d1, d2 = max(-np.pi/2, d[0]), min(np.pi/2, d[1])
	return (np.sin(d2)-np.sin(d1))*abs(wcs.wcs.cdelt[0])*shape[-1]*DEG

def pixsize(shape, wcs): return area(shape, wcs)/np.prod(shape[-2:])

def extent(shape, wcs, signed=False):
	"""enmap.extent_cyl (pixell/enmap.py:998-1014): [height, width] in radians with height*width = area"""
	dec1, dec2 = geometry.dec_of(wcs, np.array([-0.5, shape[-2]-0.5]))
	ysign = 1
	if dec1 > dec2: dec1, dec2, ysign = dec2, dec1, -1
	dec1, dec2 = max(-np.pi/2, dec1), min(np.pi/2, dec2)
	mean_cos = (np.sin(dec2)-np.sin(dec1))/(dec2-dec1)
	ext = np.array([(dec2-dec1)*ysign, shape[-1]*wcs.wcs.cdelt[0]*mean_cos*DEG])
	return ext if signed else np.abs(ext)

def laxes(shape, wcs, oversample=1, broadcastable=False):
	"""pixell/enmap.py:1273-1294 (oversample=1 only)"""
	if oversample != 1: raise NotImplementedError("oversample != 1")
	step = extent(shape, wcs, signed=True)/shape[-2:]
	ly = np.fft.fftfreq(shape[-2], step[0])*2*np.pi
	lx = np.fft.fftfreq(shape[-1], step[1])*2*np.pi
	if broadcastable: ly, lx = ly[:, None], lx[None, :]
	return ly, lx

def lmap(shape, wcs):
	ly, lx = laxes(shape, wcs)
	data = np.empty((2, ly.size, lx.size))
	data[0] = ly[:, None]; data[1] = lx[None, :]
	return geometry.ndmap(data, wcs)

def modlmap(shape, wcs, min=0):
	ly, lx = laxes(shape, wcs)
	l = (ly[:, None]**2 + lx[None, :]**2)**0.5
	if min > 0: l = np.maximum(l, min)
	return geometry.ndmap(l, wcs)

def _complex_like(emap):
	dt = L.buffer_info(emap)[2]
	return np.result_type(dt, 0j)

def _norm(shape, wcs, normalize, flip_phys):
	norm = 1.0
	if normalize: norm /= np.prod(shape[-2:])**0.5
	if normalize in ["phy", "phys", "physical"]:
		norm = norm/pixsize(shape, wcs)**0.5 if flip_phys else norm*pixsize(shape, wcs)**0.5
	return norm

def _dct2(emap, omap, normalize, flip_phys, wcs, inverse):
	"""dct=True branch of fft / ifft (pixell/enmap.py:1314-1320, 1327-1333): DCT-I over the last two axes; the reference
	normalises by prod(2 n - 1)^(1/2) per call, mirrored as written"""
	res = enfft.idct(emap, omap, axes=[-2, -1], normalize=False) if inverse else enfft.dct(emap, omap, axes=[-2, -1])
	norm = 1.0
	if normalize: norm /= float(np.prod(2*np.array(emap.shape[-2:])-1))**0.5
	if normalize in ["phy", "phys", "physical"]:
		w = _wcs(emap, wcs)
		norm = norm/pixsize(emap.shape, w)**0.5 if flip_phys else norm*pixsize(emap.shape, w)**0.5
	if norm != 1: res *= norm
	if not L.is_torch(res) and getattr(emap, "wcs", wcs) is not None: res = geometry.ndmap(res, getattr(emap, "wcs", wcs))
	return res

def dct(emap, omap=None, nthread=0, normalize=True, wcs=None):
	"""pixell/enmap.py:1339-1340"""
	return fft(emap, omap=omap, nthread=nthread, normalize=normalize, dct=True, wcs=wcs)
def idct(emap, omap=None, nthread=0, normalize=True, wcs=None):
	"""pixell/enmap.py:1341-1342"""
	return ifft(emap, omap=omap, nthread=nthread, normalize=normalize, dct=True, wcs=wcs)

def fft(emap, omap=None, nthread=0, normalize=True, adjoint_ifft=False, dct=False, wcs=None):
	"""2-D FFT of the map pixels -> complex map (pixell/enmap.py:1307-1322)."""
	if dct: return _dct2(emap, omap, normalize, adjoint_ifft, wcs, False)
	ctype = _complex_like(emap)
	if omap is None: omap = enfft._empty_like(emap, emap.shape, ctype)
	src = emap
	if L.buffer_info(emap)[2].kind != "c":      # the reference promotes real input to complex (fft.py:148-151)
		src = emap.to(omap.dtype) if L.is_torch(emap) else np.asarray(emap).astype(ctype)
	w = None
	if normalize in ["phy", "phys", "physical"]: w = _wcs(emap, wcs)
	norm = _norm(emap.shape, w, normalize, adjoint_ifft)
	enfft.transform(src, omap, (-2, -1), True, norm)
	if not L.is_torch(omap) and getattr(emap, "wcs", wcs) is not None: omap = geometry.ndmap(omap, getattr(emap, "wcs", wcs))
	return omap

def ifft(emap, omap=None, nthread=0, normalize=True, adjoint_fft=False, dct=False, wcs=None):
	"""2-D inverse FFT (pixell/enmap.py:1323-1337)."""
	if dct: return _dct2(emap, omap, normalize, not adjoint_fft, wcs, True)
	if omap is None: omap = enfft._empty_like(emap, emap.shape, L.buffer_info(emap)[2])
	w = None
	if normalize in ["phy", "phys", "physical"]: w = _wcs(emap, wcs)
	norm = _norm(emap.shape, w, normalize, not adjoint_fft)
	enfft.transform(emap, omap, (-2, -1), False, norm)
	if not L.is_torch(omap) and getattr(emap, "wcs", wcs) is not None: omap = geometry.ndmap(omap, getattr(emap, "wcs", wcs))
	return omap

def smooth_gauss(emap, sigma, wcs=None):
	"""Gaussian smoothing in harmonic space (pixell/enmap.py:1429-1439) through the real transforms:
	irfft2(rfft2(m) exp(-l^2 sigma^2/2)) / npix."""
	wcs = _wcs(emap, wcs)
	if sigma == 0: return emap.clone() if L.is_torch(emap) else emap.copy()
	ny, nx = emap.shape[-2:]
	ly, lx = laxes(emap.shape, wcs)
	filt = np.exp(-0.5*sigma**2*(ly[:, None]**2 + lx[None, :nx//2+1]**2))
	if sigma < 0: filt = 1-filt          # negative sigma: the complementary high-pass filter (pixell/enmap.py:1437-1438)
	f = enfft.rfft(emap, axes=[-2, -1])
	if L.is_torch(f): enfft.fourier_filter(f, f2=filt)
	else: f *= filt.astype(f.real.dtype)
	out = enfft.irfft(f, n=nx, axes=[-2, -1], normalize=True)
	if not L.is_torch(out): out = geometry.ndmap(out, wcs)
	return out

def pixwin_1d(f, order=0):
	"""1-D pixel window at dimensionless frequency f (pixell/utils.py:856-868)"""
	if order is None or order == "none": return f*0+1
	if order == 0 or order == "nn": return np.sinc(f)
	if order == 1 or order == "lin": return np.sinc(f)**2/(1/3*(2+np.cos(2*np.pi*f)))
	raise ValueError("Unsupported pixwin order %s" % str(order))

def calc_window(shape, order=0, scale=1):
	"""separable Fourier-space pixel window wy, wx (pixell/enmap.py:1470-1482)"""
	return pixwin_1d(np.fft.fftfreq(shape[-2], scale), order=order), pixwin_1d(np.fft.fftfreq(shape[-1], scale), order=order)

def apply_window(emap, pow=1.0, order=0, scale=1, nofft=False, wcs=None):
	"""Multiply by the pixel window to the given power in Fourier space (pixell/enmap.py:1484-1496); real maps go
	through the real transforms (half the spectrum), nofft=True takes and returns a Fourier map."""
	wy, wx = calc_window(emap.shape, order=order, scale=scale)
	wy, wx = wy**pow, wx**pow
	def mul(f, wy, wx):
		if L.is_torch(f) and f.is_contiguous(): enfft.fourier_filter(f, fy=wy, fx=wx)      # one pass (b2_fourier_filter)
		elif L.is_torch(f):
			import torch
			f *= torch.as_tensor(wy, device=f.device).to(f.real.dtype)[:, None]
			f *= torch.as_tensor(wx, device=f.device).to(f.real.dtype)[None, :]
		else:
			f *= wy.astype(f.real.dtype)[:, None]; f *= wx.astype(f.real.dtype)[None, :]
		return f
	if nofft: return mul(emap.clone() if L.is_torch(emap) else emap.copy(), wy, wx)
	nx = emap.shape[-1]
	if L.buffer_info(emap)[2].kind == "c":
		out = ifft(mul(fft(emap, wcs=wcs), wy, wx), wcs=wcs)
		return out.real                  # the reference returns ifft(...).real (pixell/enmap.py:1495)
	f = mul(enfft.rfft(emap, axes=[-2, -1]), wy, wx[:nx//2+1])
	out = enfft.irfft(f, n=nx, axes=[-2, -1], normalize=True)
	w = getattr(emap, "wcs", wcs)
	if not L.is_torch(out) and w is not None: out = geometry.ndmap(out, w)
	return out

def unapply_window(emap, pow=1.0, order=0, scale=1, nofft=False, wcs=None):
	return apply_window(emap, pow=-pow, order=order, scale=scale, nofft=nofft, wcs=wcs)

# ------------------------------------------------------------------ T,Q,U <-> T,E,B (pixell/enmap.py:1358-1427)

def spin_helper(spin, n):
	"""(spin, first, last+1) component groups; the spin list is cycled (pixell/enmap.py:3378-3388)"""
	spin = np.array(spin).reshape(-1)
	i1 = 0; ci = 0
	while i1 < n:
		s = int(spin[ci % len(spin)])
		i2 = i1 + (2 if s != 0 else 1)
		if i2 > n: raise IndexError("Unpaired component in spin transform")
		yield s, i1, i2
		i1 = i2; ci += 1

def queb_rotmat(lmap, inverse=False, iau=False, spin=2, wcs=None):
	"""2x2 rotation [[c,-s],[s,c]], angle spin*atan2(+-lx, ly) (pixell/enmap.py:1391-1400).  lmap: [2,ny,nx] or the
	broadcastable (ly[:,None], lx[None,:]) pair."""
	sign = 1
	if iau: sign = -sign
	if inverse: sign = -sign
	a = spin*np.arctan2(sign*lmap[1], lmap[0])
	c, s = np.cos(a), np.sin(a)
	return np.array([[c, -s], [s, c]])

def map_mul(mat, vec):
	"""element-wise matrix product along the last non-pixel axes (pixell/enmap.py:1418-1427)"""
	mat = np.asanyarray(mat)
	if mat.ndim <= 3: return mat*vec
	return np.einsum("...abyx,...byx->...ayx", mat, vec)

def rotate_pol(emap, angle, comps=[-2, -1], spin=2, axis=-3):
	"""pixell/enmap.py:1402-1416"""
	if spin == 0: return emap
	axis %= emap.ndim
	c, s = np.cos(spin*angle), np.sin(spin*angle)
	res = emap.clone() if L.is_torch(emap) else emap.copy()
	pre = (slice(None),)*axis
	res[pre+(comps[0],)] = c*emap[pre+(comps[0],)] - s*emap[pre+(comps[1],)]
	res[pre+(comps[1],)] = s*emap[pre+(comps[0],)] + c*emap[pre+(comps[1],)]
	return res

def _rotate_pairs(hmap, wcs, spin, sign):
	"""in place QU<->EB rotation of a complex Fourier map [..., ncomp, ny, nx] on the device (b2_queb_rotate)"""
	ny, nx = hmap.shape[-2:]
	ncomp = hmap.shape[-3]
	ly, lx = laxes(hmap.shape, wcs)
	ly = np.ascontiguousarray(ly, dtype=np.float64); lx = np.ascontiguousarray(lx, dtype=np.float64)
	ptr, mem, dt = L.buffer_info(hmap)
	st = L.strides_elems(hmap)
	if st[-1] != 1 or st[-2] != nx: raise ValueError("map2harm/harm2map: the pixel axes must be contiguous")
	pre = hmap.shape[:-3]
	prec = L.F64 if dt.itemsize == 16 else L.F32
	stream = L.current_stream(hmap)
	L.init()
	for s, i1, i2 in spin_helper(spin, ncomp):
		if s == 0: continue
		for idx in (np.ndindex(*pre) if pre else [()]):
			off = sum(i*k for i, k in zip(idx, st[:-3])) + i1*st[-3]
			L.check(L.lib().b2_queb_rotate(ptr + off*dt.itemsize, int(st[-3]), 1, 0, int(ny), int(nx), L.p_dbl(ly), L.p_dbl(lx),
				int(s), int(sign), prec, mem, stream))
	return hmap

def _on_device(emap):
	"""(torch CUDA tensor, was_numpy): numpy maps visit the device once for the whole map2harm / harm2map"""
	if L.is_torch(emap): return emap, False
	import torch
	L.init()
	return torch.from_numpy(np.ascontiguousarray(emap)).cuda(), True

def map2harm(emap, nthread=0, normalize=True, iau=False, spin=[0, 2], adjoint_harm2map=False, wcs=None):
	"""2-D FFT of the pixels followed by the Q,U -> E,B rotation of every spin-s pair (pixell/enmap.py:1358-1372)."""
	wcs = _wcs(emap, wcs)
	dev, was_np = _on_device(emap)
	hmap = fft(dev, normalize=normalize, adjoint_ifft=adjoint_harm2map, wcs=wcs)
	if hmap.ndim > 2: _rotate_pairs(hmap, wcs, spin, -1 if iau else 1)
	if was_np: hmap = geometry.ndmap(hmap.cpu().numpy(), wcs)
	return hmap

def harm2map(emap, nthread=0, normalize=True, iau=False, spin=[0, 2], keep_imag=False, adjoint_map2harm=False, wcs=None):
	"""E,B -> Q,U rotation followed by the inverse 2-D FFT (pixell/enmap.py:1373-1383)."""
	wcs = _wcs(emap, wcs)
	dev, was_np = _on_device(emap)
	if dev.ndim > 2:
		if not was_np: dev = dev.clone()
		_rotate_pairs(dev, wcs, spin, 1 if iau else -1)
	res = ifft(dev, normalize=normalize, adjoint_fft=adjoint_map2harm, wcs=wcs)
	if not keep_imag: res = res.real
	if was_np: res = geometry.ndmap(res.cpu().numpy(), wcs)
	return res

def map2harm_adjoint(emap, nthread=0, normalize=True, iau=False, spin=[0, 2], keep_imag=False, wcs=None):
	return harm2map(emap, nthread=nthread, normalize=normalize, iau=iau, spin=spin, keep_imag=keep_imag, adjoint_map2harm=True, wcs=wcs)

def harm2map_adjoint(emap, nthread=0, normalize=True, iau=False, spin=[0, 2], wcs=None):
	return map2harm(emap, nthread=nthread, normalize=normalize, iau=iau, spin=spin, adjoint_harm2map=True, wcs=wcs)
