"""pixell_b200.enmap -- the flat-sky harmonic functions of pixell.enmap that sit on the FFT engine
(reference pixell/enmap.py): fft :1307-1322, ifft :1323-1337, laxes :1273-1294, lmap :1242-1250,
modlmap :1252-1258, extent (cylindrical) :998-1014, area :1032-1036, pixsize :1097-1099,
smooth_gauss :1429-1439.  Maps are geometry.ndmap (numpy + wcs) or torch CUDA tensors with wcs=.
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

def fft(emap, omap=None, nthread=0, normalize=True, adjoint_ifft=False, dct=False, wcs=None):
	"""2-D FFT of the map pixels -> complex map (pixell/enmap.py:1307-1322)."""
	if dct: raise NotImplementedError("dct=True is not provided by pixell_b200")
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
	if dct: raise NotImplementedError("dct=True is not provided by pixell_b200")
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
	f = enfft.rfft(emap, axes=[-2, -1])
	if L.is_torch(f):
		import torch
		f *= torch.as_tensor(filt, device=f.device).to(f.real.dtype)
	else: f *= filt.astype(f.real.dtype)
	out = enfft.irfft(f, n=nx, axes=[-2, -1], normalize=True)
	if not L.is_torch(out): out = geometry.ndmap(out, wcs)
	return out
