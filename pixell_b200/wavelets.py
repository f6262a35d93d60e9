"""pixell_b200.wavelets -- curved-sky wavelet (needlet) transform on the B200 engine (reference pixell/wavelets.py:
bases :48-75 ButterTrim, :131-161 CosineNeedlet; WaveletTransform :206-417; scale geometries :472-495).

map2wave: one exact map2alm, then per scale transfer_alm (to the scale's lmax) -> lmul (filter / norm) -> alm2map on the
scale's own full-sky grid; wave2map is the reverse with the alm summed over scales.  The alm stays on the device
between the steps when the map is a torch CUDA tensor.  Provided: curved mode on maps that cover all of RA (full sky or
declination bands).  The flat-sky mode (enmap.resample_fft), the variance bases and HaarTransform are not provided."""
import numpy as np
from . import curvedsky, geometry, _lib as L
from .geometry import DEG

def trim_kernel(a, tol): return np.clip(a*(1+2*tol)-tol, 0, 1)

class ButterTrim:
	"""Butterworth wavelet basis made harmonically compact by clipping its tails (reference wavelets.py:48-75)"""
	def __init__(self, step=2, shape=7, trim=1e-2, lmin=None, lmax=None):
		self.step, self.shape, self.trim, self.lmin, self.lmax = step, shape, trim, lmin, lmax
		if lmin is not None and lmax is not None: self._finalize()
	def with_bounds(self, lmin, lmax): return ButterTrim(step=self.step, shape=self.shape, trim=self.trim, lmin=lmin, lmax=lmax)
	def __call__(self, i, l):
		profile = np.full(np.shape(l), 1.0) if i == self.n-1 else self.kernel(i, l)
		if i > 0: profile = profile - self.kernel(i-1, l)
		return profile**0.5
	def kernel(self, i, l):
		return trim_kernel(1/(1 + (l/(self.lmin*self.step**(i+0.5)))**(self.shape/np.log(self.step))), self.trim)
	def _finalize(self):
		self.n = int((np.log(self.lmax)-np.log(self.lmin))/np.log(self.step))
		self.lmaxs = np.ceil(self.lmin*((1+2*self.trim)/self.trim-1)**(np.log(self.step)/self.shape)*self.step**(np.arange(self.n)+0.5)).astype(int)
		self.lmaxs[-1] = self.lmax

class CosineNeedlet:
	"""Cosine-shaped needlets peaking at the multipoles lpeaks (reference wavelets.py:131-161)"""
	def __init__(self, lpeaks):
		self.lpeaks = np.asarray(lpeaks)
		self.lmaxs = np.append(self.lpeaks[1:], self.lpeaks[-1])
		self.lmins = np.append(self.lpeaks[0], self.lpeaks[:-1])
		self.lmin, self.lmax = self.lpeaks[0], self.lpeaks[-1]
	@property
	def n(self): return len(self.lpeaks)
	def with_bounds(self, lmin, lmax): return self
	def __call__(self, i, l):
		l = np.asarray(l, dtype=np.float64)
		out = l*0.
		lp = self.lpeaks[i]
		if i > 0:
			lm = self.lpeaks[i-1]; sel = (l >= lm) & (l < lp)
			out[sel] = np.cos(np.pi*(lp-l[sel])/(lp-lm)/2.)
		if i < self.n-1:
			ln = self.lpeaks[i+1]; sel = (l >= lp) & (l < ln)
			out[sel] = np.cos(np.pi*(l[sel]-lp)/(ln-lp)/2.)
		return out

class multimap:
	"""a group of maps with common leading dimensions and per-scale geometries (the part of pixell.multimap used here)"""
	def __init__(self, maps, geometries): self.maps, self.geometries = list(maps), list(geometries)
	@property
	def pre(self): return tuple(self.maps[0].shape[:-2])
	@property
	def dtype(self): return L.buffer_info(self.maps[0])[2]
	@property
	def nmap(self): return len(self.maps)

def make_wavelet_geometry_curved(ishape, iwcs, ores, minres=2*DEG):
	"""full-sky quadrature grid of resolution <= ores, cropped in declination to the rows the input map covers
	(reference wavelets.py:472-495; maps must span all of RA)"""
	res = min(np.pi/np.ceil(np.pi/ores), minres)
	shape, wcs = geometry.fullsky_geometry(res=res)
	if abs(abs(iwcs.wcs.cdelt[0])*ishape[-1]-360) > 1e-6: raise NotImplementedError("pixell_b200.wavelets: maps must cover all of RA")
	d1, d2 = np.sort(geometry.dec_of(iwcs, np.array([-0.5, ishape[-2]-0.5])))
	y = np.sort(geometry.ypix_of(wcs, np.rad2deg(np.clip([d1, d2], -np.pi/2, np.pi/2))))
	y1, y2 = max(0, int(np.floor(y[0]+0.5))), min(shape[0], int(np.ceil(y[1]+0.5)))
	if y1 == 0 and y2 == shape[0]: return shape, wcs
	return geometry.slice_geometry(shape, wcs, y1, y2)

class WaveletTransform:
	"""Curved-sky wavelet transform (reference wavelets.py:206-417).  uht: a pixell_b200.uharm.UHT in "curved" mode."""
	def __init__(self, uht, basis=ButterTrim(), ores=None, norms=None, geometries=None):
		if uht.mode != "curved": raise NotImplementedError("pixell_b200.wavelets: only the curved-sky mode is provided")
		self.uht = uht
		ires = np.min(np.abs(uht.wcs.wcs.cdelt))*DEG
		if basis.lmin is None or basis.lmax is None: basis = basis.with_bounds(int(np.ceil(np.pi/np.max(np.array(uht.shape)*ires))), uht.lmax)
		self.basis = basis
		self.geometries = geometries
		if self.geometries is None:
			oress = np.maximum(np.pi/np.asarray(self.basis.lmaxs), ires) if ores is None else np.zeros(self.basis.n)+ores
			self.geometries = [make_wavelet_geometry_curved(uht.shape, uht.wcs, o) for o in oress]
		self.filters, self.norms, self.lmids = self._prepare_filters()
		if norms is not None: self.norms[:] = norms
	@property
	def shape(self): return self.uht.shape
	@property
	def wcs(self): return self.uht.wcs
	@property
	def nlevel(self): return len(self.geometries)
	def _prepare_filters(self):
		filters, norms, lmids = [], [], []
		ls = np.arange(self.basis.lmax+1, dtype=np.float64)
		for i in range(self.nlevel):
			F = self.basis(i, ls)
			W = F**2*(2*ls+1)/(4*np.pi)
			Wtot = np.sum(W)
			filters.append(F); norms.append(Wtot**0.5); lmids.append(np.sum(W*ls)/Wtot)
		return filters, np.asarray(norms), np.asarray(lmids)
	def _zeros(self, ref, shape, wcs, dtype):
		if L.is_torch(ref):
			import torch
			return torch.zeros(tuple(shape), dtype={np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32}[np.dtype(dtype)], device=ref.device)
		return geometry.zeros(tuple(shape), wcs, dtype)
	def map2wave(self, map, owave=None, fl=None, scales=None, fill_value=None):
		"""map[..., ny, nx] -> multimap of wavelet coefficient maps (spin-0 transforms of every leading component)"""
		scales = range(self.nlevel) if scales is None else scales
		rdt = L.buffer_info(map)[2]
		pre = tuple(map.shape[:-2])
		if owave is None: owave = multimap([self._zeros(map, pre+tuple(s[-2:]), w, rdt) for s, w in self.geometries], self.geometries)
		ainfo = curvedsky.alm_info(lmax=int(self.basis.lmax))
		alm = curvedsky.map2alm(map, ainfo=ainfo, spin=[0], wcs=self.uht.wcs)
		if fl is not None: alm = curvedsky.almxfl(alm, fl, ainfo=ainfo)
		for i, (shape, wcs) in enumerate(self.geometries):
			if i in scales:
				small = curvedsky.alm_info(lmax=int(self.basis.lmaxs[i]))
				asmall = curvedsky.transfer_alm(ainfo, alm, small)
				small.lmul(asmall, self.filters[i][:small.lmax+1]/self.norms[i], asmall)
				curvedsky.alm2map(asmall, owave.maps[i], spin=[0], ainfo=small, wcs=wcs)
			elif fill_value is not None: owave.maps[i][...] = fill_value
		return owave
	def wave2map(self, wave, omap=None):
		"""multimap of wavelet coefficients -> map on the transform's own geometry"""
		ainfo = curvedsky.alm_info(lmax=int(self.basis.lmax))
		oalm = None
		for i, (shape, wcs) in enumerate(self.geometries):
			small = curvedsky.alm_info(lmax=int(self.basis.lmaxs[i]))
			asmall = curvedsky.map2alm(wave.maps[i], ainfo=small, spin=[0], wcs=wcs)
			small.lmul(asmall, self.filters[i][:small.lmax+1]*self.norms[i], asmall)
			part = curvedsky.transfer_alm(small, asmall, ainfo)
			oalm = part if oalm is None else oalm + part
		if omap is None: omap = self._zeros(wave.maps[0], wave.pre+tuple(self.uht.shape), self.uht.wcs, wave.dtype)
		return curvedsky.alm2map(oalm, omap, spin=[0], ainfo=ainfo, wcs=self.uht.wcs)
	def get_ls(self, i): return self.uht.l
