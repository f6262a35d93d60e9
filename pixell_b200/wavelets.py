"""pixell_b200.wavelets -- wavelet (needlet) transforms on the B200 engine (reference pixell/wavelets.py: bases :15-161
Butterworth, ButterTrim, DigitalButterTrim, AdriSD, CosineNeedlet; variance basis :168-206 VarButter; WaveletTransform
:206-417; HaarTransform :419-456; scale geometries :463-495).

Curved mode: map2wave is one exact map2alm, then per scale transfer_alm (to the scale's lmax) -> lmul (filter / norm) ->
alm2map on the scale's own full-sky grid; wave2map is the reverse with the alm summed over scales.  The alm stays on the
device between the steps when the map is a torch CUDA tensor (maps must cover all of RA: full sky or declination bands).
Flat mode: one 2-D FFT of the map (the engine's TMA path), per scale the low-frequency corners of the Fourier array are
cut out (enmap.resample_fft, corner=True), filtered and transformed back on the scale's own smaller grid.
HaarTransform is the reference's pixel-space transform (block means; host arrays)."""
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
	def get_variance_basis(self): return VarButter(step=self.step, shape=self.shape, lmin=self.lmin, lmax=self.lmax)
	def kernel(self, i, l):
		return trim_kernel(1/(1 + (l/(self.lmin*self.step**(i+0.5)))**(self.shape/np.log(self.step))), self.trim)
	def _finalize(self):
		self.n = int((np.log(self.lmax)-np.log(self.lmin))/np.log(self.step))
		self.lmaxs = np.ceil(self.lmin*((1+2*self.trim)/self.trim-1)**(np.log(self.step)/self.shape)*self.step**(np.arange(self.n)+0.5)).astype(int)
		self.lmaxs[-1] = self.lmax

class Butterworth:
	"""Differences of Butterworth low-pass filters (reference wavelets.py:15-46); tol sets where a scale's map may stop"""
	def __init__(self, step=2, shape=7, tol=1e-3, lmin=None, lmax=None):
		self.step, self.shape, self.tol, self.lmin, self.lmax = step, shape, tol, lmin, lmax
		if lmin is not None and lmax is not None: self._finalize()
	def with_bounds(self, lmin, lmax): return Butterworth(step=self.step, shape=self.shape, tol=self.tol, lmin=lmin, lmax=lmax)
	def __call__(self, i, l):
		profile = np.full(np.shape(l), 1.0) if i == self.n-1 else self.kernel(i, l)
		if i > 0: profile = profile - self.kernel(i-1, l)
		return profile**0.5
	def get_variance_basis(self): return VarButter(step=self.step, shape=self.shape, tol=self.tol, lmin=self.lmin, lmax=self.lmax)
	def kernel(self, i, l): return 1/(1 + (l/(self.lmin*self.step**(i+0.5)))**(self.shape/np.log(self.step)))
	def _finalize(self):
		self.n = int((np.log(self.lmax)-np.log(self.lmin))/np.log(self.step))
		self.lmaxs = np.round(self.lmin*(1/self.tol-1)**(np.log(self.step)/self.shape)*self.step**(np.arange(self.n)+0.5)).astype(int)
		self.lmaxs[-1] = self.lmax

def digitize(a):
	"""on/off array approximating a smooth array with values in [0, 1] (reference wavelets.py:458-462)"""
	f = np.round(np.cumsum(a))
	return np.concatenate([[1], f[1:] != f[:-1]])

class DigitalButterTrim(ButterTrim):
	"""ButterTrim with every filter replaced by a comb of top hats: orthogonal scales (reference wavelets.py:77-107)"""
	def with_bounds(self, lmin, lmax): return DigitalButterTrim(step=self.step, shape=self.shape, trim=self.trim, lmin=lmin, lmax=lmax)
	def __call__(self, i, l):
		idx = np.rint(np.asarray(l, dtype=np.float64)).astype(int)            # nearest sample (utils.interpol order 0); zero outside
		ok = (idx >= 0) & (idx < self.profiles.shape[1])
		return np.where(ok, self.profiles[i][np.clip(idx, 0, self.profiles.shape[1]-1)], 0.0)
	def get_variance_basis(self): raise NotImplementedError
	def _finalize(self):
		ButterTrim._finalize(self)
		l = np.arange(self.lmax)
		kernels = np.array([np.zeros(l.size)] + [digitize(self.kernel(i, l)) for i in range(self.n-1)] + [np.full(l.size, 1.0)])
		kernels = np.sort(kernels, 0)
		self.profiles = kernels[1:]-kernels[:-1]

class AdriSD:
	"""Scale-discrete basis from the optweight library (reference wavelets.py:109-129); needs `optweight`, as the reference does"""
	def __init__(self, lamb=2, lmin=None, lmax=None):
		self.lamb, self.lmin, self.lmax = lamb, lmin, lmax
		if lmin is not None: self._finalize()
	def with_bounds(self, lmin, lmax): return AdriSD(lamb=self.lamb, lmin=lmin, lmax=lmax)
	@property
	def n(self): return len(self.profiles)
	def __call__(self, i, l): return np.interp(l, np.arange(self.profiles[i].size), self.profiles[i])
	def get_variance_basis(self): raise NotImplementedError
	def _finalize(self):
		from optweight import wlm_utils
		self.profiles, self.lmaxs = wlm_utils.get_sd_kernels(self.lamb, self.lmax, lmin=self.lmin)

class RadialFourierTransform:
	"""log-spaced radial (Hankel) transform pair on the flat sky (reference pixell/utils.py:3206-3290, scipy.fft.fht)"""
	def __init__(self, lrange=None, rrange=None, n=512, pad=256):
		if lrange is None and rrange is None: lrange = [0.1, 1e7]
		if lrange is None: lrange = [1/rrange[1], 1/rrange[0]]
		logl1, logl2 = np.log(lrange)
		self.dlog = (logl2-logl1)/n
		self.l = np.exp((logl2+logl1)/2 + (np.arange(1, n+2*pad+1)-((n+1)/2+pad))*self.dlog)
		self.r = 1/self.l[::-1]
		self.pad = pad
	def real2harm(self, rprof):
		import scipy.fft
		if callable(rprof): rprof = rprof(self.r)
		return 2*np.pi*scipy.fft.fht(rprof*self.r, self.dlog, 0)/self.l
	def harm2real(self, lprof):
		import scipy.fft
		if callable(lprof): lprof = lprof(self.l)
		return scipy.fft.ifht(lprof/(2*np.pi)*self.l, self.dlog, 0)/self.r
	def unpad(self, *arrs):
		res = arrs if self.pad == 0 else tuple(a[..., self.pad:-self.pad] for a in arrs)
		return res[0] if len(arrs) == 1 else res

class VarButter:
	"""Variance basis of the Butterworth wavelets: how white noise transforms (reference wavelets.py:168-206): the harmonic
	profile of the squared real-space kernel of every scale"""
	def __init__(self, step=2, shape=7, tol=1e-3, lmin=None, lmax=None):
		self.step, self.shape, self.tol, self.lmin, self.lmax, self.basis = step, shape, tol, lmin, lmax, None
		if lmin is not None: self._finalize()
	@property
	def n(self): return self.basis.n
	@property
	def lmaxs(self): return self.basis.lmaxs
	def with_bounds(self, lmin, lmax): return VarButter(step=self.step, shape=self.shape, tol=self.tol, lmin=lmin, lmax=lmax)
	def __call__(self, i, l): return np.interp(l, self.l, self.kernels[i])
	def _kernel_helper(self, i, rft):
		if i < self.basis.n-1: F = self.basis(i, rft.l)
		else:
			# the last, unbounded scale gets a cutoff at lmax: the map holds no power beyond it
			kernel = 1/(1 + (rft.l/self.basis.lmax)**(self.basis.shape/np.log(self.basis.step)))
			F = (kernel - self.basis.kernel(i-1, rft.l))**0.5
		return rft.unpad(rft.real2harm(rft.harm2real(F)**2))
	def _finalize(self):
		self.basis = Butterworth(step=self.step, shape=self.shape, tol=self.tol, lmin=self.lmin, lmax=self.lmax)
		rft = RadialFourierTransform()
		self.kernels = [self._kernel_helper(i, rft) for i in range(self.n)]
		self.l = rft.unpad(rft.l)

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

def scale_wcs(wcs, scale, corner=True):
	"""wcsutils.scale with rowmajor=True (reference wcsutils.py:188-204): pixel density times scale = (sy, sx)"""
	sy, sx = (np.zeros(2)+scale)
	w = wcs.wcs
	crpix = np.array(w.crpix, float) - (0.5 if corner else 0)
	crpix = crpix*np.array([sx, sy]) + (0.5 if corner else 0)
	return geometry.CarWCS(w.crval, np.array(w.cdelt, float)/np.array([sx, sy]), crpix, getattr(w, "ctype", ("RA---CAR", "DEC--CAR")))

def make_wavelet_geometry_flat(ishape, iwcs, ires, ores, margin=4):
	"""smaller pixelisation of the same patch for one scale (reference wavelets.py:463-470)"""
	oshape = (np.ceil(np.array(ishape[-2:])*ires/ores)).astype(int)+margin
	oshape = np.minimum(oshape, ishape[-2:])
	return tuple(int(v) for v in oshape), scale_wcs(iwcs, oshape[-2:]/np.array(ishape[-2:], float), corner=True)

def resample_fft(fimap, oshape, fomap=None, add=False, norm=1.0):
	"""enmap.resample_fft(..., corner=True, norm=None) (reference enmap.py:3328-3376): the four low-frequency corners of a
	Fourier array copied (or added) into a Fourier array of another size, then the half-pixel realignment phase.
	numpy arrays or torch tensors."""
	tor = L.is_torch(fimap)
	iy, ix = fimap.shape[-2:]; oy, ox = int(oshape[-2]), int(oshape[-1])
	if fomap is None:
		if tor:
			import torch
			fomap = torch.zeros(tuple(fimap.shape[:-2])+(oy, ox), dtype=fimap.dtype, device=fimap.device)
		else: fomap = np.zeros(tuple(fimap.shape[:-2])+(oy, ox), fimap.dtype)
	cny, cnx = min(iy, oy), min(ix, ox)
	hny, hnx = cny//2, cnx//2
	ty, tx = cny-hny, cnx-hnx
	src = fimap if norm == 1 else fimap*norm
	# realignment with the pixel centres: a shift by off (output pixels) = phase exp(-2 pi i off k/n) per axis
	off = -(0.5 - 0.5*np.array([oy, ox], float)/np.array([iy, ix], float))
	blocks = [(slice(0, hny), slice(0, hnx)), (slice(0, hny), slice(-tx, None)), (slice(-ty, None), slice(0, hnx)), (slice(-ty, None), slice(-tx, None))]
	tmp = fomap if not add else (fomap*0)
	for sy, sx in blocks:
		if (sy.stop == 0 and sy.start == 0) or (sx.stop == 0 and sx.start == 0): continue
		tmp[..., sy, sx] = src[..., sy, sx]
	ky, kx = np.fft.fftfreq(oy), np.fft.fftfreq(ox)
	py, px = np.exp(-2j*np.pi*ky*off[0]), np.exp(-2j*np.pi*kx*off[1])
	if tor:
		import torch
		tmp *= torch.as_tensor(py, device=tmp.device).to(tmp.dtype)[:, None]; tmp *= torch.as_tensor(px, device=tmp.device).to(tmp.dtype)[None, :]
	else:
		tmp *= py.astype(tmp.dtype)[:, None]; tmp *= px.astype(tmp.dtype)[None, :]
	if add: fomap += tmp
	return fomap

class WaveletTransform:
	"""Wavelet transform (reference wavelets.py:206-417).  uht: a pixell_b200.uharm.UHT ("curved" or "flat" mode)."""
	def __init__(self, uht, basis=ButterTrim(), ores=None, norms=None, geometries=None):
		from . import enmap
		self.uht = uht
		cd = np.abs(uht.wcs.wcs.cdelt)*DEG
		if uht.mode == "flat": ires = float(np.max(cd))                        # largest pixel side (enmap.pixshapebounds; separable CAR)
		else: ires = float(np.min(cd))
		if basis.lmin is None or basis.lmax is None:
			lmin, lmax = basis.lmin, basis.lmax
			if uht.mode == "flat":
				if lmax is None: lmax = min(int(np.ceil(np.pi/ires)), uht.lmax)
				if lmin is None: lmin = min(int(np.ceil(np.pi/np.max(enmap.extent(uht.shape, uht.wcs)))), lmax)
			else:
				if lmax is None: lmax = uht.lmax
				if lmin is None: lmin = int(np.ceil(np.pi/np.max(np.array(uht.shape)*ires)))
			basis = basis.with_bounds(lmin, lmax)
		self.basis = basis
		self.geometries = geometries
		if self.geometries is None:
			oress = np.maximum(np.pi/np.asarray(self.basis.lmaxs), ires) if ores is None else np.zeros(self.basis.n)+ores
			if uht.mode == "flat":
				self.geometries = [make_wavelet_geometry_flat(uht.shape, uht.wcs, ires, o) for o in oress[:-1]] + [(tuple(uht.shape), uht.wcs)]
			else:
				self.geometries = [make_wavelet_geometry_curved(uht.shape, uht.wcs, o) for o in oress]
		self.filters, self.norms, self.lmids = self._prepare_filters()
		if norms is not None: self.norms[:] = norms
	@property
	def shape(self): return self.uht.shape
	@property
	def wcs(self): return self.uht.wcs
	@property
	def nlevel(self): return len(self.geometries)
	def get_variance_transform(self):
		"""the transform white-noise variance maps follow (reference wavelets.py:383-384)"""
		return WaveletTransform(self.uht, basis=self.basis.get_variance_basis(), norms=self.norms**2, geometries=self.geometries)
	def _prepare_filters(self):
		filters, norms, lmids = [], [], []
		if self.uht.mode == "flat":
			from . import enmap
			for i, (shape, wcs) in enumerate(self.geometries):
				ls = np.asarray(self.get_ls(i))
				F = self.basis(i, ls)
				W = F**2/enmap.area(shape, wcs)
				Wtot = np.sum(W)
				filters.append(F); norms.append(Wtot**0.5); lmids.append(np.sum(W*ls)/Wtot)
			return filters, np.asarray(norms), np.asarray(lmids)
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
		if self.uht.mode == "flat":
			from . import enmap
			if fl is not None: raise NotImplementedError("Pre-filtering not yet implemented for flat-sky wavelets.")
			fmap = enmap.fft(map, normalize=False, wcs=self.uht.wcs)
			npix = map.shape[-2]*map.shape[-1]
			for i, (shape, wcs) in enumerate(self.geometries):
				if i in scales:
					fsmall = resample_fft(fmap, shape)
					f = self.filters[i]/(self.norms[i]*npix)
					if L.is_torch(fsmall):
						import torch
						fsmall *= torch.as_tensor(f, device=fsmall.device).to(fsmall.real.dtype)
					else: fsmall *= f.astype(fsmall.real.dtype)
					res = enmap.ifft(fsmall, normalize=False, wcs=wcs)
					owave.maps[i][...] = res.real
				elif fill_value is not None: owave.maps[i][...] = np.nan
			return owave
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
		if self.uht.mode == "flat":
			from . import enmap
			fomap = None
			for i, (shape, wcs) in enumerate(self.geometries):
				fsmall = enmap.fft(wave.maps[i], normalize=False, wcs=wcs)
				f = self.filters[i]*(self.norms[i]/(shape[-2]*shape[-1]))
				if L.is_torch(fsmall):
					import torch
					fsmall *= torch.as_tensor(f, device=fsmall.device).to(fsmall.real.dtype)
				else: fsmall = np.asarray(fsmall)*f.astype(fsmall.real.dtype)
				fomap = resample_fft(fsmall, self.uht.shape, fomap=fomap, add=fomap is not None)
			res = enmap.ifft(fomap, normalize=False, wcs=self.uht.wcs).real
			if omap is None: return res
			omap[...] = res
			return omap
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
	def get_ls(self, i):
		"""multipoles of scale i: the map of |l| of the scale's Fourier grid (flat) or 0..lmax (curved)"""
		if self.uht.mode == "flat":
			from . import enmap
			shape, wcs = self.geometries[i]
			# the scale's Fourier grid holds the low-frequency corners of the full grid: |l| of the FULL map at those modes
			ly, lx = enmap.laxes(self.uht.shape, self.uht.wcs)
			def corners(v, n):
				c = min(len(v), n); h = c//2
				out = np.zeros(n); out[:h] = v[:h]; out[n-(c-h):] = v[len(v)-(c-h):]
				return out
			return np.hypot(corners(ly, shape[-2])[:, None], corners(lx, shape[-1])[None, :])
		return self.uht.l

# ------------------------------------------------------------------ Haar-like pixel-space transform (reference wavelets.py:419-456)

def _block_reduce(a, bsize, axis, off):
	"""means over blocks of bsize along axis; incomplete blocks before `off` and at the end are kept (utils.block_reduce, inclusive)"""
	a = np.asarray(a); axis %= a.ndim
	nwhole = (a.shape[axis]-off)//bsize
	pre, mid, tail = np.split(a, [off, off+nwhole*bsize], axis)
	parts = []
	if pre.size > 0: parts.append(np.expand_dims(np.mean(pre, axis), axis))
	if mid.size > 0: parts.append(np.mean(mid.reshape(mid.shape[:axis]+(nwhole, bsize)+mid.shape[axis+1:]), axis+1))
	if tail.size > 0: parts.append(np.expand_dims(np.mean(tail, axis), axis))
	return np.concatenate(parts, axis) if parts else a

def _block_expand(a, bsize, osize, axis, off):
	"""nearest-neighbour inverse of _block_reduce (utils.block_expand, inclusive)"""
	a = np.asarray(a); axis %= a.ndim
	nwhole = (osize-off)//bsize; nrest = osize-off-nwhole*bsize
	pre, mid, tail = np.split(a, [int(off > 0), int(off > 0)+nwhole], axis)
	parts = []
	if pre.size > 0: parts.append(np.repeat(pre, off, axis))
	if mid.size > 0: parts.append(np.repeat(mid, bsize, axis))
	if tail.size > 0: parts.append(np.repeat(tail, nrest, axis))
	return np.concatenate(parts, axis)

class HaarTransform:
	"""Simple 2-d Haar-like transform in pixel space: every level halves the resolution by block means and keeps the
	difference to the re-expanded coarse map (reference wavelets.py:419-456).  ref = [dec, ra] keeps the blocks of different
	patches aligned (enmap.get_downgrade_offset, reference enmap.py:2026-2031); None: no offset."""
	def __init__(self, nlevel, ref=[0, 0]):
		self.nlevel, self.ref = nlevel, ref
	def _offset(self, shape, wcs):
		if self.ref is None: return np.zeros(2, int)
		y = geometry.ypix_of(wcs, np.rad2deg(self.ref[0]))
		x = wcs.wcs.crpix[0]-1 + (np.rad2deg(self.ref[1])-wcs.wcs.crval[0])/wcs.wcs.cdelt[0]
		return np.rint([y, x]).astype(int) % 2
	def _down_wcs(self, wcs, off):
		w = wcs.wcs
		# geometry of map[off::2] averaged in pairs: pixel centres move by half an input pixel; one extra pixel in front if off > 0
		crpix = (np.array(w.crpix, float) - off[::-1] - 0.5)/2 + 0.5 + (off[::-1] > 0)
		return geometry.CarWCS(w.crval, np.array(w.cdelt, float)*2, crpix, getattr(w, "ctype", ("RA---CAR", "DEC--CAR")))
	def map2wave(self, map, wcs=None):
		wcs = geometry.wcs_of(map, wcs)
		cur = np.asarray(map); omaps, geos = [], []
		for i in range(self.nlevel):
			off = self._offset(cur.shape, wcs)
			down = _block_reduce(_block_reduce(cur, 2, -2, off[0]), 2, -1, off[1])
			up = _block_expand(_block_expand(down, 2, cur.shape[-2], -2, off[0]), 2, cur.shape[-1], -1, off[1])
			omaps.append(geometry.ndmap(cur-up, wcs)); geos.append((cur.shape[-2:], wcs))
			cur, wcs = down, self._down_wcs(wcs, off)
		omaps.append(geometry.ndmap(cur, wcs)); geos.append((cur.shape[-2:], wcs))
		return multimap(omaps[::-1], geos[::-1])
	def wave2map(self, wave):
		omap = np.array(wave.maps[0])
		for i in range(1, wave.nmap):
			shape, wcs = wave.geometries[i]
			off = self._offset(shape, wcs)
			omap = np.asarray(wave.maps[i]) + _block_expand(_block_expand(omap, 2, shape[-2], -2, off[0]), 2, shape[-1], -1, off[1])
		return geometry.ndmap(omap, wave.geometries[-1][1])
