"""pixell_b200.uharm -- the Unified Harmonic Transform of pixell/uharm.py:8-209 on the B200 engine: one interface over
flat-sky FFTs (enmap.map2harm / harm2map, "phys" normalisation) and curved-sky SHTs (curvedsky.map2alm / alm2map),
so that filtering code is written once:

	uht  = UHT(shape, wcs)
	beam = uht.lprof2hprof(bl)
	omap = uht.harm2map(uht.hmul(beam, uht.map2harm(map)))

Provided: map2harm, harm2map and their adjoints, quad_weights, lprof2hprof, rprof2hprof / hprof2rprof (curved mode),
hprof2harm, hmul, hrand and hprof_rpow (curved mode), harm2powspec, sum_hprof, mean_hprof, beam2res, beam2rmax.  The
flat-sky radial profile helpers (profile2harm_flat(_2d), harm2profile_flat_2d: enmap's rbin / lbin / modrmap) are not provided.
"""
import numpy as np
from . import enmap, curvedsky, geometry, _lib as L

def res2lmax(res): return int(round(np.pi/res))

def estimate_distortion(shape, wcs):
	"""maximum relative scale difference across a cylindrical map (pixell/uharm.py:271-276)"""
	dec1, dec2 = geometry.dec_of(wcs, np.array([-0.5, shape[-2]-0.5]))
	rmin = min(np.cos(dec1), np.cos(dec2))
	rmax = 1 if not dec1*dec2 > 0 else max(np.cos(dec1), np.cos(dec2))
	return rmax/rmin-1

class UHT:
	def __init__(self, shape, wcs, mode="auto", lmax=None, max_distortion=0.1, niter=0):
		self.shape, self.wcs = tuple(shape[-2:]), wcs
		self.area = enmap.area(self.shape, self.wcs)
		self.fsky = self.area/(4*np.pi)
		if mode == "auto": mode = "flat" if estimate_distortion(shape, wcs) <= max_distortion else "curved"
		self.mode, self.quad, self.niter = mode, None, niter
		if mode == "flat":
			self.l = enmap.modlmap(self.shape, wcs)
			self.lmax = int(round(float(np.max(np.asarray(self.l)))))
			self.nper = 1/self.fsky
			self.ntot = self.nper*self.shape[-2]*self.shape[-1]
		elif mode == "curved":
			if lmax is None: lmax = res2lmax(np.min(np.abs(wcs.wcs.cdelt))*geometry.DEG)
			self.lmax = lmax
			self.l = np.arange(lmax+1)
			self.ainfo = curvedsky.alm_info(lmax=lmax)
			self.nper = 2*self.l+1
			self.ntot = np.sum(self.nper)
		else: raise ValueError("Unrecognized mode in UHT: '%s'" % (str(mode)))
	@property
	def npix(self): return self.shape[-2]*self.shape[-1]
	def _omap(self, harm):
		rdt = np.zeros(1, L.buffer_info(harm)[2]).real.dtype
		oshape = tuple(harm.shape[:-1])+self.shape
		if L.is_torch(harm):
			import torch
			return torch.zeros(oshape, dtype={np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32}[np.dtype(rdt)], device=harm.device)
		return geometry.zeros(oshape, self.wcs, rdt)
	def map2harm(self, map, spin=0):
		if self.mode == "flat": return enmap.map2harm(map, spin=spin, normalize="phys", wcs=self.wcs)
		return curvedsky.map2alm(map, ainfo=self.ainfo, spin=spin, niter=self.niter, wcs=self.wcs)
	def harm2map(self, harm, spin=0):
		if self.mode == "flat": return enmap.harm2map(harm, spin=spin, normalize="phys", wcs=self.wcs)
		return curvedsky.alm2map(harm, self._omap(harm), ainfo=self.ainfo, spin=spin, wcs=self.wcs)
	def harm2map_adjoint(self, map, spin=0):
		if self.mode == "flat": return enmap.harm2map_adjoint(map, spin=spin, normalize="phys", wcs=self.wcs)
		return curvedsky.alm2map_adjoint(map, ainfo=self.ainfo, spin=spin, wcs=self.wcs)
	def map2harm_adjoint(self, harm, spin=0):
		if self.mode == "flat": return enmap.map2harm_adjoint(harm, spin=spin, normalize="phys", wcs=self.wcs)
		return curvedsky.map2alm_adjoint(harm, self._omap(harm), ainfo=self.ainfo, spin=spin, niter=self.niter, wcs=self.wcs)
	def quad_weights(self):
		"""quadrature weights W broadcasting against maps: map2harm = harm2map_adjoint * W"""
		if self.quad is None:
			if self.mode == "flat": self.quad = geometry.pixsize_rows(self.shape, self.wcs)[:, None]
			else: self.quad = curvedsky.quad_weights(self.shape, self.wcs)[:, None]
		return self.quad
	def rprof2hprof(self, br, r):
		if self.mode == "flat": raise NotImplementedError("flat-sky radial profiles are not provided by pixell_b200")
		return curvedsky.profile2harm(br, r, lmax=self.lmax)
	def hprof2rprof(self, harm, r):
		if self.mode == "flat": raise NotImplementedError("flat-sky radial profiles are not provided by pixell_b200")
		return curvedsky.harm2profile(harm, r)
	def lprof2hprof(self, lprof):
		lprof = np.asarray(lprof)
		if self.mode == "flat":
			# linear interpolation of lprof at the map's |l|, zero beyond its last entry (utils.interpol order 1, constant border)
			l = np.asarray(self.l)
			i0 = np.clip(np.floor(l).astype(int), 0, lprof.shape[-1]-1); i1 = np.clip(i0+1, 0, lprof.shape[-1]-1)
			w = l-np.floor(l)
			res = lprof[..., i0]*(1-w) + lprof[..., i1]*w
			res = np.where(l <= lprof.shape[-1]-1, res, 0.0)
			return geometry.ndmap(res, self.wcs)
		if lprof.shape[-1] >= self.lmax+1: return lprof[..., :self.lmax+1]
		return np.concatenate([lprof, np.zeros(lprof.shape[:-1]+(self.lmax+1-lprof.shape[-1],), lprof.dtype)], -1)
	def hprof2harm(self, hprof):
		if self.mode == "flat": return hprof.copy()
		lval = np.zeros(self.ainfo.nelem, int)
		for m in range(self.ainfo.mmax+1): lval[self.ainfo.lm2ind(np.arange(m, self.lmax+1), m)] = np.arange(m, self.lmax+1)
		return np.asarray(hprof)[..., lval]
	def hmul(self, hprof, harm, inplace=False):
		"""hprof*harm -> harm; flat: hprof [ny,nx], [ncomp,ny,nx] or [ncomp,ncomp,ny,nx]; curved: [nl], [ncomp,nl] or [ncomp,ncomp,nl]"""
		if self.mode == "flat":
			if L.is_torch(harm):
				import torch
				h = torch.as_tensor(np.asarray(hprof), device=harm.device)
				res = h*harm if h.ndim <= 3 else torch.einsum("...abyx,...byx->...ayx", h.to(harm.dtype), harm)
			else: res = enmap.map_mul(np.asarray(hprof), harm)
			if inplace: harm[...] = res; return harm
			return res
		out = harm if inplace else None
		if not L.is_torch(harm): harm = np.asanyarray(harm).astype(np.result_type(harm, 0j), copy=False)
		return self.ainfo.lmul(harm, np.asarray(hprof), out=out)
	def harm2powspec(self, harm, harm2=None, patch=False):
		if self.mode == "flat":
			h2 = harm if harm2 is None else harm2
			return (harm*h2.conj()).real
		powspec = curvedsky.alm2cl(harm, harm2, ainfo=self.ainfo)
		if patch: powspec = powspec/self.fsky
		return powspec
	def sum_hprof(self, hprof):
		hprof = np.asanyarray(hprof)
		if self.mode == "flat": return np.sum(hprof*self.nper, (-2, -1))
		return np.sum(hprof*self.nper, -1)
	def mean_hprof(self, hprof): return self.sum_hprof(hprof)/self.ntot
	def hrand(self, hprof, seed=None):
		"""random realisation with harmonic profile hprof (pixell/uharm.py:166-172; curved mode: curvedsky.rand_alm)"""
		if self.mode == "flat": raise NotImplementedError("flat-sky hrand (enmap.rand_gauss_harm) is not provided by pixell_b200")
		return curvedsky.rand_alm(hprof, lmax=self.lmax, seed=seed)
	def hprof_rpow(self, hprof, power):
		"""hprof raised to `power` in real space: map2harm(harm2map(hprof)**power) on profiles (pixell/uharm.py:191-208)"""
		if self.mode == "flat": raise NotImplementedError("flat-sky hprof_rpow is not provided by pixell_b200")
		hprof = np.asarray(hprof)
		sigma = 1/max(1, np.where(hprof > np.max(hprof)*np.exp(-0.5))[0][-1])
		r = np.arange(0, 20*sigma, sigma/10)
		return self.rprof2hprof(self.hprof2rprof(hprof, r)**power, r)

def beam2res(br, r):
	"""a third of the beam's full width at half maximum (pixell/uharm.py:255-258)"""
	return 2*r[np.where(br >= br[0]*0.5)[0][-1]]/3

def beam2rmax(br, r, tol=1e-5, return_index=False):
	"""radius beyond which the beam stays below tol of its peak (pixell/uharm.py:260-263)"""
	imax = np.where(br >= br[0]*tol)[0][-1]
	return (r[imax], imax) if return_index else r[imax]
