"""pixell_b200.curvedsky -- the curved-sky hot path of pixell.curvedsky on the B200 engine.

Same names, argument meaning and error behaviour as the reference (pixell/curvedsky.py):
  alm_info :409-476   alm2map :83-164   alm2map_adjoint :166-172   map2alm :209-302
  map2alm_adjoint :304-310   rand_map :17-36   rand_alm :61-77   rand_alm_white :620-628
  almxfl :630-651   filter :653-669   alm2cl :672-712   transfer_alm :744-750
  analyse_geometry :1252-1306   get_method :478-488   quad_weights :492-505
for methods "2d" and "cyl" (cylindrical CAR maps, spin 0 and spin s, deriv, adjoints).
The "general" method (non-uniform FFT at the pixel centres), HEALPix rings and rotate_alm follow further down
(SURVEY.md 8f).

Differences that are deliberate:
  * maps may be numpy arrays / ndmaps (host) or torch CUDA tensors (pass wcs=); nothing is
    flipped, padded or copied on the host: the reference's map2buffer/buffer2map round trip
    (:1384-1411) is index arithmetic inside the ring-FFT kernels;
  * nthread is accepted and ignored.
"""
import numpy as np
from . import _lib as L, sht, cmisc, geometry
from .geometry import DEG

class ShapeError(Exception): pass

def nalm2lmax(nalm): return int((-1+(1+8*nalm)**0.5)/2)-1

class alm_info:
	"""Layout of an alm array (pixell/curvedsky.py:409-476)."""
	def __init__(self, lmax=None, mmax=None, nalm=None, stride=1, layout="triangular"):
		if lmax is not None: lmax = int(lmax)
		if mmax is not None: mmax = int(mmax)
		if nalm is not None: nalm = int(nalm)
		if isinstance(layout, str):
			if layout in ("triangular", "tri"):
				if lmax is None: lmax = nalm2lmax(nalm)
				if mmax is None: mmax = lmax
				m = np.arange(mmax+1)
				mstart = stride*(m*(2*lmax+1-m)//2)
			elif layout in ("rectangular", "rect"):
				if lmax is None: lmax = int(nalm**0.5)-1
				if mmax is None: mmax = lmax
				mstart = np.arange(mmax+1)*(lmax+1)*stride
			else: raise ValueError("unkonwn layout: %s" % layout)
		else: mstart = np.asarray(layout)
		self.lmax, self.mmax, self.stride = lmax, mmax, int(stride)
		self.nelem = int(np.max(mstart) + (lmax+1)*stride)
		self.nreal = lmax**2+2*lmax+2
		if nalm is not None:
			assert self.nelem == nalm, "lmax must be explicitly specified when lmax != mmax"
		self.mstart = mstart.astype(np.uint64, copy=False)
	@property
	def nl(self): return self.lmax+1
	@property
	def nm(self): return self.mmax+1
	def lm2ind(self, l, m): return (self.mstart[m].astype(int, copy=False)+l*self.stride).astype(int, copy=False)
	def transpose_alm(self, alm, out=None): return cmisc.transpose_alm(self, alm, out=out)
	def alm2cl(self, alm, alm2=None, dtype=None): return cmisc.alm2cl(self, alm, alm2=alm2, cl_dtype=dtype)
	def lmul(self, alm, lmat, out=None): return cmisc.lmul(self, alm, lmat, out=out)
	def __repr__(self): return "alm_info(lmax=%s,mmax=%s,mstart=%s)" % (str(self.lmax), str(self.mmax), str(self.mstart))

# ------------------------------------------------------------------ geometry analysis

class _Bunch(dict):
	__getattr__ = dict.__getitem__
	__setattr__ = dict.__setitem__

def _hasoff(val, off, tol): return abs((val-off+0.5) % 1 - 0.5) < tol

def _flipped(shape, wcs, flip):
	w = wcs.wcs
	crpix, cdelt = np.array(w.crpix, float), np.array(w.cdelt, float)
	if flip[0]: crpix[1] = shape[-2]+1-crpix[1]; cdelt[1] = -cdelt[1]
	if flip[1]: crpix[0] = shape[-1]+1-crpix[0]; cdelt[0] = -cdelt[0]
	return geometry.CarWCS(w.crval, cdelt, crpix)

def _is_cyl(wcs):
	"""plate carree with the equator as reference latitude: the only separable projection whose rows are
	equidistant in declination (what the ring kernels assume); anything else is the reference's "general" case"""
	ctype = getattr(wcs.wcs, "ctype", None)
	proj = "CAR" if ctype is None else str(ctype[0])[-3:].upper()
	return proj in ("CAR", "") and abs(wcs.wcs.crval[1]) < 1e-12

def get_ducc_geo(wcs, shape, tol=1e-6):
	"""pixell/curvedsky.py:1308-1347 on an already flipped (north-first, ra increasing) geometry."""
	nx = 360/wcs.wcs.cdelt[0]
	if not _hasoff(nx, 0, tol): return None
	y1, y2 = geometry.ypix_of(wcs, 90.0), geometry.ypix_of(wcs, -90.0)
	Ny = shape[-2]
	near = lambda a, b: abs(a-b) < tol
	if _hasoff(y1, 0, tol) and _hasoff(y2, 0, tol):
		if   near(y1, -1) and near(y2, Ny): name, o1, o2 = "F2", 1, 1
		elif near(y1, 0) and near(y2, Ny):  name, o1, o2 = "DH", 1, 0
		else: name, o1, o2 = "CC", 0, 0
	elif _hasoff(y1, 0.5, tol) and _hasoff(y2, 0.5, tol): name, o1, o2 = "F1", 0.5, 0.5
	elif _hasoff(y1, 0.5, tol) and _hasoff(y2, 0.0, tol): name, o1, o2 = "MW", 0.5, 0.0
	elif _hasoff(y1, 0.0, tol) and _hasoff(y2, 0.5, tol): name, o1, o2 = "MWflip", 0.0, 0.5
	else: return None
	ny = int(np.rint(y2-y1+1-o1-o2)); yoff = int(np.rint(-y1-o1))
	return _Bunch(name=name, nx=int(np.rint(nx)), ny=ny, pole_offs=[o1, o2], yoff=yoff, lmax=sht.maxlmax(name, ny))

def analyse_geometry(shape, wcs, tol=1e-6):
	"""pixell/curvedsky.py:1252-1306.  Adds phi0_user / xdir: azimuth of the caller's pixel x=0 and
	the sign of d(phi)/dx, which is how the engine handles x flips without copying."""
	divides = _hasoff(360/abs(wcs.wcs.cdelt[0]), 0, tol)
	if not _is_cyl(wcs) or not divides:
		return _Bunch(case="general", flip=[False, False], ducc_geo=None, ypad=(0, 0), xpad=(0, 0), phi0=0)
	flip = [bool(wcs.wcs.cdelt[1] > 0), bool(wcs.wcs.cdelt[0] < 0)]
	wwcs = _flipped(shape, wcs, flip)
	phi0 = float(geometry.ra_of(wwcs, 0))
	extra = dict(phi0_user=float(geometry.ra_of(wcs, 0)), xdir=-1 if flip[1] else 1, wwcs=wwcs)
	dg = get_ducc_geo(wwcs, shape, tol)
	if dg is not None and shape[-2] == dg.ny and shape[-1] == dg.nx and abs(dg.yoff) < tol:
		return _Bunch(case="2d", flip=flip, ducc_geo=dg, ypad=(0, 0), xpad=(0, 0), phi0=phi0, **extra)
	ypad = (dg.yoff, dg.ny-dg.yoff-shape[-2]) if dg is not None else (0, 0)
	nx = int(np.rint(360/wwcs.wcs.cdelt[0]))
	if shape[-1] == nx:
		return _Bunch(case="cyl", flip=flip, ducc_geo=dg, ypad=ypad, xpad=(0, 0), phi0=phi0, **extra)
	return _Bunch(case="partial", flip=flip, ducc_geo=dg, ypad=ypad, xpad=(0, nx-shape[-1]), phi0=phi0, **extra)

def get_method(shape, wcs, minfo=None, pix_tol=1e-6):
	"""pixell/curvedsky.py:478-488"""
	if minfo is None: minfo = analyse_geometry(shape, wcs, tol=pix_tol)
	if   minfo.case == "general": return "general"
	elif minfo.case == "2d":      return "2d"
	else:                         return "cyl"

def quad_weights(shape, wcs, pix_tol=1e-6):
	"""pixell/curvedsky.py:492-505: per-row quadrature weights in the caller's row order."""
	minfo = analyse_geometry(shape, wcs, tol=pix_tol)
	if minfo.ducc_geo is None or minfo.ducc_geo.name is None:
		raise ValueError("Quadrature weights not available for geometry %s,%s" % (str(shape), str(wcs)))
	w = _ring_weights(shape, wcs, minfo)
	return w[::-1] if minfo.flip[0] else w

def _ring_weights(shape, wcs, minfo):
	"""weights per ring in north-first order (pixell/curvedsky.py:852-861)"""
	if minfo.ducc_geo is not None and minfo.ducc_geo.name is not None:
		ny = shape[-2]+int(np.sum(minfo.ypad))
		w = sht.get_gridweights(minfo.ducc_geo.name, ny)
		w = w[minfo.ypad[0]:len(w)-minfo.ypad[1]]/minfo.ducc_geo.nx
		return w
	w = geometry.pixsize_rows(shape, wcs)
	return w[::-1] if minfo.flip[0] else w

def get_ring_info(shape, wcs, minfo=None):
	"""pixell/curvedsky.py:1170-1190 for the caller's (unflipped) array: one ring per row."""
	if minfo is None: minfo = analyse_geometry(shape, wcs)
	ny, nx = shape[-2:]
	theta = np.pi/2 - geometry.dec_of(wcs, np.arange(ny))
	nphi = int(np.rint(360/abs(wcs.wcs.cdelt[0])))
	return _Bunch(theta=theta, nphi=nphi, phi0=minfo.phi0_user, xdir=minfo.xdir, npix=nx,
		offsets=np.arange(ny, dtype=np.int64)*nx)

def spin_helper(spin, n):
	"""enmap.spin_helper (pixell/enmap.py:3378-3388)"""
	spin = np.array(spin).reshape(-1)
	scomp = 1+(spin != 0)
	ci, i1 = 0, 0
	while True:
		i2 = min(i1+scomp[ci], n)
		if i2-i1 != scomp[ci]: raise IndexError("Unpaired component in spin transform")
		yield int(spin[ci]), i1, i2
		if i2 == n: break
		i1 = i2; ci = (ci+1) % len(spin)

# ------------------------------------------------------------------ array plumbing

def _rdtype(a): return L.buffer_info(a)[2]
def _ctype_of(rdt): return np.result_type(rdt, 0j)

def _zeros_like_kind(ref, shape, dtype):
	return sht._empty_like(ref, shape, dtype)

def _astype(a, dtype):
	if L.is_torch(a):
		import torch
		td = {np.dtype(np.complex128): torch.complex128, np.dtype(np.complex64): torch.complex64,
			np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32}[np.dtype(dtype)]
		return a.to(td)
	return a.astype(dtype, copy=False)

def prepare_alm(alm=None, ainfo=None, lmax=None, pre=(), dtype=np.float64, convert=False, like=None):
	"""pixell/curvedsky.py:1413-1427"""
	ctype = _ctype_of(dtype)
	if alm is None:
		if ainfo is None:
			if lmax is None: raise ValueError("prepare_alm needs either alm, ainfo or lmax to be specified")
			ainfo = alm_info(lmax)
		alm = _zeros_like_kind(like, tuple(pre)+(ainfo.nelem,), ctype) if like is not None else np.zeros(tuple(pre)+(ainfo.nelem,), ctype)
	if ainfo is None: ainfo = alm_info(nalm=alm.shape[-1])
	if not convert and _rdtype(alm) != ctype:
		raise ValueError("alm had dtype '%s', but expected '%s'" % (str(_rdtype(alm)), str(ctype)))
	return _astype(alm, ctype), ainfo

def _atleast(a, n):
	while a.ndim < n: a = a[None]
	return a

def _contig(a, nlast):
	"""make the last nlast axes C-contiguous (copy only if needed)"""
	st = L.strides_elems(a); want = 1; ok = True
	for i in range(1, nlast+1):
		if a.shape[-i] != 1 and st[-i] != want: ok = False
		want *= a.shape[-i]
	if ok: return a, False
	return (a.contiguous() if L.is_torch(a) else np.ascontiguousarray(a)), True

def _comp_block(a, j1, j2, nlast):
	"""a[j1:j2] with contiguous trailing axes and a uniform component stride"""
	blk = a[j1:j2]
	blk, copied = _contig(blk, nlast)
	return blk, copied

# ------------------------------------------------------------------ transforms

def _plan_kwargs(shape, wcs, minfo, method, ainfo, lmax, mmax, weights=None):
	"""arguments that select the engine plan for this geometry"""
	if method == "2d":
		dg = minfo.ducc_geo
		return dict(kind="2d", geometry=dg.name, phi0=minfo.phi0_user, flip_y=minfo.flip[0], flip_x=minfo.flip[1],
			lmax=lmax, mmax=mmax, mstart=np.asarray(ainfo.mstart)[:mmax+1], lstride=ainfo.stride, ntheta=shape[-2], nphi=shape[-1])
	ri = get_ring_info(shape, wcs, minfo)
	return dict(kind="rings", theta=ri.theta, nphi=ri.nphi, phi0=ri.phi0, ringstart=ri.offsets, xdir=ri.xdir, npix=ri.npix,
		lmax=lmax, mmax=mmax, mstart=np.asarray(ainfo.mstart)[:mmax+1], lstride=ainfo.stride, weight=weights)

def _synth(pk, alm, map, spin, mode="STANDARD", adjoint=False):
	"""alm[nca, nalm] <-> map[ncm, ny, nx] through the right engine entry point"""
	if pk["kind"] == "2d":
		kw = {k: pk[k] for k in ("geometry", "phi0", "flip_y", "flip_x", "lmax", "mmax", "mstart", "lstride")}
		if adjoint: sht.adjoint_synthesis_2d(map=map, alm=alm, spin=spin, mode=mode, **kw)
		else:       sht.synthesis_2d(alm=alm, map=map, spin=spin, mode=mode, **kw)
	else:
		kw = {k: pk[k] for k in ("theta", "nphi", "phi0", "ringstart", "xdir", "npix", "lmax", "mmax", "mstart", "lstride")}
		flat = map.reshape(map.shape[0], -1)
		if adjoint: sht.adjoint_synthesis(map=flat, alm=alm, spin=spin, mode=mode, weight=pk.get("weight"), **kw)
		else:       sht.synthesis(alm=alm, map=flat, spin=spin, mode=mode, **kw)

def _plan_of(pk):
	if pk["kind"] == "2d":
		return sht.plan_2d(pk["geometry"], pk["ntheta"], pk["nphi"], pk["phi0"], pk["lmax"], pk["mmax"], pk["mstart"], pk["lstride"], pk["flip_y"], pk["flip_x"])
	return sht.plan_rings(pk["theta"], pk["nphi"], pk["phi0"], pk["ringstart"], pk["lmax"], pk["mmax"], pk["mstart"], pk["lstride"],
		pk.get("weight"), pk["xdir"], pk["npix"])

def _grouped(pk, op, groups, alm2, map3):
	"""All spin groups of one [ncomp, nalm] / [ncomp, ny, nx] pair in a single engine call when no copies are
	needed (the usual case): lets the engine overlap host<->device copies of one group with the kernels of the
	next.  Returns False when the arrays need the per-group path."""
	if len(groups) < 2 or len(groups) > 8: return False
	sa, sm = L.strides_elems(alm2), L.strides_elems(map3)
	if sa[-1] != 1 or sm[-1] != 1 or sm[-2] != map3.shape[-1]: return False
	sht.run_groups(_plan_of(pk), op, groups, alm2, map3)
	return True

def _linear_car(wcs):
	"""True when pixel centres follow the linear plate-carree formulas of geometry.py (dec_of / ra_of / pixsize_rows):
	CAR (or no projection given) with the equator as reference latitude.  Other projections (CEA, TAN, ...; CAR with
	crval[1] != 0 is oblique) need a real WCS library for enmap.pix2sky / pixsizemap (pixell/curvedsky.py:1355-1382)."""
	return _is_cyl(wcs)

def _require_linear_car(wcs, what):
	if not _linear_car(wcs):
		ctype = getattr(wcs.wcs, "ctype", ["?"])
		raise NotImplementedError("%s: pixel positions of projection %s with crval[1]=%g are not the linear plate-carree ones "
			"pixell_b200.geometry computes; pass locinfo= (pixel centres from a WCS library) to the *_general functions instead"
			% (what, str(ctype[0]), wcs.wcs.crval[1]))

def _check_method(method, minfo, shape):
	if method == "general": return          # handled by *_general, which locate the pixel centres (linear CAR only, or locinfo=)
	if method not in ("2d", "cyl"): raise ValueError("Unrecognized alm2map method '%s'" % str(method))
	if minfo.case == "general": raise NotImplementedError("non-cylindrical geometry: only the reference's 'general' method applies")
	if method == "2d" and minfo.case != "2d":
		raise NotImplementedError("method='2d' on a map that is not a full ducc grid needs padding; use method='cyl'")

def alm2map(alm, map, spin=[0,2], deriv=False, adjoint=False, copy=False, method="auto", ainfo=None,
		verbose=False, nthread=None, epsilon=1e-6, pix_tol=1e-6, locinfo=None, tweak=False, wcs=None):
	"""Spherical harmonics synthesis (pixell/curvedsky.py:83-164).  alm[...,ncomp,nelem] -> map[...,ncomp,ny,nx];
	deriv=True: alm[...,nelem] -> map[...,2,ny,nx] = (d/ddec, d/dra / cos(dec)); adjoint=True applies the transpose
	(map -> alm)."""
	wcs = geometry.wcs_of(map, wcs)
	minfo = analyse_geometry(map.shape, wcs, tol=pix_tol)
	if method == "auto": method = get_method(map.shape, wcs, minfo=minfo)
	_check_method(method, minfo, map.shape)
	if verbose: print("method: %s" % method)
	if method == "general":
		return alm2map_general(alm, map, ainfo=ainfo, spin=spin, deriv=deriv, copy=copy, adjoint=adjoint, locinfo=locinfo,
			epsilon=None if epsilon == 1e-6 else epsilon, wcs=wcs)
	if copy:
		if adjoint and alm is not None: alm = alm.clone() if L.is_torch(alm) else alm.copy()
		else: map = map.clone() if L.is_torch(map) else map.copy()
	rdt = _rdtype(map)
	if adjoint: alm, ainfo = prepare_alm(alm=alm, ainfo=ainfo, pre=map.shape[:-2] if not deriv else map.shape[:-3], dtype=rdt, convert=False, like=map)
	else:       alm, ainfo = prepare_alm(alm=alm, ainfo=ainfo, pre=(), dtype=rdt, convert=True)
	alm_full = _atleast(alm, 2 if deriv else 3)
	map_full = _atleast(map if L.is_torch(map) else np.asarray(map), 4)
	if deriv:
		assert map_full.shape[-3] == 2, "map must have shape [...,2,ny,nx] when deriv is True"
		assert tuple(map_full.shape[:-3]) == tuple(alm_full.shape[:-1]), "map and alm must agree on pre-dimensions"
	else:
		assert tuple(map_full.shape[:-2]) == tuple(alm_full.shape[:-1]), "map and alm must agree on pre-dimensions"
	pk = _plan_kwargs(map.shape, wcs, minfo, method, ainfo, ainfo.lmax, ainfo.mmax)
	for I in np.ndindex(*map_full.shape[:-3]):
		if deriv:
			# ducc returns (d/dtheta, 1/sin(theta) d/dphi); flipping the first gives d/ddec (curvedsky.py:918-920)
			a = _contig(alm_full[I][None], 1)[0]
			m, mcopied = _contig(map_full[I], 2)
			if adjoint:
				mm = m.clone() if L.is_torch(m) else m.copy(); mm[0] *= -1
				_synth(pk, a, mm, 1, "DERIV1", adjoint=True)
				if a is not alm_full[I][None] and not _same(a, alm_full[I]): alm_full[I] = a[0]
			else:
				_synth(pk, a, m, 1, "DERIV1")
				m[0] *= -1
				if mcopied: map_full[I] = m
		else:
			groups = list(spin_helper(spin, alm_full.shape[-2]))
			# all spin groups in one engine call: the copies of one group then overlap the kernels of the next
			if pk["kind"] != "rings" or pk.get("weight") is None:
				if _grouped(pk, "adjoint_synthesis" if adjoint else "synthesis", groups, alm_full[I], map_full[I]): continue
			for s, j1, j2 in groups:
				a, acopied = _comp_block(alm_full[I], j1, j2, 1)
				m, mcopied = _comp_block(map_full[I], j1, j2, 2)
				_synth(pk, a, m, s, adjoint=adjoint)
				if adjoint and acopied: alm_full[I][j1:j2] = a
				if not adjoint and mcopied: map_full[I][j1:j2] = m
	return alm if adjoint else map

# ------------------------------------------------------------------ HEALPix and radial ring sets

def npix2nside(npix): return int(round((npix/12)**0.5))

def prepare_healmap(healmap, nside=None, pre=(), dtype=np.float64):
	if healmap is not None: return healmap
	return np.zeros(tuple(pre)+(12*nside**2,), dtype)

def get_ring_info_healpix(nside, rings=None):
	"""theta, nphi, phi0 and pixel offset of every HEALPix ring (reference curvedsky.py:1192-1222)"""
	nside = int(nside)
	rings = np.arange(4*nside-1) if rings is None else np.asarray(rings)
	nring, npix = len(rings), 12*nside**2
	theta, phi0, nphi = np.zeros(nring), np.zeros(nring), np.zeros(nring, np.uint64)
	rings = rings+1                              # one-based
	north = np.where(rings > 2*nside, 4*nside-rings, rings)
	cap = np.where(north < nside)[0]
	theta[cap] = 2*np.arcsin(north[cap]/(6**0.5*nside))
	nphi[cap] = 4*north[cap]
	phi0[cap] = np.pi/(4*north[cap])
	rest = np.where(north >= nside)[0]
	theta[rest] = np.arccos((2*nside-north[rest])*(8*nside/npix))
	nphi[rest] = 4*nside
	phi0[rest] = np.pi/(4*nside)*(((north[rest]-nside) & 1) == 0)
	south = np.where(north != rings)[0]
	theta[south] = np.pi-theta[south]
	offsets = np.concatenate([[0], np.cumsum(nphi)[:-1]]).astype(np.uint64)
	return _Bunch(theta=theta, nphi=nphi, phi0=phi0, offsets=offsets, stride=np.ones(nring, np.int32), npix=npix, nrow=nring)

def get_ring_info_radial(r):
	"""one single-pixel ring per colatitude r (reference curvedsky.py:1224-1234): radially symmetric (mmax = 0) transforms"""
	theta = np.asarray(r, dtype=np.float64)
	assert theta.ndim == 1, "r must be one-dimensional!"
	n = len(theta)
	return _Bunch(theta=theta, nphi=np.ones(n, np.uint64), phi0=np.zeros(n), offsets=np.arange(n, dtype=np.uint64),
		stride=np.ones(n, np.int32), npix=n, nrow=n)

def apply_minfo_theta_lim(minfo, theta_min=None, theta_max=None):
	if theta_min is None and theta_max is None: return minfo
	mask = np.full(minfo.nrow, True, bool)
	if theta_min is not None: mask &= minfo.theta >= theta_min
	if theta_max is not None: mask &= minfo.theta <= theta_max
	res = _Bunch(minfo)
	for key in ["theta", "nphi", "phi0", "offsets"]: res[key] = res[key][mask]
	return res

def _ring_kwargs(rinfo, ainfo):
	return dict(theta=rinfo.theta, nphi=rinfo.nphi, phi0=rinfo.phi0, ringstart=rinfo.offsets, lmax=ainfo.lmax, mmax=ainfo.mmax,
		mstart=ainfo.mstart, lstride=ainfo.stride)

def alm2map_healpix(alm, healmap=None, spin=[0,2], deriv=False, adjoint=False, copy=False, ainfo=None, nside=None,
		theta_min=None, theta_max=None, nthread=None):
	"""alm[..., ncomp, nalm] -> HEALPix map[..., ncomp, npix] (RING order), or its transpose (reference curvedsky.py:312-353)."""
	rdt = np.zeros(1, _rdtype(alm)).real.dtype
	if ainfo is None: ainfo = alm_info(nalm=alm.shape[-1])
	healmap = prepare_healmap(healmap, nside, alm.shape[:-1] if not deriv else alm.shape[:-1]+(2,), rdt)
	if copy:
		if adjoint: alm = alm.copy()
		else: healmap = healmap.copy()
	alm_full = _atleast(alm, 2 if deriv else 3)
	map_full = _atleast(healmap, 3)
	if deriv and (alm_full.shape[:-1] != map_full.shape[:-2] or map_full.shape[-2] != 2):
		raise ValueError("When deriv is True, alm must have shape [...,nelem] and map shape [...,2,npix]")
	if not deriv and (alm_full.shape[:-1] != map_full.shape[:-1]):
		raise ValueError("alm must have shape [...,[ncomp,]nelem] and map shape [...,[ncomp,]npix]")
	func = sht.adjoint_synthesis if adjoint else sht.synthesis
	rinfo = apply_minfo_theta_lim(get_ring_info_healpix(npix2nside(map_full.shape[-1])), theta_min, theta_max)
	if (theta_min is not None or theta_max is not None) and not adjoint: map_full[:] = 0
	kw = _ring_kwargs(rinfo, ainfo)
	ctype = np.result_type(rdt, 0j)
	for I in np.ndindex(*map_full.shape[:-2]):
		if deriv:
			a = np.ascontiguousarray(alm_full[I][None]).astype(ctype, copy=False); m = np.ascontiguousarray(map_full[I])
			if adjoint:
				m = m.copy(); m[0] *= -1
				func(alm=a, map=m, mode="DERIV1", spin=1, **kw); alm_full[I] = a[0]
			else:
				func(alm=a, map=m, mode="DERIV1", spin=1, **kw)
				m[0] *= -1; map_full[I] = m
		else:
			for s, j1, j2 in spin_helper(spin, alm_full[I].shape[-2]):
				Ij = I+(slice(j1, j2),)
				a = np.ascontiguousarray(alm_full[Ij]).astype(ctype, copy=False); m = np.ascontiguousarray(map_full[Ij])
				func(alm=a, map=m, spin=s, **kw)
				if adjoint: alm_full[Ij] = a
				else: map_full[Ij] = m
	return alm if adjoint else healmap

def map2alm_healpix(healmap, alm=None, ainfo=None, lmax=None, spin=[0,2], weights=None, deriv=False, copy=False, verbose=False,
		adjoint=False, niter=0, theta_min=None, theta_max=None, nthread=None):
	"""HEALPix map -> alm with pixel-area weights and niter Jacobi iterations, like healpy's map2alm
	(reference curvedsky.py:355-405)."""
	if deriv: raise NotImplementedError("map2alm_healpix with deriv=True is broken")
	if copy:
		if adjoint: healmap = healmap.copy()
		elif alm is not None: alm = alm.copy()
	pre = healmap.shape[:-1]
	alm, ainfo = prepare_alm(alm=alm, ainfo=ainfo, lmax=lmax, pre=pre, dtype=_rdtype(healmap), convert=adjoint, like=healmap)
	alm_full = _atleast(alm, 3)
	map_full = _atleast(healmap, 3)
	rinfo = apply_minfo_theta_lim(get_ring_info_healpix(npix2nside(map_full.shape[-1])), theta_min, theta_max)
	kw = _ring_kwargs(rinfo, ainfo)
	if weights is None: weights = 4*np.pi/rinfo.npix
	for I in np.ndindex(*map_full.shape[:-2]):
		for s, j1, j2 in spin_helper(spin, alm_full.shape[-2]):
			Ij = I+(slice(j1, j2),)
			def Y(a): return sht.synthesis(map=np.zeros_like(np.ascontiguousarray(map_full[Ij])), alm=np.ascontiguousarray(a), spin=s, **kw)
			def YT(m): return sht.adjoint_synthesis(map=np.ascontiguousarray(m), spin=s, **kw)
			def YTW(m): return YT(m*weights)
			def WY(a): return Y(a)*weights
			if adjoint:
				a = np.ascontiguousarray(alm_full[Ij]); x = WY(a)
				for it in range(niter): x -= WY(YT(x)-a)
				map_full[Ij] = x
			else:
				y = np.ascontiguousarray(map_full[Ij]); x = YTW(y)
				for it in range(niter): x -= YTW(Y(x)-y)
				alm_full[Ij] = x
	return healmap if adjoint else alm

def profile2harm(br, r, lmax=None, oversample=1, left=None, right=None):
	"""Radial profile br[..., nr] at ascending radii r -> bl[..., nl]: an mmax = 0 transform on single-pixel
	Clenshaw-Curtis rings (reference curvedsky.py:1544-1577)."""
	br, r = np.asarray(br), np.asarray(r)
	dr = (r[-1]-r[0])/(len(r)-1)
	nfull = int(round(np.pi/dr))+1
	dr = np.pi/(nfull-1)
	ncut = int(np.ceil(r[-1]/dr))
	if lmax is None: lmax = int(nfull//2-1)
	l = np.arange(lmax+1)
	rinfo = get_ring_info_radial(np.arange(ncut)*dr)
	weights = sht.get_gridweights("CC", nfull)[:ncut]
	harm = np.zeros(br.shape[:-1]+(lmax+1,), br.dtype)
	for I in np.ndindex(*br.shape[:-1]):
		map = np.interp(rinfo.theta, r, br[I], left=left, right=right).reshape(1, -1)
		alm = sht.adjoint_synthesis(map=np.ascontiguousarray(map*weights, dtype=np.float64), theta=rinfo.theta, nphi=rinfo.nphi,
			phi0=rinfo.phi0, ringstart=rinfo.offsets, spin=0, lmax=lmax, mmax=0)[0]
		harm[I] = alm.real*(4*np.pi/(2*l+1))**0.5
	return harm

def harm2profile(bl, r):
	"""bl[..., nl] -> br[..., nr] = sum_l bl (2l+1)/(4 pi) P_l(cos r) (reference curvedsky.py:1579-1593)"""
	bl, r = np.asarray(bl), np.asarray(r)
	l = np.arange(bl.shape[-1])
	rinfo = get_ring_info_radial(r.reshape(-1))
	alm = (bl*((2*l+1)/(4*np.pi))**0.5).astype(np.complex128)
	br = np.zeros(bl.shape[:-1]+(r.size,), np.float64)
	for I in np.ndindex(*bl.shape[:-1]):
		br[I] = sht.synthesis(alm=np.ascontiguousarray(alm[I][None]), theta=rinfo.theta, nphi=rinfo.nphi, phi0=rinfo.phi0,
			ringstart=rinfo.offsets, spin=0, lmax=bl.shape[-1]-1, mmax=0)[0]
	return br.astype(bl.dtype, copy=False) if bl.dtype.kind == "f" else br

# zyz Euler angles between coordinate systems (pixell/curvedsky.py:714-716)
euler_angs = {}
euler_angs[("gal", "equ")] = np.array([57.06793215, 62.87115487, -167.14056929])*DEG
euler_angs[("equ", "gal")] = -euler_angs[("gal", "equ")][::-1]

def rotate_alm(alm, psi, theta, phi, lmax=None, method="auto", nthread=None, inplace=False):
	"""Rotate alm[..., :] by the zyz Euler angles psi, theta, phi (reference curvedsky.py:717-742, ducc0.sht.rotate_alm):
	the field is rotated actively by R = Rz(phi) Ry(theta) Rz(psi), f'(x) = f(R^-1 x), every component as a scalar.
	Instead of Wigner matrices the engine evaluates f at the back-rotated nodes of a Clenshaw-Curtis grid that carries
	lmax exactly (K8, 1e-12) and analyses the result (exact quadrature), which is the same operator for band-limited f."""
	import torch
	if lmax is None: lmax = nalm2lmax(alm.shape[-1])
	ainfo = alm_info(lmax)
	if ainfo.nelem != alm.shape[-1]: raise ValueError("rotate_alm needs the triangular layout with mmax = lmax")
	tor = L.is_torch(alm)
	out = alm if inplace else (alm.clone() if tor else alm.copy())
	dev = torch.device("cuda", L.init())
	nt, nphi = lmax+2, sht._fast_len(2*lmax+2)
	th = torch.arange(nt, device=dev, dtype=torch.float64)*(np.pi/(nt-1))
	ph = torch.arange(nphi, device=dev, dtype=torch.float64)*(2*np.pi/nphi) - phi       # Rz(-phi)
	st, ct = torch.sin(th)[:, None], torch.cos(th)[:, None]
	x, y, z = st*torch.cos(ph)[None, :], st*torch.sin(ph)[None, :], ct.expand(nt, nphi)
	c, sn = np.cos(theta), np.sin(theta)                                                 # Ry(-theta)
	x2, z2 = x*c - z*sn, x*sn + z*c
	loc = torch.stack([torch.atan2(torch.sqrt(x2*x2 + y*y), z2), torch.atan2(y, x2) - psi], -1).reshape(-1, 2).contiguous()
	del x, y, z, x2, z2
	flat = out.reshape(-1, out.shape[-1])
	for i in range(flat.shape[0]):
		a = flat[i]
		a = (a if tor else torch.from_numpy(np.ascontiguousarray(a))).to(dev).to(torch.complex128).reshape(1, -1)
		m = sht.synthesis_general(alm=a, loc=loc, spin=0, lmax=lmax).reshape(1, nt, nphi)
		b = sht.analysis_2d(map=m, spin=0, lmax=lmax, geometry="CC")
		if tor: flat[i] = b[0].to(flat.dtype)
		else: flat[i] = b[0].cpu().numpy().astype(flat.dtype, copy=False)
	return out

def alm2map_raw_general(alm, map, loc, ainfo=None, spin=[0,2], deriv=False, copy=False, verbose=False, adjoint=False, nthread=None, epsilon=None):
	"""alm[..., ncomp, nelem] <-> map[..., ncomp, npos] (deriv: alm[..., nelem], map[..., 2, npos]) at loc[npos, 2] =
	(codec, ra) by the non-uniform-FFT synthesis or, with adjoint=True, its transpose (reference curvedsky.py:993-1016)."""
	if copy:
		if adjoint: alm = alm.clone() if L.is_torch(alm) else alm.copy()
		else: map = map.clone() if L.is_torch(map) else map.copy()
	if ainfo is None: ainfo = alm_info(nalm=alm.shape[-1])
	if epsilon is None: epsilon = 1e-10 if _rdtype(map) == np.float64 else 1e-6
	epsilon = max(epsilon, 1e-12)
	alm_full = _atleast(alm, 2 if deriv else 3)
	map_full = _atleast(map, 3)
	kw = dict(loc=loc, lmax=ainfo.lmax, mmax=ainfo.mmax, mstart=ainfo.mstart, lstride=ainfo.stride, epsilon=epsilon)
	ctype = np.complex128
	for I in np.ndindex(*map_full.shape[:-2]):
		if deriv:
			if adjoint:
				m = _astype(map_full[I], np.float64)
				m = m.clone() if L.is_torch(m) else m.copy()
				m[0] *= -1
				_assign(alm_full[I], sht.adjoint_synthesis_general(map=m, spin=1, mode="DERIV1", **kw)[0])
			else:
				out = sht.synthesis_general(alm=_astype(alm_full[I][None], ctype), spin=1, mode="DERIV1", **kw)
				out[0] *= -1                          # theta derivative -> dec derivative (reference :1010-1011)
				_assign(map_full[I], out)
		else:
			for s, j1, j2 in spin_helper(spin, alm_full.shape[-2]):
				Ij = I+(slice(j1, j2),)
				if adjoint: _assign(alm_full[Ij], sht.adjoint_synthesis_general(map=_astype(map_full[Ij], np.float64), spin=s, **kw))
				else: _assign(map_full[Ij], sht.synthesis_general(alm=_astype(alm_full[Ij], ctype), spin=s, **kw))
	return alm if adjoint else map

def calc_locinfo(shape, wcs, bsize=1000):
	"""(codec, ra) of every pixel centre, [npix, 2] (reference curvedsky.py:1355-1382; separable CAR maps: no invalid pixels)"""
	dec = geometry.dec_of(wcs, np.arange(shape[-2])); ra = geometry.ra_of(wcs, np.arange(shape[-1]))
	loc = np.empty((shape[-2], shape[-1], 2))
	loc[..., 0] = (np.pi/2-dec)[:, None]
	loc[..., 1] = np.mod(ra, 2*np.pi)[None, :]
	return _Bunch(loc=loc.reshape(-1, 2), mask=np.ones(tuple(shape[-2:]), bool), masked=False)

def alm2map_general(alm, map, ainfo=None, spin=[0,2], deriv=False, copy=False, verbose=False, adjoint=False, nthread=None,
		locinfo=None, epsilon=None, wcs=None):
	"""alm2map through the arbitrary-position synthesis at the pixel centres (reference curvedsky.py:796-820): works on any
	pixelisation whose pixel centres are known, at a few times the cost of the ring methods."""
	wcs = geometry.wcs_of(map, wcs)
	if L.is_torch(map): raise NotImplementedError("alm2map_general: pass numpy maps (use sht.synthesis_general directly for device tensors)")
	if copy:
		if adjoint and alm is not None: alm = alm.copy()
		else: map = map.copy()
	if adjoint: alm, ainfo = prepare_alm(alm=alm, ainfo=ainfo, pre=map.shape[:-2] if not deriv else map.shape[:-3], dtype=_rdtype(map), convert=False, like=map)
	if locinfo is None:
		_require_linear_car(wcs, "alm2map_general")
		locinfo = calc_locinfo(map.shape, wcs)
	mview = np.asarray(map)
	for I in np.ndindex(*mview.shape[:-3]):
		tmap = np.ascontiguousarray(mview[I].reshape(mview[I].shape[:-2]+(-1,)))
		alm2map_raw_general(alm[I], tmap, locinfo.loc, ainfo=ainfo, spin=spin, deriv=deriv, epsilon=epsilon, adjoint=adjoint)
		if not adjoint: mview[I] = tmap.reshape(mview[I].shape)
	return alm if adjoint else map

def map2alm_raw_general(map, loc, alm=None, ainfo=None, lmax=None, spin=[0,2], weights=None, deriv=False, copy=False, verbose=False,
		adjoint=False, nthread=None, niter=0, epsilon=None):
	"""map[..., ncomp, npix] at loc -> alm = Y^T (W map), refined by niter Jacobi iterations, or the adjoint of that
	(reference curvedsky.py:1088-1120)."""
	if deriv: raise NotImplementedError("map2alm_raw_general: deriv=True is not provided")
	if epsilon is None: epsilon = 1e-10 if _rdtype(map) == np.float64 else 1e-6
	epsilon = max(epsilon, 1e-12)
	if ainfo is None: ainfo = alm_info(lmax=lmax) if alm is None else alm_info(nalm=alm.shape[-1])
	if weights is None: weights = np.ones(1)
	alm_full, map_full = _atleast(alm, 3), _atleast(map, 3)
	kw = dict(loc=loc, lmax=ainfo.lmax, mmax=ainfo.mmax, mstart=ainfo.mstart, lstride=ainfo.stride, epsilon=epsilon)
	for I in np.ndindex(*map_full.shape[:-2]):
		for s, j1, j2 in spin_helper(spin, alm_full.shape[-2]):
			Ij = I+(slice(j1, j2),)
			def Y(a): return sht.synthesis_general(alm=np.ascontiguousarray(a, dtype=np.complex128), spin=s, **kw)
			def YT(m): return sht.adjoint_synthesis_general(map=np.ascontiguousarray(m, dtype=np.float64), spin=s, **kw)
			def YTW(m): return YT(m*weights)
			def WY(a): return Y(a)*weights
			if adjoint:
				a = alm_full[Ij]; x = WY(a)
				for it in range(niter): x -= WY(YT(x)-a)
				map_full[Ij] = x
			else:
				y = map_full[Ij]; x = YTW(y)
				for it in range(niter): x -= YTW(Y(x)-y)
				alm_full[Ij] = x
	return map if adjoint else alm

def map2alm_general(map, alm=None, ainfo=None, minfo=None, lmax=None, spin=[0,2], weights=None, deriv=False, copy=False,
		verbose=False, adjoint=False, nthread=None, locinfo=None, epsilon=None, niter=0, wcs=None):
	"""map2alm with pixel-area weights at arbitrary pixel centres (reference curvedsky.py:875-898)"""
	wcs = geometry.wcs_of(map, wcs)
	if L.is_torch(map): raise NotImplementedError("map2alm_general: pass numpy maps")
	if adjoint:
		if copy and map is not None: map = map.copy()
	elif copy and alm is not None: alm = alm.copy()
	alm, ainfo = prepare_alm(alm=alm, ainfo=ainfo, lmax=lmax, pre=map.shape[:-2], dtype=_rdtype(map), convert=adjoint, like=map)
	if locinfo is None or weights is None: _require_linear_car(wcs, "map2alm_general")
	if locinfo is None: locinfo = calc_locinfo(map.shape, wcs)
	if weights is None: weights = np.repeat(geometry.pixsize_rows(map.shape, wcs), map.shape[-1]).astype(_rdtype(map), copy=False)
	mview = np.asarray(map)
	for I in np.ndindex(*mview.shape[:-3]):
		tmap = np.ascontiguousarray(mview[I].reshape(mview[I].shape[:-2]+(-1,)))
		map2alm_raw_general(tmap, locinfo.loc, alm[I], ainfo=ainfo, lmax=lmax, spin=spin, deriv=deriv, weights=weights,
			adjoint=adjoint, niter=niter, epsilon=epsilon)
		if adjoint: mview[I] = tmap.reshape(mview[I].shape)
	return map if adjoint else alm

def alm2map_pos(alm, pos=None, loc=None, ainfo=None, map=None, spin=[0,2], deriv=False, copy=False, verbose=False, adjoint=False, nthread=None, epsilon=None):
	"""Like alm2map, but evaluated at arbitrary positions (reference curvedsky.py:174-207):
	pos: [{dec,ra},...] radians, or loc: [...,{codec,ra}] radians.  adjoint=True: map -> alm (alm must be given)."""
	if adjoint:
		if copy and alm is not None: alm = alm.copy()
	elif copy and map is not None: map = map.copy()
	if loc is None:
		loc = np.moveaxis(np.asarray(pos, dtype=np.float64), 0, -1).copy(order="C")
		loc[..., 0] *= -1
		loc[..., 0] += np.pi/2
		loc[loc[..., 1] < 0, 1] += 2*np.pi
	lpre = loc.shape[:-1]
	loc = loc.reshape(-1, 2)
	if deriv: oshape = alm.shape[:-1]+(2, len(loc))
	else:     oshape = alm.shape[:-1]+(len(loc),)
	if map is None: map = np.zeros(oshape, np.zeros(1, _rdtype(alm)).real.dtype)
	map = map.reshape(oshape)
	if map.ndim < 2: alm2map_raw_general(alm[None], map[None], loc, ainfo=ainfo, spin=spin, deriv=deriv, epsilon=epsilon, adjoint=adjoint)
	else:
		for I in np.ndindex(*map.shape[:-2]):
			alm2map_raw_general(alm[I], map[I], loc, ainfo=ainfo, spin=spin, deriv=deriv, epsilon=epsilon, adjoint=adjoint)
	map = map.reshape(map.shape[:-1]+lpre)
	return alm if adjoint else map

def _same(a, b):
	return L.buffer_info(a)[0] == L.buffer_info(b)[0]

def alm2map_adjoint(map, alm=None, spin=[0,2], deriv=False, copy=False, method="auto", ainfo=None,
		verbose=False, nthread=None, epsilon=None, pix_tol=1e-6, locinfo=None, wcs=None):
	"""pixell/curvedsky.py:166-172"""
	return alm2map(alm, map, spin=spin, deriv=deriv, adjoint=True, copy=copy, method=method, ainfo=ainfo,
		verbose=verbose, nthread=nthread, pix_tol=pix_tol, wcs=wcs)

def map2alm(map, alm=None, lmax=None, spin=[0,2], deriv=False, adjoint=False, copy=False, method="auto",
		ainfo=None, verbose=False, nthread=None, niter=0, epsilon=None, pix_tol=1e-6, weights=None,
		locinfo=None, tweak=False, wcs=None):
	"""Spherical harmonics analysis (pixell/curvedsky.py:209-302).  method "2d": exact quadrature
	(ducc analysis_2d), lmax clipped to what the grid supports (:1027); method "cyl": alm = Y^T W map
	refined by niter Jacobi iterations (:1079-1084, :1122-1136)."""
	wcs = geometry.wcs_of(map, wcs)
	minfo = analyse_geometry(map.shape, wcs, tol=pix_tol)
	if method == "auto": method = get_method(map.shape, wcs, minfo=minfo)
	_check_method(method, minfo, map.shape)
	if verbose: print("method: %s" % method)
	if method == "general":
		return map2alm_general(map, alm=alm, ainfo=ainfo, lmax=lmax, spin=spin, weights=weights, deriv=deriv, copy=copy,
			adjoint=adjoint, locinfo=locinfo, epsilon=epsilon, niter=niter, wcs=wcs)
	if adjoint:
		if copy and map is not None: map = map.clone() if L.is_torch(map) else map.copy()
	elif copy and alm is not None: alm = alm.clone() if L.is_torch(alm) else alm.copy()
	rdt = _rdtype(map)
	alm, ainfo = prepare_alm(alm=alm, ainfo=ainfo, lmax=lmax, pre=map.shape[:-2], dtype=rdt, convert=adjoint, like=map)
	if deriv: raise NotImplementedError("ducc does not support derivatives for map2alm operations. Can be worked around if necessary.")
	alm_full = _atleast(alm, 3)
	map_full = _atleast(map if L.is_torch(map) else np.asarray(map), 4)
	assert tuple(map_full.shape[:-2]) == tuple(alm_full.shape[:-1]), "map and alm must agree on pre-dimensions"
	if method == "2d":
		lm = min(ainfo.lmax, minfo.ducc_geo.lmax); mm = min(ainfo.mmax, lm)
		pk = _plan_kwargs(map.shape, wcs, minfo, "2d", ainfo, lm, mm)
		kw = {k: pk[k] for k in ("geometry", "phi0", "flip_y", "flip_x", "lmax", "mmax", "mstart", "lstride")}
		for I in np.ndindex(*map_full.shape[:-3]):
			groups = list(spin_helper(spin, alm_full.shape[-2]))
			if _grouped(pk, "adjoint_analysis_2d" if adjoint else "analysis_2d", groups, alm_full[I], map_full[I]): continue
			for s, j1, j2 in groups:
				a, acopied = _comp_block(alm_full[I], j1, j2, 1)
				m, mcopied = _comp_block(map_full[I], j1, j2, 2)
				if adjoint:
					sht.adjoint_analysis_2d(alm=a, map=m, spin=s, **kw)
					if mcopied: map_full[I][j1:j2] = m
				else:
					sht.analysis_2d(map=m, alm=a, spin=s, **kw)
					if acopied: alm_full[I][j1:j2] = a
		return map if adjoint else alm
	# ---- cyl: Jacobi-refined weighted adjoint synthesis
	# ring weights in buffer (north-first) order, as the reference applies them to its flipped buffer (:852-868)
	wring = _ring_weights(map.shape, wcs, minfo) if weights is None else np.asarray(weights, dtype=np.float64)
	wrow = wring[::-1] if minfo.flip[0] else wring                             # caller's row order
	pk = _plan_kwargs(map.shape, wcs, minfo, "cyl", ainfo, ainfo.lmax, ainfo.mmax, weights=np.ascontiguousarray(wrow))
	ny, nx = map.shape[-2:]
	nphi = pk["nphi"]
	# The Jacobi iteration runs on full rings: the reference pads cut rows with zeros (xpad, :864-868), so the
	# residual Y(x) - y also lives on the pixels outside the patch.  Full-width temporaries reproduce that.
	pkf = dict(pk, npix=nphi, ringstart=np.arange(ny, dtype=np.int64)*nphi)
	pkf_now = dict(pkf, weight=None)
	def wmul(m):
		if L.is_torch(m):
			import torch
			return m*torch.as_tensor(np.ascontiguousarray(wrow), device=m.device, dtype=m.dtype)[:, None]
		return m*wrow.astype(m.dtype)[:, None]
	def widen(m):
		if nx == nphi: return m
		out = _zeros_like_kind(m, tuple(m.shape[:-1])+(nphi,), rdt); out[..., :nx] = m; return out
	for I in np.ndindex(*map_full.shape[:-3]):
		for s, j1, j2 in spin_helper(spin, alm_full.shape[-2]):
			a, acopied = _comp_block(alm_full[I], j1, j2, 1)
			m, mcopied = _comp_block(map_full[I], j1, j2, 2)
			if niter == 0 and not adjoint:
				_synth(pk, a, m, s, adjoint=True)
				if acopied: alm_full[I][j1:j2] = a
				continue
			fshape = tuple(m.shape[:-1])+(nphi,)
			def Y(x):
				out = _zeros_like_kind(m, fshape, rdt); _synth(pkf_now, x, out, s); return out
			def YT(y):
				out = _zeros_like_kind(a, a.shape, _ctype_of(rdt)); _synth(pkf_now, out, y, s, adjoint=True); return out
			def YTW(y):
				out = _zeros_like_kind(a, a.shape, _ctype_of(rdt)); _synth(pkf, out, y, s, adjoint=True); return out
			def WY(x): return wmul(Y(x))
			if adjoint:
				x = WY(a)
				for it in range(niter): x -= WY(YT(x)-a)
				map_full[I][j1:j2] = x[..., :nx]
			else:
				y = widen(m)
				x = YTW(y)
				for it in range(niter): x -= YTW(Y(x)-y)
				_assign(a, x)
				if acopied: alm_full[I][j1:j2] = a
	return map if adjoint else alm

def _assign(dst, src):
	if L.is_torch(dst): dst.copy_(src)
	else: dst[...] = src

def map2alm_adjoint(alm, map, lmax=None, spin=[0,2], deriv=False, copy=False, method="auto", ainfo=None,
		verbose=False, nthread=None, niter=0, epsilon=None, pix_tol=1e-6, weights=None, locinfo=None, wcs=None):
	"""pixell/curvedsky.py:304-310"""
	return map2alm(map, alm, lmax=lmax, spin=spin, deriv=deriv, adjoint=True, copy=copy, method=method, ainfo=ainfo,
		verbose=verbose, nthread=nthread, niter=niter, pix_tol=pix_tol, weights=weights, wcs=wcs)

# ------------------------------------------------------------------ alm helpers

def almxfl(alm, lfilter=None, ainfo=None, out=None):
	"""pixell/curvedsky.py:630-651"""
	if not L.is_torch(alm): alm = np.asarray(alm)
	ainfo = alm_info(nalm=alm.shape[-1]) if ainfo is None else ainfo
	if callable(lfilter): lfilter = lfilter(np.arange(ainfo.lmax+1.0))
	return ainfo.lmul(alm, lfilter, out=out)

def alm2cl(alm, alm2=None, ainfo=None, dtype=None):
	"""pixell/curvedsky.py:672-712"""
	if not L.is_torch(alm): alm = np.asarray(alm)
	ainfo = alm_info(nalm=alm.shape[-1]) if ainfo is None else ainfo
	return ainfo.alm2cl(alm, alm2=alm2, dtype=dtype)

def transfer_alm(iainfo, ialm, oainfo, oalm=None, op=None):
	"""pixell/curvedsky.py:744-750"""
	return cmisc.transfer_alm(iainfo, ialm, oainfo, oalm=oalm, op=op)

def filter(imap, lfilter, ainfo=None, lmax=None, wcs=None):
	"""pixell/curvedsky.py:653-669: alm2map(almxfl(map2alm(imap)))"""
	wcs = geometry.wcs_of(imap, wcs)
	alm = almxfl(map2alm(imap, ainfo=ainfo, lmax=lmax, spin=0, wcs=wcs), lfilter=lfilter, ainfo=ainfo)
	omap = _zeros_like_kind(imap, tuple(imap.shape), _rdtype(imap))
	if not L.is_torch(omap): omap = geometry.ndmap(omap, wcs)
	return alm2map(alm, omap, spin=0, ainfo=ainfo, wcs=wcs)

# ------------------------------------------------------------------ random fields

def pad_spectrum(ps, lmax):
	ps = np.asarray(ps)
	ops = np.zeros(ps.shape[:-1]+(lmax+1,), ps.dtype)
	ops[..., :ps.shape[-1]] = ps[..., :ps.shape[-1]]
	return ops

def sym_expand(ps, scheme="diag"):
	"""powspec.sym_expand for the diagonal-first scheme (pixell/powspec.py:5-20)"""
	ps = np.asarray(ps)
	ncomp = int(((1+8*ps.shape[0])**0.5-1)/2)
	out = np.zeros((ncomp, ncomp)+ps.shape[1:], ps.dtype)
	k = 0
	for d in range(ncomp):
		for i in range(ncomp-d):
			out[i, i+d] = out[i+d, i] = ps[k]; k += 1
	return out

def prepare_ps(ps, ainfo=None, lmax=None):
	"""pixell/curvedsky.py:608-618"""
	ps = np.asarray(ps)
	if ainfo is None:
		if lmax is None: lmax = ps.shape[-1]-1
		if lmax > ps.shape[-1]-1: ps = pad_spectrum(ps, lmax)
		ainfo = alm_info(lmax)
	if   ps.ndim == 1: wps = ps[None, None]
	elif ps.ndim == 2: wps = sym_expand(ps, scheme="diag")
	elif ps.ndim == 3: wps = ps
	else: raise ValueError("power spectrum must be [nl], [nspec,nl] or [ncomp,ncomp,nl]")
	return wps, ainfo

def multi_pow_half(ps):
	"""enmap.multi_pow(ps, 0.5) (pixell/enmap.py:2021-2024 -> utils.eigpow :2789-2830): symmetric square
	root per l by eigendecomposition, negative eigenvalues -> 0.  [ncomp,ncomp,nl] host arrays, tiny."""
	A = np.moveaxis(np.asarray(ps, np.float64), -1, 0)
	E, V = np.linalg.eigh(A)
	E = np.where(E < 0, 0, np.abs(E)**0.5)
	return np.moveaxis(np.einsum("...ij,...kj->...ik", V*E[..., None, :], V), 0, -1)

def fill_gauss(arr, bsize=0x10000):
	"""pixell/curvedsky.py:602-606: numpy's legacy global stream, 65536 numbers at a time"""
	rtype = np.zeros([0], arr.dtype).real.dtype
	flat = arr.reshape(-1).view(rtype)
	for i in range(0, flat.size, bsize):
		flat[i:i+bsize] = np.random.standard_normal(min(bsize, flat.size-i))

def rand_alm_white(ainfo, pre=None, alm=None, seed=None, dtype=np.complex128, m_major=True):
	"""pixell/curvedsky.py:620-628.  The random stream is numpy's (host) so that seeds reproduce the
	reference bit for bit; the l-major -> m-major transpose runs on the GPU."""
	if seed is not None: np.random.seed(seed)
	if alm is None:
		alm = np.empty(ainfo.nelem if pre is None else tuple(pre)+(ainfo.nelem,), dtype)
	fill_gauss(alm)
	if m_major: ainfo.transpose_alm(alm, alm)
	return alm

def rand_alm(ps, ainfo=None, lmax=None, seed=None, dtype=np.complex128, m_major=True, return_ainfo=False):
	"""pixell/curvedsky.py:61-77"""
	ps = np.asarray(ps)
	rtype = np.zeros([0], dtype=dtype).real.dtype
	wps, ainfo = prepare_ps(ps, ainfo=ainfo, lmax=lmax)
	alm = rand_alm_white(ainfo, pre=[wps.shape[0]], seed=seed, dtype=dtype, m_major=m_major)
	ps12 = multi_pow_half(wps)
	ainfo.lmul(alm, (ps12/2**0.5).astype(rtype, copy=False), alm)
	alm[:, :ainfo.lmax+1].imag = 0
	alm[:, :ainfo.lmax+1].real *= 2**0.5
	if ps.ndim == 1: alm = alm[0]
	return (alm, ainfo) if return_ainfo else alm

def rand_alm_healpy(ps, lmax=None, seed=None, dtype=np.complex128):
	"""pixell/curvedsky.py:44-59 restated without healpy.  The scalar stream (what the reference's
	golden MM_041121.pkl pins) is healpy.synalm's: re = N(0,1)[nalm], im = N(0,1)[nalm] in m-major order,
	a_l0 = sqrt(C_l) re, a_lm = sqrt(C_l/2)(re + i im).  For several components healpy's stream is not
	pinned by any reference fixture (parity unpinned): we colour white alm with the symmetric square root."""
	ps = np.asarray(ps)
	if lmax is None: lmax = ps.shape[-1]-1
	if ps.ndim == 1:
		if seed is not None: np.random.seed(seed)
		ainfo = alm_info(lmax)
		re = np.random.standard_normal(ainfo.nelem); im = np.random.standard_normal(ainfo.nelem)
		alm = (re + 1j*im).astype(dtype)
		cl = np.zeros(lmax+1); n = min(lmax+1, ps.shape[-1]); cl[:n] = ps[:n]
		alm = ainfo.lmul(alm, np.sqrt(cl/2).astype(alm.real.dtype), alm)
		alm[:lmax+1] = re[:lmax+1]*np.sqrt(cl)
		return alm
	import warnings
	warnings.warn("pixell_b200.rand_alm_healpy: for several components the alm follow pixell's own rand_alm stream (l-major fill, symmetric "
		"square root), not healpy.synalm's (m-major, Cholesky-like mixing): same covariance, different realisation for a given seed", stacklevel=2)
	return rand_alm(ps, lmax=lmax, seed=seed, dtype=dtype)

def rand_map(shape, wcs, ps, lmax=None, dtype=np.float64, seed=None, spin=[0,2], method="auto", verbose=False):
	"""pixell/curvedsky.py:17-36"""
	ps = np.asarray(ps)
	if ps.ndim == 1: ps3 = ps[None, None]
	elif ps.ndim == 2: ps3 = sym_expand(ps)
	else: ps3 = ps
	if not ps3.shape[0] == ps3.shape[1]: raise ShapeError("ps must be [ncomp,ncomp,nl] or [nl]")
	if not (len(shape) == 2 or len(shape) == 3): raise ShapeError("shape must be (ncomp,ny,nx) or (ny,nx)")
	ncomp = 1 if len(shape) == 2 else shape[-3]
	ps3 = ps3[:ncomp, :ncomp]
	ctype = np.result_type(dtype, 0j)
	if lmax is None: lmax = ps3.shape[-1]-1
	alm = rand_alm_healpy(ps3[0, 0] if ncomp == 1 else ps3, lmax=lmax, seed=seed, dtype=ctype)
	alm = np.atleast_2d(alm)
	map = geometry.empty((ncomp,)+tuple(shape[-2:]), wcs, dtype=dtype)
	alm2map(alm, map, spin=spin, method=method, verbose=verbose)
	if len(shape) == 2: map = map[0]
	return map

# ------------------------------------------------------------------ named entry points of the reference's transform layers
# pixell/curvedsky.py:756-873 (per-method helpers) and :900-1086 (raw helpers).  In the reference the "raw" functions work on
# flipped / padded buffers (map2buffer / buffer2map, :1384-1411); here flips and cuts are index arithmetic inside the engine,
# so every layer forwards to alm2map / map2alm with the method fixed and no copy is ever made.

def alm2map_2d(alm, map, ainfo=None, minfo=None, spin=[0,2], deriv=False, copy=False, verbose=False, adjoint=False, nthread=None, pix_tol=1e-6, wcs=None):
	"""pixell/curvedsky.py:756-774"""
	return alm2map(alm, map, spin=spin, deriv=deriv, adjoint=adjoint, copy=copy, method="2d", ainfo=ainfo, verbose=verbose, pix_tol=pix_tol, wcs=wcs)

def alm2map_cyl(alm, map, ainfo=None, minfo=None, spin=[0,2], deriv=False, copy=False, verbose=False, adjoint=False, nthread=None, pix_tol=1e-6, wcs=None):
	"""pixell/curvedsky.py:776-794"""
	return alm2map(alm, map, spin=spin, deriv=deriv, adjoint=adjoint, copy=copy, method="cyl", ainfo=ainfo, verbose=verbose, pix_tol=pix_tol, wcs=wcs)

def map2alm_2d(map, alm=None, ainfo=None, minfo=None, lmax=None, spin=[0,2], deriv=False, copy=False, verbose=False, adjoint=False, nthread=None, pix_tol=1e-6, wcs=None):
	"""pixell/curvedsky.py:822-841"""
	return map2alm(map, alm=alm, lmax=lmax, spin=spin, deriv=deriv, adjoint=adjoint, copy=copy, method="2d", ainfo=ainfo, verbose=verbose, pix_tol=pix_tol, wcs=wcs)

def map2alm_cyl(map, alm=None, ainfo=None, minfo=None, lmax=None, spin=[0,2], weights=None, deriv=False, copy=False, verbose=False, adjoint=False,
		nthread=None, pix_tol=1e-6, niter=0, wcs=None):
	"""pixell/curvedsky.py:843-873"""
	return map2alm(map, alm=alm, lmax=lmax, spin=spin, deriv=deriv, adjoint=adjoint, copy=copy, method="cyl", ainfo=ainfo, verbose=verbose,
		niter=niter, pix_tol=pix_tol, weights=weights, wcs=wcs)

def alm2map_raw_2d(alm, map, ainfo=None, spin=[0,2], deriv=False, copy=False, verbose=False, adjoint=False, nthread=None, wcs=None):
	"""pixell/curvedsky.py:900-924: the map must be a full ducc grid (any orientation: no buffer is needed here)"""
	return alm2map_2d(alm, map, ainfo=ainfo, spin=spin, deriv=deriv, copy=copy, verbose=verbose, adjoint=adjoint, wcs=wcs)

def alm2map_raw_cyl(alm, map, ainfo=None, minfo=None, spin=[0,2], deriv=False, copy=False, verbose=False, adjoint=False, nthread=None, wcs=None):
	"""pixell/curvedsky.py:926-962"""
	return alm2map_cyl(alm, map, ainfo=ainfo, spin=spin, deriv=deriv, copy=copy, verbose=verbose, adjoint=adjoint, wcs=wcs)

def map2alm_raw_2d(map, alm=None, ainfo=None, lmax=None, spin=[0,2], deriv=False, copy=False, verbose=False, adjoint=False, nthread=None, wcs=None):
	"""pixell/curvedsky.py:1018-1046"""
	return map2alm_2d(map, alm=alm, ainfo=ainfo, lmax=lmax, spin=spin, deriv=deriv, copy=copy, verbose=verbose, adjoint=adjoint, wcs=wcs)

def map2alm_raw_cyl(map, alm=None, ainfo=None, lmax=None, spin=[0,2], weights=None, deriv=False, copy=False, verbose=False, adjoint=False, niter=0,
		nthread=None, wcs=None):
	"""pixell/curvedsky.py:1050-1086"""
	return map2alm_cyl(map, alm=alm, ainfo=ainfo, lmax=lmax, spin=spin, weights=weights, deriv=deriv, copy=copy, verbose=verbose, adjoint=adjoint,
		niter=niter, wcs=wcs)

def get_ducc_maxlmax(name, ny):
	"""pixell/curvedsky.py:1349-1353"""
	return sht.maxlmax(name, ny)

def dangerous_dtype(dtype):
	"""pixell/curvedsky.py:1448-1449"""
	return np.dtype(dtype).byteorder not in ("=", "|", "<" if np.little_endian else ">")

# buffers (pixell/curvedsky.py:1236-1250, 1384-1411): kept for callers that use them directly; the transforms above do not
def flip2slice(flips):
	res = (Ellipsis,)
	for flip in flips: res = res + (slice(None, None, 1-2*int(flip)),)
	return res
def flip_array(arr, flips): return arr[flip2slice(flips)]
def flip_geometry(shape, wcs, flips):
	return tuple(shape), _flipped(shape, wcs, [bool(f) for f in flips])
def pad_geometry(shape, wcs, pad):
	pad = np.asarray(pad)
	w, h = int(pad[0, 0]+shape[-2]+pad[1, 0]), int(pad[0, 1]+shape[-1]+pad[1, 1])
	ww = wcs.wcs
	owcs = geometry.CarWCS(ww.crval, ww.cdelt, np.array(ww.crpix, float)+pad[0, ::-1], getattr(ww, "ctype", ("RA---CAR", "DEC--CAR")))
	return tuple(shape[:-2])+(w, h), owcs
def map2buffer(map, flip, pad, obuf=False, wcs=None):
	"""north-first, ra-increasing, zero-padded copy of a host map (pixell/curvedsky.py:1384-1404)"""
	pad = np.asarray(pad)
	wcs = geometry.wcs_of(map, wcs)
	shape, w = pad_geometry(*flip_geometry(map.shape, wcs, flip), pad)
	buf = geometry.zeros(shape, w, np.asarray(map).dtype.newbyteorder("="))
	if not obuf: buf[..., pad[0, 0]:buf.shape[-2]-pad[1, 0], pad[0, 1]:buf.shape[-1]-pad[1, 1]] = flip_array(np.asarray(map), flip)
	return buf
def buffer2map(map, flip, pad):
	"""pixell/curvedsky.py:1406-1411"""
	pad = np.array(pad)
	map = map[..., pad[0, 0]:map.shape[-2]-pad[1, 0], pad[0, 1]:map.shape[-1]-pad[1, 1]]
	return flip_array(map, flip)

def prepare_raw(alm, map, ainfo=None, lmax=None, deriv=False, verbose=False, nthread=None, pixdims=2, convert_alm=False):
	"""pixell/curvedsky.py:1429-1446: validated, dimension-padded views of alm and map"""
	alm, ainfo = prepare_alm(alm, ainfo, lmax=lmax, pre=map.shape[:-pixdims], dtype=_rdtype(map), convert=convert_alm)
	alm_full = _atleast(alm, 2 if deriv else 3)
	map_full = _atleast(map if L.is_torch(map) else np.asarray(map), pixdims+2)
	if deriv:
		assert map_full.ndim >= pixdims+1 and map_full.shape[-pixdims-1] == 2, "map must have shape [...,2,%s] when deriv is True" % ("nloc" if pixdims == 1 else "ny,nx")
		assert tuple(map_full.shape[:-1-pixdims]) == tuple(alm_full.shape[:-1]), "map and alm must agree on pre-dimensions"
	else:
		assert tuple(map_full.shape[:-pixdims]) == tuple(alm_full.shape[:-1]), "map and alm must agree on pre-dimensions"
	return alm_full, map_full, ainfo, 0

# ------------------------------------------------------------------ iterative inverses (pixell/curvedsky.py:1122-1168)

def jacobi_inverse(forward, approx_backward, y, niter=0):
	"""x from y = forward(x) by Jacobi iteration with the approximate inverse approx_backward (pixell/curvedsky.py:1122-1136)"""
	x = approx_backward(y)
	for i in range(niter):
		x -= approx_backward(forward(x)-y)
	return x

class _Minres:
	"""MINRES for a symmetric operator A given as a function on flat real vectors (Paige & Saunders 1975; what
	pixell/utils.py's Minres class provides to minres_inverse)"""
	def __init__(self, A, b):
		self.A = A; self.x = np.zeros_like(b)
		self.r = b.copy(); self.beta = float(np.sqrt(self.r @ self.r)); self.b2 = max(self.beta**2, 1e-300)
		self.v_old = np.zeros_like(b); self.v = self.r/self.beta if self.beta > 0 else self.r
		self.w_old = np.zeros_like(b); self.w = np.zeros_like(b)
		self.c_old = self.c = 1.0; self.s_old = self.s = 0.0; self.eta = self.beta; self.i = 0
		self.abserr = self.beta**2/self.b2
	def step(self):
		Av = self.A(self.v)
		alpha = float(self.v @ Av)
		v_new = Av - alpha*self.v - self.beta*self.v_old
		beta_new = float(np.sqrt(v_new @ v_new))
		if beta_new > 0: v_new = v_new/beta_new
		# previous rotations
		delta = self.c*alpha - self.c_old*self.s*self.beta
		rho2 = self.s*alpha + self.c_old*self.c*self.beta
		rho3 = self.s_old*self.beta
		rho1 = np.sqrt(delta*delta + beta_new*beta_new)
		c_new, s_new = (delta/rho1, beta_new/rho1) if rho1 > 0 else (1.0, 0.0)
		w_new = (self.v - rho3*self.w_old - rho2*self.w)/rho1 if rho1 > 0 else np.zeros_like(self.v)
		self.x = self.x + c_new*self.eta*w_new
		self.eta = -s_new*self.eta
		self.v_old, self.v, self.beta = self.v, v_new, beta_new
		self.w_old, self.w = self.w, w_new
		self.c_old, self.c, self.s_old, self.s = self.c, c_new, self.s, s_new
		self.i += 1
		self.abserr = self.eta**2/self.b2

def minres_inverse(forward, approx_backward, y, epsilon=1e-6, maxiter=100, zip=None, unzip=None, verbose=False):
	"""the maximum-likelihood x of y = forward(x) by MINRES on approx_backward(forward(x)) = approx_backward(y)
	(pixell/curvedsky.py:1138-1168)"""
	rhs = approx_backward(y)
	rhs = np.asarray(rhs.cpu().numpy() if L.is_torch(rhs) else rhs)
	rtype = np.zeros(1, rhs.dtype).real.dtype
	if zip is None:
		def zip(a): return np.ascontiguousarray(a).view(rtype).reshape(-1)
	if unzip is None:
		def unzip(x): return x.view(rhs.dtype).reshape(rhs.shape)
	def A(x): return zip(approx_backward(forward(unzip(x)))).astype(rtype, copy=True)
	solver = _Minres(A, zip(rhs).astype(rtype, copy=True))
	while solver.abserr**0.5 > epsilon and solver.i < maxiter:
		solver.step()
		if verbose: print("Minres %4d %15.7e" % (solver.i, solver.abserr**0.5))
	return unzip(solver.x)

# ------------------------------------------------------------------ real packing of alm (pixell/curvedsky.py:1451-1473)

def alm_complex2real(alm, ainfo=None):
	alm = np.asarray(alm)
	dtype = np.zeros(1, alm.dtype).real.dtype
	if ainfo is None: ainfo = alm_info(nalm=alm.shape[-1])
	i = int(ainfo.mstart[1]+1)
	return np.concatenate([alm[..., :i].real, 2**0.5*np.ascontiguousarray(alm[..., i:]).view(dtype)], -1)

def alm_real2complex(ralm, ainfo=None):
	ralm = np.asarray(ralm)
	ctype = np.result_type(ralm.dtype, 0j)
	if ainfo is None:
		lmax = int(np.rint((ralm.shape[-1]-1)**0.5))-1
		ainfo = alm_info(lmax=lmax)
	i = int(ainfo.mstart[1]+1)
	oalm = np.zeros(ralm.shape[:-1]+(ainfo.nelem,), ctype)
	oalm[..., :i] = ralm[..., :i]
	oalm[..., i:] = np.ascontiguousarray(ralm[..., i:]).view(ctype)/2**0.5
	return oalm

# ------------------------------------------------------------------ profiles on the sky (pixell/curvedsky.py:558-585)

def prof2alm(profile, dir=[0, np.pi/2], spin=0, geometry="CC", nthread=None, norot=False):
	"""alm of a 1-d equispaced profile[..., n] (colatitude 0..pi on the named ring grid) oriented along dir = [ra, dec]:
	exact analysis of the m = 0 map, expansion to mmax = lmax, rotation of the pole to the target direction"""
	profile = np.asarray(profile)
	n = profile.shape[-1]
	lmax = get_ducc_maxlmax(geometry, n)
	iainfo = alm_info(lmax=lmax, mmax=0)
	oainfo = alm_info(lmax=lmax, mmax=lmax if not norot else 0)
	ctype = np.result_type(profile.dtype, 0j)
	oalm = np.zeros(profile.shape[:-1]+(oainfo.nelem,), ctype)
	spins = np.array(spin).reshape(-1)
	for I in np.ndindex(*profile.shape[:-1]):
		s = int(spins[(I[-1] if len(I) else 0) % len(spins)]) if len(spins) > 1 else int(spins[0])
		if s != 0: raise NotImplementedError("prof2alm: only scalar profiles (spin 0) are provided")
		prof = np.ascontiguousarray(profile[I], dtype=np.float64)[None, :, None]
		alm = np.asarray(sht.analysis_2d(map=prof, spin=0, lmax=lmax, mmax=0, geometry=geometry))
		if not norot:
			alm = transfer_alm(iainfo, alm, oainfo)
			alm = rotate_alm(alm, 0, np.pi/2-dir[1], dir[0])
		oalm[I] = alm[0] if alm.ndim == 2 else alm
	return oalm
