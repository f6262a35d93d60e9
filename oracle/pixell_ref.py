"""oracle/pixell_ref.py -- TEST INFRASTRUCTURE ONLY (CPU checker; never imported by the product).

Compact CPU restatement of pixell's curvedsky host logic on top of sht_oracle (our CPU stand-in
for ducc0): geometry analysis, flip/pad buffers, spin loop, Jacobi iterations.  It deliberately
mirrors the reference's *copying* implementation (map2buffer/buffer2map) so that the product's
zero-copy index arithmetic is checked against the straightforward version.

Follows reference pixell/curvedsky.py:
  analyse_geometry :1252-1306   get_ducc_geo :1308-1347   get_ducc_maxlmax :1349-1353
  alm2map_2d/_cyl :756-794      map2alm_2d/_cyl :822-873   *_raw_2d/_raw_cyl :900-962, 1018-1086
  jacobi_inverse :1122-1136     get_ring_info :1170-1190   map2buffer/buffer2map :1384-1411
  spin_helper: pixell/enmap.py:3378-3388
Geometry is a plain CAR description (shape, crval, cdelt, crpix in degrees; FITS 1-based crpix),
because astropy (and therefore pixell.enmap / wcsutils) is not available.
"""
import numpy as np
from . import sht_oracle as so
from .alm_oracle import AlmInfo

DEG = np.pi/180

class Geo:
	def __init__(self, shape, crval, cdelt, crpix):
		self.shape = tuple(shape[-2:]); self.crval = np.array(crval, float)
		self.cdelt = np.array(cdelt, float); self.crpix = np.array(crpix, float)
		assert abs(self.crval[1]) < 1e-12, "CAR with crval_dec != 0 is not a plain cylindrical grid"
	def dec(self, y): return (self.crval[1] + (np.asarray(y)+1-self.crpix[1])*self.cdelt[1])*DEG
	def ra(self, x):  return (self.crval[0] + (np.asarray(x)+1-self.crpix[0])*self.cdelt[0])*DEG
	def ypix(self, dec_deg): return self.crpix[1]-1 + (dec_deg-self.crval[1])/self.cdelt[1]
	def flipped(self, flip):
		crpix, cdelt = self.crpix.copy(), self.cdelt.copy()
		if flip[0]: crpix[1] = self.shape[0]+1-crpix[1]; cdelt[1] = -cdelt[1]
		if flip[1]: crpix[0] = self.shape[1]+1-crpix[0]; cdelt[0] = -cdelt[0]
		return Geo(self.shape, self.crval, cdelt, crpix)

def fullsky_geo(shape=None, res=None, variant="fejer1"):
	"""enmap.fullsky_geometry (pixell/enmap.py:1713-1740), res in radians."""
	yo = {"cc": 1, "fejer1": 0}[variant.lower()]
	if shape is None:
		res = np.zeros(2)+res
		shape = tuple(np.rint(np.array([np.pi, 2*np.pi])/res + (yo, 0)).astype(int))
	ny, nx = shape
	return Geo((ny, nx), [360./nx/2, 0], [-360./nx, 180./(ny-yo)], [nx//2+0.5, (ny+1)/2])

def hasoff(val, off, tol): return abs((val-off+0.5) % 1 - 0.5) < tol

def maxlmax(name, ny): return so.maxlmax(name, ny)

def ducc_geo(geo, tol=1e-6):
	"""get_ducc_geo on an already-flipped geometry."""
	nx = 360/geo.cdelt[0]
	if not hasoff(nx, 0, tol): return None
	y1, y2 = geo.ypix(90.0), geo.ypix(-90.0)
	Ny = geo.shape[0]
	near = lambda a, b: abs(a-b) < tol
	if hasoff(y1, 0, tol) and hasoff(y2, 0, tol):
		if   near(y1, -1) and near(y2, Ny): name, o1, o2 = "F2", 1, 1
		elif near(y1, 0) and near(y2, Ny):  name, o1, o2 = "DH", 1, 0
		else: name, o1, o2 = "CC", 0, 0
	elif hasoff(y1, 0.5, tol) and hasoff(y2, 0.5, tol): name, o1, o2 = "F1", 0.5, 0.5
	elif hasoff(y1, 0.5, tol) and hasoff(y2, 0.0, tol): name, o1, o2 = "MW", 0.5, 0.0
	elif hasoff(y1, 0.0, tol) and hasoff(y2, 0.5, tol): name, o1, o2 = "MWflip", 0.0, 0.5
	else: return None
	ny = int(np.rint(y2-y1+1-o1-o2)); yoff = int(np.rint(-y1-o1))
	return dict(name=name, nx=int(np.rint(nx)), ny=ny, yoff=yoff, lmax=maxlmax(name, ny))

def analyse_geometry(geo, tol=1e-6):
	if not hasoff(360/abs(geo.cdelt[0]), 0, tol):
		return dict(case="general")
	flip = [geo.cdelt[1] > 0, geo.cdelt[0] < 0]
	w = geo.flipped(flip)
	phi0 = w.ra(0)
	dg = ducc_geo(w, tol)
	if dg is not None and geo.shape[0] == dg["ny"] and geo.shape[1] == dg["nx"] and abs(dg["yoff"]) < tol:
		return dict(case="2d", flip=flip, ducc_geo=dg, ypad=(0,0), xpad=(0,0), phi0=phi0, wgeo=w)
	ypad = (dg["yoff"], dg["ny"]-dg["yoff"]-geo.shape[0]) if dg is not None else (0,0)
	nx = int(np.rint(360/w.cdelt[0]))
	if geo.shape[1] == nx:
		return dict(case="cyl", flip=flip, ducc_geo=dg, ypad=ypad, xpad=(0,0), phi0=phi0, wgeo=w)
	return dict(case="partial", flip=flip, ducc_geo=dg, ypad=ypad, xpad=(0, nx-geo.shape[1]), phi0=phi0, wgeo=w)

def get_method(minfo):
	if minfo["case"] == "general": return "general"
	return "2d" if minfo["case"] == "2d" else "cyl"

def spin_helper(spin, n):
	spin = np.array(spin).reshape(-1); scomp = 1+(spin != 0)
	ci, i1 = 0, 0
	while True:
		i2 = min(i1+scomp[ci], n)
		if i2-i1 != scomp[ci]: raise IndexError("Unpaired component in spin transform")
		yield int(spin[ci]), i1, i2
		if i2 == n: break
		i1 = i2; ci = (ci+1) % len(spin)

def _flip(a, flip):
	if flip[0]: a = a[..., ::-1, :]
	if flip[1]: a = a[..., :, ::-1]
	return a

def _map2buffer(map, flip, ypad, xpad, obuf=False):
	ny, nx = map.shape[-2:]
	buf = np.zeros(map.shape[:-2]+(ypad[0]+ny+ypad[1], xpad[0]+nx+xpad[1]), map.dtype)
	if not obuf: buf[..., ypad[0]:ypad[0]+ny, xpad[0]:xpad[0]+nx] = _flip(map, flip)
	return buf

def _buffer2map(buf, flip, ypad, xpad):
	b = buf[..., ypad[0]:buf.shape[-2]-ypad[1], xpad[0]:buf.shape[-1]-xpad[1]]
	return _flip(b, flip)

def _ring_info(minfo, ny):
	"""get_ring_info for the flipped (north-first, phi increasing), x-padded buffer."""
	w = minfo["wgeo"]
	theta = np.pi/2 - w.dec(np.arange(ny))
	nx = int(np.rint(360/w.cdelt[0]))
	nphi = np.full(ny, nx, np.int64)
	phi0 = np.full(ny, w.ra(0))
	return dict(theta=theta, nphi=nphi, phi0=phi0, ringstart=np.arange(ny, dtype=np.int64)*nx)

def _prep_alm(alm, ainfo, lmax, pre, rdtype):
	ctype = np.result_type(rdtype, 0j)
	if alm is None:
		if ainfo is None:
			if lmax is None: raise ValueError("need alm, ainfo or lmax")
			ainfo = AlmInfo(lmax)
		alm = np.zeros(pre+(ainfo.nelem,), ctype)
	if ainfo is None: ainfo = AlmInfo(nalm=alm.shape[-1])
	return alm, ainfo

def alm2map(alm, map, geo, spin=[0,2], deriv=False, adjoint=False, method="auto", ainfo=None):
	"""curvedsky.alm2map for methods 2d / cyl.  map[...,ncomp,ny,nx] is overwritten (or alm if adjoint)."""
	minfo = analyse_geometry(geo)
	if method == "auto": method = get_method(minfo)
	if method not in ("2d", "cyl"): raise NotImplementedError(method)
	map3 = map.reshape((-1,)+map.shape[-2:]) if map.ndim <= 3 else map
	if adjoint: alm, ainfo = _prep_alm(alm, ainfo, None, map.shape[:-2], map.dtype)
	elif ainfo is None: ainfo = AlmInfo(nalm=alm.shape[-1])
	almN = alm.reshape((-1,)+alm.shape[-1:]) if alm.ndim <= 2 else alm
	assert map3.ndim == 3 and almN.ndim == 2, "oracle supports [ncomp,...] inputs only"
	ypad = minfo["ypad"] if method == "2d" else (0,0)
	buf = _map2buffer(map3, minfo["flip"], ypad, minfo["xpad"], obuf=not adjoint).astype(np.float64)
	kw = dict(lmax=ainfo.lmax, mmax=ainfo.mmax, mstart=ainfo.mstart)
	if method == "2d":
		kw.update(geometry=minfo["ducc_geo"]["name"], phi0=minfo["phi0"], ntheta=buf.shape[-2], nphi=buf.shape[-1])
		syn, asyn = so.synthesis_2d, so.adjoint_synthesis_2d
		view = lambda b: b
	else:
		ri = _ring_info(minfo, buf.shape[-2])
		kw.update(theta=ri["theta"], nphi=ri["nphi"], phi0=ri["phi0"], ringstart=ri["ringstart"])
		syn, asyn = so.synthesis, so.adjoint_synthesis
		view = lambda b: b.reshape(b.shape[0], -1)
	if deriv:
		if adjoint:
			b2 = buf.copy(); b2[0] *= -1
			almN[:] = asyn(map=view(b2), spin=1, mode="DERIV1", **kw)
		else:
			view(buf)[:] = syn(alm=almN[:1].astype(np.complex128), spin=1, mode="DERIV1", **kw)
			buf[0] *= -1
	else:
		for s, i1, i2 in spin_helper(spin, almN.shape[0]):
			if adjoint: almN[i1:i2] = asyn(map=view(buf[i1:i2]), spin=s, **kw)
			else: view(buf[i1:i2])[:] = syn(alm=almN[i1:i2].astype(np.complex128), spin=s, **kw)
	if adjoint: return alm
	map3[:] = _buffer2map(buf, minfo["flip"], ypad, minfo["xpad"])
	return map

def quad_weights(geo):
	"""curvedsky.quad_weights / the weights block of map2alm_cyl (:852-861)."""
	minfo = analyse_geometry(geo)
	dg = minfo["ducc_geo"]
	ny = geo.shape[0]
	if dg is not None:
		w = so.get_gridweights(dg["name"], ny+sum(minfo["ypad"]))
		w = w[minfo["ypad"][0]:len(w)-minfo["ypad"][1]]/dg["nx"]
		return w          # north-first order (buffer order)
	# pixel area per row: |cdelt_ra| * (sin(dec+h)-sin(dec-h)) (enmap.pixsizemap separable CAR)
	wg = minfo["wgeo"]
	dec = wg.dec(np.arange(ny)); h = abs(wg.cdelt[1])*DEG/2
	return abs(wg.cdelt[0])*DEG*(np.sin(dec+h)-np.sin(dec-h))

def map2alm(map, geo, alm=None, lmax=None, spin=[0,2], adjoint=False, method="auto", ainfo=None, niter=0, weights=None):
	"""curvedsky.map2alm for methods 2d / cyl."""
	minfo = analyse_geometry(geo)
	if method == "auto": method = get_method(minfo)
	if method not in ("2d", "cyl"): raise NotImplementedError(method)
	map3 = map.reshape((-1,)+map.shape[-2:]) if map.ndim <= 3 else map
	alm, ainfo = _prep_alm(alm, ainfo, lmax, map.shape[:-2], map.dtype)
	almN = alm.reshape((-1,)+alm.shape[-1:]) if alm.ndim <= 2 else alm
	ypad = minfo["ypad"] if method == "2d" else (0,0)
	buf = _map2buffer(map3, minfo["flip"], ypad, minfo["xpad"], obuf=adjoint).astype(np.float64)
	if method == "2d":
		lm = min(ainfo.lmax, minfo["ducc_geo"]["lmax"]); mm = min(ainfo.mmax, lm)
		kw = dict(lmax=lm, mmax=mm, mstart=ainfo.mstart[:mm+1], geometry=minfo["ducc_geo"]["name"], phi0=minfo["phi0"])
		for s, i1, i2 in spin_helper(spin, almN.shape[0]):
			if adjoint: buf[i1:i2] = so.adjoint_analysis_2d(alm=almN[i1:i2].astype(np.complex128), spin=s, ntheta=buf.shape[-2], nphi=buf.shape[-1], **kw)
			else:
				res = np.zeros((i2-i1, almN.shape[-1]), np.complex128); res[:] = almN[i1:i2]
				so.analysis_2d(map=buf[i1:i2], spin=s, alm=res, **kw)
				almN[i1:i2] = res
	else:
		if weights is None: weights = quad_weights(geo)
		ri = _ring_info(minfo, buf.shape[-2])
		kw = dict(theta=ri["theta"], nphi=ri["nphi"], phi0=ri["phi0"], ringstart=ri["ringstart"],
			lmax=ainfo.lmax, mmax=ainfo.mmax, mstart=ainfo.mstart)
		shp = buf.shape[-2:]
		for s, i1, i2 in spin_helper(spin, almN.shape[0]):
			Y   = lambda a: so.synthesis(alm=a, spin=s, **kw)
			YT  = lambda m: so.adjoint_synthesis(map=m, spin=s, **kw)
			wm  = lambda m: (m.reshape((-1,)+shp)*weights[:,None]).reshape(m.shape)
			YTW = lambda m: YT(wm(m)); WY = lambda a: wm(Y(a))
			if adjoint:
				x = WY(almN[i1:i2].astype(np.complex128))
				y = almN[i1:i2].astype(np.complex128)
				for it in range(niter): x -= WY(YT(x)-y)
				buf[i1:i2] = x.reshape((-1,)+shp)
			else:
				y = buf[i1:i2].reshape(i2-i1, -1)
				x = YTW(y)
				for it in range(niter): x -= YTW(Y(x)-y)
				almN[i1:i2] = x
	if adjoint:
		map3[:] = _buffer2map(buf, minfo["flip"], ypad, minfo["xpad"])
		return map
	return alm
