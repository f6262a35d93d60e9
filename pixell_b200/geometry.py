"""pixell_b200.geometry -- the little of pixell's ndmap/WCS data model the SHT hot path reads.

pixell's enmap.ndmap is a numpy array plus an astropy WCS (pixell/enmap.py:33-163).  The hot path
only ever reads four numbers per axis from it (cdelt, crval, crpix, ctype; pixell/curvedsky.py:
1252-1347), so this module provides a duck-typed stand-in that works without astropy: anything
with `.wcs.cdelt`, `.wcs.crval`, `.wcs.crpix` (degrees, FITS 1-based crpix, axis order ra, dec)
is accepted wherever a `wcs` is expected -- a real astropy.wcs.WCS included.
"""
import numpy as np

DEG = np.pi/180

class _Inner:
	def __init__(self, crval, cdelt, crpix, ctype):
		self.crval = np.array(crval, dtype=np.float64); self.cdelt = np.array(cdelt, dtype=np.float64)
		self.crpix = np.array(crpix, dtype=np.float64); self.ctype = list(ctype)

class CarWCS:
	"""Plate-carree WCS description: mirrors the attributes of astropy.wcs.WCS that pixell uses."""
	def __init__(self, crval, cdelt, crpix, ctype=("RA---CAR", "DEC--CAR")):
		self.wcs = _Inner(crval, cdelt, crpix, ctype)
	def deepcopy(self): return CarWCS(self.wcs.crval, self.wcs.cdelt, self.wcs.crpix, self.wcs.ctype)
	def __repr__(self):
		w = self.wcs
		return "car:{cdelt:[%.4g,%.4g],crval:[%.4g,%.4g],crpix:[%.2f,%.2f]}" % (*w.cdelt, *w.crval, *w.crpix)
	def __eq__(self, other):
		try: return all(np.allclose(getattr(self.wcs, k), getattr(other.wcs, k)) for k in ("crval", "cdelt", "crpix"))
		except AttributeError: return False

class ndmap(np.ndarray):
	"""numpy array + wcs, like pixell.enmap.ndmap (pixell/enmap.py:33-60); slicing that changes the
	pixel grid is not tracked (use the arrays' leading axes only)."""
	def __new__(cls, arr, wcs):
		obj = np.asarray(arr).view(cls); obj.wcs = wcs
		return obj
	def __array_finalize__(self, obj):
		if obj is None: return
		self.wcs = getattr(obj, "wcs", None)

def zeros(shape, wcs, dtype=np.float64): return ndmap(np.zeros(shape, dtype), wcs)
def empty(shape, wcs, dtype=np.float64): return ndmap(np.empty(shape, dtype), wcs)

def fullsky_geometry(res=None, shape=None, dims=(), variant="fejer1"):
	"""enmap.fullsky_geometry for proj="car" (pixell/enmap.py:1713-1740): res in radians.
	variant "fejer1" has pixel centres half a pixel from the poles, "cc" has pixels on the poles."""
	yo = {"cc": 1, "fejer1": 0}[variant.lower()]
	if shape is None:
		res = np.zeros(2)+res
		shape = tuple(np.rint(np.array([np.pi, 2*np.pi])/res + (yo, 0)).astype(int))
	ny, nx = (int(v) for v in shape[-2:])
	wcs = CarWCS(crval=[360./nx/2, 0], cdelt=[-360./nx, 180./(ny-yo)], crpix=[nx//2+0.5, (ny+1)/2])
	return tuple(dims)+(ny, nx), wcs

def band_geometry(dec_cut, res, dims=(), variant="fejer1"):
	"""A full-width declination band cut from the full-sky geometry (enmap.band_geometry,
	pixell/enmap.py:1742-1777, CAR only).  dec_cut in radians: scalar (symmetric) or (dec1, dec2)."""
	dec_cut = np.atleast_1d(dec_cut)
	if dec_cut.size == 1: dec_cut = np.array([-dec_cut[0], dec_cut[0]])
	shape, wcs = fullsky_geometry(res=res, variant=variant)
	ny = shape[-2]
	y = (np.sort(dec_cut)/DEG - wcs.wcs.crval[1])/wcs.wcs.cdelt[1] + wcs.wcs.crpix[1] - 1
	y1, y2 = int(max(0, np.floor(y[0]+0.5))), int(min(ny, np.floor(y[1]+0.5)+1))
	return slice_geometry(shape, wcs, y1, y2, dims=dims)

def slice_geometry(shape, wcs, y1, y2, x1=0, x2=None, dims=()):
	"""Geometry of map[..., y1:y2, x1:x2] (positive unit steps)."""
	ny, nx = shape[-2:]
	if x2 is None: x2 = nx
	w = wcs.wcs
	owcs = CarWCS(w.crval, w.cdelt, [w.crpix[0]-x1, w.crpix[1]-y1], getattr(w, "ctype", ("RA---CAR", "DEC--CAR")))
	return tuple(dims)+(y2-y1, x2-x1), owcs

def wcs_of(map, wcs=None):
	if wcs is not None: return wcs
	w = getattr(map, "wcs", None)
	if w is None or not hasattr(w, "wcs"): raise ValueError("map has no wcs: pass an ndmap or wcs=")
	return w

# ---- pixel <-> sky for separable cylindrical maps (what enmap.pix2sky gives for CAR)
def dec_of(wcs, y): w = wcs.wcs; return (w.crval[1] + (np.asarray(y, dtype=np.float64)+1-w.crpix[1])*w.cdelt[1])*DEG
def ra_of(wcs, x):  w = wcs.wcs; return (w.crval[0] + (np.asarray(x, dtype=np.float64)+1-w.crpix[0])*w.cdelt[0])*DEG
def ypix_of(wcs, dec_deg): w = wcs.wcs; return w.crpix[1]-1 + (dec_deg-w.crval[1])/w.cdelt[1]

def pixsize_rows(shape, wcs):
	"""Per-row pixel area of a separable CAR map (one column of enmap.pixsizemap): |dra| (sin(dec+h)-sin(dec-h))."""
	dec = dec_of(wcs, np.arange(shape[-2])); h = abs(wcs.wcs.cdelt[1])*DEG/2
	return abs(wcs.wcs.cdelt[0])*DEG*(np.sin(dec+h)-np.sin(dec-h))
