"""pixell_b200.reproject -- harmonic reprojection between CAR maps and HEALPix maps, optionally with a coordinate
rotation (reference pixell/reproject.py:118-247 map2healpix, :249-361 healpix2map, method="harm"): map2alm(_healpix) ->
rotate_alm -> alm2map(_healpix), all on the engine.  The pixel-space method="spline" (healpy interpolation) is not provided."""
import numpy as np
from . import curvedsky, enmap, geometry
from .geometry import DEG

def rot2euler(rot):
	"""zyz Euler angles of a rotation given as those angles or as "isys,osys" with systems cel / equ / gal
	(pixell/reproject.py:363-384)"""
	from scipy.spatial.transform import Rotation
	gal2cel = np.array([57.06793215, 62.87115487, -167.14056929])*DEG
	if isinstance(rot, str):
		try: isys, osys = rot.split(",")
		except ValueError: raise ValueError("Rotation string must be of form 'isys,osys', but got '%s'" % str(rot))
		R = Rotation.identity()
		if isys in ["cel", "equ"]: pass
		elif isys == "gal": R *= Rotation.from_euler("zyz", gal2cel)
		else: raise ValueError("Unrecognized system '%s'" % isys)
		if osys in ["cel", "equ"]: pass
		elif osys == "gal": R *= Rotation.from_euler("zyz", gal2cel).inv()
		else: raise ValueError("Unrecognized system '%s'" % osys)
		return R.as_euler("zyz")
	return np.asarray(rot, dtype=np.float64)

def restrict_nside(nside, mode="mul32", round="ceil"):
	"""pixell/reproject.py:388-418"""
	if isinstance(round, str): round = {"floor": np.floor, "round": np.round, "ceil": np.ceil}[round]
	if mode == "any": nside = round(nside)
	elif mode == "mul32":
		if 12*nside**2 > 1024: nside = round(nside/32)*32
	elif mode == "pow2": nside = 2**round(np.log2(nside))
	else: raise ValueError("Unrecognized nside mode '%s'" % str(mode))
	return max(1, int(nside))

def _only_harm(method):
	if method not in ["harm", "harmonic"]:
		if method == "spline": raise NotImplementedError("pixell_b200.reproject: method='spline' (healpy pixel interpolation) is not provided")
		raise ValueError("Map reprojection method '%s' not recognized" % str(method))

def map2healpix(imap, nside=None, lmax=None, out=None, rot=None, spin=[0,2], method="harm", order=1, extensive=False,
		bsize=100000, nside_mode="pow2", boundary="constant", verbose=False, niter=0, wcs=None):
	"""CAR map [..., ny, nx] -> HEALPix map [..., npix] (RING), optionally rotated (pixell/reproject.py:118-247)."""
	_only_harm(method)
	wcs = geometry.wcs_of(imap, wcs)
	ires = np.mean(np.abs(wcs.wcs.cdelt))*DEG
	if out is None:
		if nside is None: nside = restrict_nside(((4*np.pi/ires**2)/12)**0.5, nside_mode)
		out = np.zeros(imap.shape[:-2]+(12*nside**2,), imap.dtype)
	npix = out.shape[-1]
	if lmax is None: lmax = int(np.pi/ires)
	if extensive: imap = geometry.ndmap(np.asarray(imap)*((4*np.pi/npix)/geometry.pixsize_rows(imap.shape, wcs)[:, None]), wcs)
	alm = curvedsky.map2alm(imap, lmax=lmax, spin=spin, niter=niter, wcs=wcs)
	if rot is not None: curvedsky.rotate_alm(alm, *rot2euler(rot), inplace=True)
	curvedsky.alm2map_healpix(alm, out, spin=spin)
	return out

def healpix2map(iheal, shape=None, wcs=None, lmax=None, out=None, rot=None, spin=[0,2], method="harm", order=1, extensive=False,
		bsize=100000, verbose=False, niter=0):
	"""HEALPix map [..., npix] (RING) -> CAR map [..., ny, nx], optionally rotated (pixell/reproject.py:249-361)."""
	_only_harm(method)
	iheal = np.asarray(iheal)
	npix = iheal.shape[-1]
	nside = curvedsky.npix2nside(npix)
	if out is None: out = geometry.zeros(iheal.shape[:-1]+tuple(shape[-2:]), wcs, dtype=iheal.dtype)
	else: wcs = geometry.wcs_of(out, wcs)
	if lmax is None: lmax = 3*nside
	alm = curvedsky.map2alm_healpix(iheal, lmax=lmax, spin=spin, niter=niter)
	if rot is not None: curvedsky.rotate_alm(alm, *rot2euler(rot), inplace=True)
	curvedsky.alm2map(alm, out, spin=spin, wcs=wcs)
	if extensive: out *= geometry.pixsize_rows(out.shape, wcs)[:, None]/(4*np.pi/npix)
	return out
