"""Test scaffolding (last row): import the reference's UNMODIFIED pixell package (staged byte for byte under
baseline/_ref by scripts/stage_reference.py, or read from /root/reference when present) with

  ducc0          -> a stand-in exposing ducc0.sht.experimental.* from `backend` (pixell_b200.sht on the GPU box, the CPU
                    oracle in the CPU tests), i.e. exactly the call sites pixell/curvedsky.py:328-1115 bind
  pixell.cmisc   -> `cmisc_backend` (pixell_b200.cmisc on the GPU box; oracle.alm_oracle wrapped in the CPU tests)
  astropy.wcs    -> a small WCS class for the plate-carree, "plain" and gnomonic projections (astropy is not installable here)
  astropy.io.fits, healpy, matplotlib, h5py and the reference's compiled extension modules -> inert placeholders

and run test methods of the reference's own tests/test_pixell.py against it.  Nothing of the reference is edited."""
import importlib, importlib.abc, importlib.machinery, os, sys, types, unittest, warnings
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

def reference_root():
	for d in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
		if os.path.isfile(os.path.join(d, "pixell", "curvedsky.py")) and os.path.isfile(os.path.join(d, "tests", "test_pixell.py")): return d
	return None

# ------------------------------------------------------------------ astropy.wcs stand-in

DEG = np.pi/180

class _Wcsprm:
	def __init__(self, naxis):
		self.naxis = naxis
		self.crval = np.zeros(naxis); self.cdelt = np.ones(naxis); self.crpix = np.zeros(naxis)
		self.ctype = [""]*naxis; self.cunit = [""]*naxis
		self.lonpole = np.nan; self.latpole = np.nan; self._pv = []; self.name = ""
	def get_pv(self): return list(self._pv)
	def set_pv(self, pv): self._pv = list(pv)
	def has_cd(self): return False
	def has_pc(self): return False
	def bounds_check(self, a, b): pass
	def compare(self, other, cmp=0, tolerance=0.0):
		tol = max(tolerance, 0.0)
		return (list(self.ctype) == list(other.ctype) and np.allclose(self.crval, other.crval, rtol=0, atol=tol)
			and np.allclose(self.cdelt, other.cdelt, rtol=0, atol=tol) and np.allclose(self.crpix, other.crpix, rtol=0, atol=tol))
	def __setattr__(self, k, v):
		if k in ("crval", "cdelt", "crpix"): v = np.array(v, dtype=np.float64)
		elif k in ("ctype", "cunit"): v = [str(x) for x in v]
		object.__setattr__(self, k, v)

class FITSFixedWarning(Warning): pass

class WCS:
	"""astropy.wcs.WCS look-alike: linear ("plain"), CAR with the equator as reference latitude, TAN"""
	def __init__(self, header=None, naxis=2, **kw):
		self.naxis = naxis
		self.wcs = _Wcsprm(naxis)
		if header is not None:
			for i in range(naxis):
				for key, attr in (("CTYPE", "ctype"), ("CRVAL", "crval"), ("CDELT", "cdelt"), ("CRPIX", "crpix")):
					k = "%s%d" % (key, i+1)
					if k in header:
						a = getattr(self.wcs, attr); a[i] = header[k]; setattr(self.wcs, attr, a)
	def deepcopy(self):
		o = WCS(naxis=self.naxis)
		for k in ("crval", "cdelt", "crpix", "ctype", "cunit"): setattr(o.wcs, k, getattr(self.wcs, k))
		o.wcs.lonpole = self.wcs.lonpole; o.wcs.latpole = self.wcs.latpole; o.wcs._pv = list(self.wcs._pv)
		return o
	copy = deepcopy
	def __deepcopy__(self, memo): return self.deepcopy()
	def sub(self, axes): return self.deepcopy()
	def _proj(self):
		c = self.wcs.ctype[0]
		return c[-3:].upper() if len(c) >= 3 else ""
	def to_header(self, relax=None):
		h = {"WCSAXES": self.naxis}
		for i in range(self.naxis):
			h["CTYPE%d" % (i+1)] = self.wcs.ctype[i]; h["CRVAL%d" % (i+1)] = float(self.wcs.crval[i])
			h["CDELT%d" % (i+1)] = float(self.wcs.cdelt[i]); h["CRPIX%d" % (i+1)] = float(self.wcs.crpix[i])
		return h
	def to_header_string(self, relax=None): return repr(self.to_header())
	# intermediate world coordinates (degrees) <-> sky
	def _x2s(self, x, y):
		p = self._proj(); w = self.wcs
		if p == "" : return w.crval[0]+x, w.crval[1]+y
		if p == "CAR":
			if abs(w.crval[1]) > 1e-12: raise NotImplementedError("WCS stand-in: CAR with crval[1] != 0")
			return w.crval[0]+x, y
		if p == "TAN":
			R = np.hypot(x, y)*DEG
			phi = np.arctan2(x, -y); theta = np.arctan2(1.0, R)
			ap, dp, pp = w.crval[0]*DEG, w.crval[1]*DEG, np.pi
			st, ct = np.sin(theta), np.cos(theta); d = phi-pp
			dec = np.arcsin(np.clip(st*np.sin(dp) + ct*np.cos(dp)*np.cos(d), -1, 1))
			ra = ap + np.arctan2(-ct*np.sin(d), st*np.cos(dp) - ct*np.sin(dp)*np.cos(d))
			return ra/DEG, dec/DEG
		raise NotImplementedError("WCS stand-in: projection %s" % p)
	def _s2x(self, ra, dec):
		p = self._proj(); w = self.wcs
		if p == "": return ra-w.crval[0], dec-w.crval[1]
		if p == "CAR":
			if abs(w.crval[1]) > 1e-12: raise NotImplementedError("WCS stand-in: CAR with crval[1] != 0")
			return ra-w.crval[0], dec
		if p == "TAN":
			ap, dp, pp = w.crval[0]*DEG, w.crval[1]*DEG, np.pi
			a, d = ra*DEG-ap, dec*DEG
			phi = pp + np.arctan2(-np.cos(d)*np.sin(a), np.sin(d)*np.cos(dp) - np.cos(d)*np.sin(dp)*np.cos(a))
			st = np.sin(d)*np.sin(dp) + np.cos(d)*np.cos(dp)*np.cos(a)
			R = np.sqrt(np.maximum(1-st*st, 0))/st/DEG
			return R*np.sin(phi), -R*np.cos(phi)
		raise NotImplementedError("WCS stand-in: projection %s" % p)
	def _args(self, args):
		if len(args) == 2:
			a = np.asarray(args[0], dtype=np.float64); return [a[..., i] for i in range(a.shape[-1])], int(args[1]), True
		return [np.asarray(a, dtype=np.float64) for a in args[:-1]], int(args[-1]), False
	def wcs_pix2world(self, *args):
		cols, origin, packed = self._args(args)
		w = self.wcs
		x = (cols[0] + (1-origin) - w.crpix[0])*w.cdelt[0]; y = (cols[1] + (1-origin) - w.crpix[1])*w.cdelt[1]
		ra, dec = self._x2s(x, y)
		return np.stack([ra, dec], -1) if packed else [ra, dec]
	all_pix2world = wcs_pix2world
	def wcs_world2pix(self, *args):
		cols, origin, packed = self._args(args)
		w = self.wcs
		x, y = self._s2x(cols[0], cols[1])
		px = x/w.cdelt[0] + w.crpix[0] - (1-origin); py = y/w.cdelt[1] + w.crpix[1] - (1-origin)
		return np.stack([px, py], -1) if packed else [px, py]
	all_world2pix = wcs_world2pix
	def __repr__(self): return "WCS(%s)" % self.to_header_string()

# ------------------------------------------------------------------ module plumbing

class _Placeholder(types.ModuleType):
	"""an importable module whose attributes are placeholders that fail only when used"""
	__path__ = []
	def __getattr__(self, name):
		if name.startswith("__"): raise AttributeError(name)
		if self.__name__ == "matplotlib" and name == "use": return lambda *a, **k: None      # called at import time by the reference's tests
		def missing(*a, **k): raise ImportError("%s.%s is not available in this test environment" % (self.__name__, name))
		missing.__name__ = name
		return missing

def _module(name, **attrs):
	m = types.ModuleType(name)
	for k, v in attrs.items(): setattr(m, k, v)
	return m

PLACEHOLDERS = ["astropy.io", "astropy.io.fits", "astropy.coordinates", "astropy.time", "astropy.units", "astropy.utils", "healpy", "matplotlib",
	"matplotlib.pyplot", "h5py", "PIL", "PIL.Image", "PIL.ImageDraw", "PIL.ImageFont", "ephem", "numba", "mpi4py", "dateutil", "dateutil.parser",
	"pixell._interpol_32", "pixell._interpol_64", "pixell._colorize", "pixell._array_ops_32", "pixell._array_ops_64", "pixell.srcsim",
	"pixell.distances", "pixell.cmisc_core", "pixell._distances", "pixell._srcsim", "pixell.sharp"]

def make_ducc0(backend):
	"""ducc0 package object: ducc0.sht.experimental.<f> = backend.<f> (keyword-only functions, as pixell calls them)"""
	exp = _module("ducc0.sht.experimental")
	for name in ("synthesis", "adjoint_synthesis", "synthesis_2d", "adjoint_synthesis_2d", "analysis_2d", "adjoint_analysis_2d",
			"get_gridweights", "synthesis_general", "adjoint_synthesis_general"):
		if hasattr(backend, name): setattr(exp, name, getattr(backend, name))
	sht = _module("ducc0.sht", experimental=exp)
	if hasattr(backend, "rotate_alm"): sht.rotate_alm = backend.rotate_alm
	d = _module("ducc0", sht=sht, __version__="0.36.0 (stand-in)")
	d.__path__ = []
	for sub in ("fft", "nufft", "misc"): setattr(d, sub, _Placeholder("ducc0."+sub))
	return d, {"ducc0": d, "ducc0.sht": sht, "ducc0.sht.experimental": exp, "ducc0.fft": d.fft, "ducc0.nufft": d.nufft, "ducc0.misc": d.misc}

_installed = {}

def install(backend, cmisc_backend, fresh=True):
	"""Put the stand-ins into sys.modules and import the reference's pixell package; returns the dict of pixell modules"""
	ref = reference_root()
	if ref is None: raise RuntimeError("reference files are not staged: run scripts/stage_reference.py where /root/reference exists")
	if fresh:
		for k in [k for k in sys.modules if k == "pixell" or k.startswith("pixell.") or k == "ducc0" or k.startswith("ducc0.") or k == "astropy" or k.startswith("astropy.")]:
			del sys.modules[k]
	astropy = _module("astropy"); astropy.__path__ = []
	wcsmod = _module("astropy.wcs", WCS=WCS, FITSFixedWarning=FITSFixedWarning)
	astropy.wcs = wcsmod
	sys.modules["astropy"] = astropy; sys.modules["astropy.wcs"] = wcsmod
	for name in PLACEHOLDERS:
		if name.startswith("pixell."): continue
		if name not in sys.modules or name.startswith("astropy"):
			try:
				if not name.startswith("astropy"): importlib.import_module(name); continue
			except Exception: pass
			sys.modules[name] = _Placeholder(name)
			parent, _, leaf = name.rpartition(".")
			if parent in sys.modules: setattr(sys.modules[parent], leaf, sys.modules[name])
	d, mods = make_ducc0(backend)
	sys.modules.update(mods)
	# the package itself: the reference's files, with its compiled extension modules replaced
	pkg = types.ModuleType("pixell"); pkg.__path__ = [os.path.join(ref, "pixell")]; pkg.__file__ = os.path.join(ref, "pixell", "__init__.py")
	pkg.__version__ = "reference (unmodified sources, stand-in backends)"
	sys.modules["pixell"] = pkg
	sys.modules["pixell.cmisc"] = cmisc_backend; pkg.cmisc = cmisc_backend
	for name in PLACEHOLDERS:
		if name.startswith("pixell."):
			sys.modules[name] = _Placeholder(name); setattr(pkg, name.split(".", 1)[1], sys.modules[name])
	out = {}
	for name in ("bunch", "utils", "wcsutils", "powspec", "fft", "enmap", "curvedsky"):
		out[name] = importlib.import_module("pixell."+name)
	return out

def load_reference_tests():
	"""The reference's tests/test_pixell.py as a module.  Its import block pulls in plotting, point-source and tiling
	modules that need compiled extensions: those imports are satisfied by placeholders (they are not used by the tests run)."""
	ref = reference_root()
	path = os.path.join(ref, "tests", "test_pixell.py")
	for name in ("lensing", "array_ops", "enplot", "reproject", "pointsrcs", "colors", "tilemap", "interpol", "coordinates", "wavelets", "uharm", "analysis"):
		full = "pixell."+name
		if full in sys.modules: continue
		try: importlib.import_module(full)
		except Exception:
			sys.modules[full] = _Placeholder(full); setattr(sys.modules["pixell"], name, sys.modules[full])
	spec = importlib.util.spec_from_file_location("reference_test_pixell", path)
	mod = importlib.util.module_from_spec(spec)
	spec.loader.exec_module(mod)
	return mod

def run_reference_tests(names, verbosity=0):
	"""Run the named test methods of the reference's PixelTests class; returns the unittest result"""
	mod = load_reference_tests()
	cls = next(v for k, v in vars(mod).items() if isinstance(v, type) and issubclass(v, unittest.TestCase) and hasattr(v, names[0]))
	suite = unittest.TestSuite([cls(n) for n in names])
	with warnings.catch_warnings():
		warnings.simplefilter("ignore")
		return unittest.TextTestRunner(verbosity=verbosity, stream=open(os.devnull, "w")).run(suite)

# ------------------------------------------------------------------ CPU stand-in for pixell.cmisc (oracle-backed)

def oracle_cmisc():
	"""pixell.cmisc look-alike on top of oracle/alm_oracle.py (numpy), for the CPU run of the scaffolding"""
	from oracle import alm_oracle as ao
	m = types.ModuleType("pixell.cmisc")
	def _ai(ainfo): return ao.AlmInfo(ainfo.lmax, ainfo.mmax, stride=ainfo.stride, layout=np.asarray(ainfo.mstart).astype(np.int64))
	def alm2cl(ainfo, alm, alm2=None, cl_dtype=None):
		alm = np.asarray(alm); alm2 = alm if alm2 is None else np.asarray(alm2)
		pre = np.broadcast_shapes(alm.shape[:-1], alm2.shape[:-1])
		a1 = np.broadcast_to(alm, pre+alm.shape[-1:]); a2 = np.broadcast_to(alm2, pre+alm2.shape[-1:])
		cl = np.zeros(pre+(ainfo.lmax+1,), cl_dtype if cl_dtype is not None else alm.real.dtype)
		for I in np.ndindex(*pre): cl[I] = ao.alm2cl(_ai(ainfo), a1[I], a2[I], dtype=cl.dtype)
		return cl
	def lmul(ainfo, alm, lfun, out=None):
		alm = np.asarray(alm); lfun = np.asarray(lfun)
		if alm.dtype not in (np.complex64, np.complex128): raise ValueError("lmul requires complex64 or complex128 arrays")
		res = ao.lmul(_ai(ainfo), alm, lfun)
		if out is None: return res.astype(alm.dtype)
		out[...] = res
		return out
	def transpose_alm(ainfo, alm, out=None):
		res = ao.transpose_alm(_ai(ainfo), np.asarray(alm))
		if out is None: return res
		out[...] = res
		return out
	def transfer_alm(iainfo, ialm, oainfo, oalm=None, op=lambda a, b: b):
		ialm = np.asarray(ialm)
		if oalm is None: oalm = np.zeros(ialm.shape[:-1]+(oainfo.nelem,), ialm.dtype)
		lmax, mmax = min(iainfo.lmax, oainfo.lmax), min(iainfo.mmax, oainfo.mmax)
		for m in range(mmax+1):
			si = slice(int(iainfo.mstart[m])+m*iainfo.stride, int(iainfo.mstart[m])+(lmax+1)*iainfo.stride, iainfo.stride)
			so_ = slice(int(oainfo.mstart[m])+m*oainfo.stride, int(oainfo.mstart[m])+(lmax+1)*oainfo.stride, oainfo.stride)
			oalm[..., so_] = op(oalm[..., so_], ialm[..., si])
		return oalm
	m.alm2cl, m.lmul, m.transpose_alm, m.transfer_alm = alm2cl, lmul, transpose_alm, transfer_alm
	return m
