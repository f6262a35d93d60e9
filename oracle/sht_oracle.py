"""oracle/sht_oracle.py -- TEST INFRASTRUCTURE ONLY (CPU checker; never imported by the product).

CPU restatement of the ducc0 entry points pixell's curvedsky calls
(reference pixell/curvedsky.py:328-329, 501, 531-555, 855, 907-924, 936-960, 1032-1046,
1068-1084).  ducc0 (PyPI "ducc0>=0.36.0", reference pyproject.toml:27) is a third-party
dependency absent from /root/reference, so its *published* algorithm is restated here:

  synthesis           alm --Legendre--> leg[comp,ring,m] --phase e^{i m phi0}, c2r FFT--> map
  adjoint_synthesis   exact transpose of the above
  analysis_2d         ring r2c FFT -> leg -> (theta-resampling to a Clenshaw-Curtis grid with
                      >= 2 lmax+2 rings when the grid is too coarse for direct quadrature)
                      -> quadrature weights -> adjoint Legendre
  adjoint_analysis_2d exact transpose of analysis_2d
  get_gridweights     interpolatory ring weights (sum = 4 pi)

The Legendre stage is C (sht_oracle.c, built by oracle/Makefile into oracle/_build/); the FFTs
use scipy.fft (pocketfft = the same FFT ducc ships).  Parity is pinned against the reference's
golden fixtures in tests/test_oracle_golden.py.
"""
import ctypes, os, subprocess
import numpy as np
import scipy.fft as sfft

_here = os.path.dirname(os.path.abspath(__file__))
_lib = None

def build(force=False):
	"""Compile sht_oracle.c -> oracle/_build/libshtoracle.so (gcc -O3 -fopenmp)."""
	out = os.path.join(_here, "_build", "libshtoracle.so")
	src = os.path.join(_here, "sht_oracle.c")
	if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
		os.makedirs(os.path.dirname(out), exist_ok=True)
		subprocess.check_call(["gcc", "-O3", "-march=native", "-fopenmp", "-shared", "-fPIC",
			"-o", out, src, "-lm"])
	return out

def _load():
	global _lib
	if _lib is None:
		_lib = ctypes.CDLL(build())
		i64p = ctypes.POINTER(ctypes.c_int64); dp = ctypes.POINTER(ctypes.c_double)
		_lib.orc_alm2leg.argtypes = [ctypes.c_int]*4 + [i64p, ctypes.c_int, dp, dp, ctypes.c_int64, dp]
		_lib.orc_leg2alm.argtypes = [ctypes.c_int]*4 + [i64p, ctypes.c_int, dp, dp, dp, ctypes.c_int64]
	return _lib

def nthreads():
	return _load().orc_num_threads()

# ------------------------------------------------------------------ tuned CPU Legendre stage (bench.py's CPU arm)

_fast = None
def build_fast(force=False):
	"""Compile sht_fast.c -> oracle/_build/libshtfast.so for the cores of THIS machine (-march=native, 512-bit vectors
	where the CPU has them).  bench.py's CPU arm builds it on the box it runs on."""
	out = os.path.join(_here, "_build", "libshtfast.so")
	src = os.path.join(_here, "sht_fast.c")
	tag = out + ".cpu"
	cpu = ""
	try:
		with open("/proc/cpuinfo") as f:
			for line in f:
				if line.startswith("model name"): cpu = line.split(":", 1)[1].strip(); break
	except OSError: pass
	stale = not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src)
	if not stale:
		try: stale = open(tag).read() != cpu
		except OSError: stale = True
	if force or stale:
		os.makedirs(os.path.dirname(out), exist_ok=True)
		subprocess.check_call(["gcc", "-O3", "-march=native", "-mprefer-vector-width=512", "-fopenmp", "-shared", "-fPIC",
			"-o", out, src, "-lm"])
		with open(tag, "w") as f: f.write(cpu)
	return out

def _load_fast():
	global _fast
	if _fast is None:
		_fast = ctypes.CDLL(build_fast())
		i64p = ctypes.POINTER(ctypes.c_int64); dp = ctypes.POINTER(ctypes.c_double); ip = ctypes.POINTER(ctypes.c_int)
		_fast.fast_alm2leg.argtypes = [ctypes.c_int]*3 + [i64p, ctypes.c_int, ip, ctypes.c_int, dp, dp, ctypes.c_int64, dp]
		_fast.fast_leg2alm.argtypes = [ctypes.c_int]*3 + [i64p, ctypes.c_int, ip, ctypes.c_int, dp, dp, dp, ctypes.c_int64]
	return _fast

def fast_nthreads(): return _load_fast().fast_num_threads()

def fast_alm2leg(alm, theta, spin, lmax, mmax, mstart, mlist=None):
	"""alm[nca, nalm] c128 -> leg[ncm, len(mlist), nring] c128 (ring fastest), only for the m in mlist (default all)"""
	lib = _load_fast()
	alm = np.ascontiguousarray(alm, dtype=np.complex128); theta = np.ascontiguousarray(theta, dtype=np.float64)
	mstart = np.ascontiguousarray(mstart).astype(np.int64)
	mlist = np.arange(mmax+1, dtype=np.int32) if mlist is None else np.ascontiguousarray(mlist, dtype=np.int32)
	ncm = 1 if spin == 0 else 2
	assert alm.shape[0] == ncm
	leg = np.empty((ncm, len(mlist), len(theta)), np.complex128)
	err = lib.fast_alm2leg(spin, lmax, mmax, _ip(mstart), len(mlist), mlist.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
		len(theta), _dp(theta), _dp(alm.view(np.float64)), alm.shape[1], _dp(leg.view(np.float64)))
	if err: raise ValueError("fast_alm2leg failed")
	return leg

def fast_leg2alm(leg, theta, spin, lmax, mmax, mstart, nalm, mlist=None, alm=None):
	"""leg[ncm, len(mlist), nring] -> alm[nca, nalm] (entries of the listed m only; others untouched / zero)"""
	lib = _load_fast()
	leg = np.ascontiguousarray(leg, dtype=np.complex128); theta = np.ascontiguousarray(theta, dtype=np.float64)
	mstart = np.ascontiguousarray(mstart).astype(np.int64)
	mlist = np.arange(mmax+1, dtype=np.int32) if mlist is None else np.ascontiguousarray(mlist, dtype=np.int32)
	ncm = 1 if spin == 0 else 2
	assert leg.shape == (ncm, len(mlist), len(theta))
	if alm is None: alm = np.zeros((ncm, nalm), np.complex128)
	err = lib.fast_leg2alm(spin, lmax, mmax, _ip(mstart), len(mlist), mlist.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
		len(theta), _dp(theta), _dp(leg.view(np.float64)), _dp(alm.view(np.float64)), alm.shape[1])
	if err: raise ValueError("fast_leg2alm failed")
	return alm

def set_mstride(s):
	"""bench.py only: restrict the Legendre stage to every s-th m (bounded CPU-baseline sample)"""
	_load().orc_set_mstride(int(s))

def _dp(a): return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
def _ip(a): return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))

def default_mstart(lmax, mmax):
	m = np.arange(mmax+1, dtype=np.int64)
	return m*(2*lmax+1-m)//2

# ------------------------------------------------------------------ Legendre stage (C)

def alm2leg(alm, theta, spin, lmax, mmax, mstart, deriv1=False):
	"""alm[ncomp_alm,nalm] c128 -> leg[ncomp_map,nring,mmax+1] c128"""
	lib = _load()
	alm = np.ascontiguousarray(alm, dtype=np.complex128)
	assert alm.ndim == 2
	theta = np.ascontiguousarray(theta, dtype=np.float64)
	mstart = np.ascontiguousarray(mstart).astype(np.int64)
	ncm = 1 if spin == 0 else 2
	assert alm.shape[0] == (1 if (spin == 0 or deriv1) else 2)
	assert alm.shape[1] >= int(mstart.max()) + lmax + 1
	leg = np.zeros((ncm, len(theta), mmax+1), np.complex128)
	err = lib.orc_alm2leg(spin, int(deriv1), lmax, mmax, _ip(mstart), len(theta), _dp(theta),
		_dp(alm.view(np.float64)), alm.shape[1], _dp(leg.view(np.float64)))
	if err: raise ValueError("orc_alm2leg failed")
	return leg

def leg2alm(leg, theta, spin, lmax, mmax, mstart, nalm, deriv1=False, alm=None):
	lib = _load()
	leg = np.ascontiguousarray(leg, dtype=np.complex128)
	theta = np.ascontiguousarray(theta, dtype=np.float64)
	mstart = np.ascontiguousarray(mstart).astype(np.int64)
	nca = 1 if (spin == 0 or deriv1) else 2
	if alm is None: alm = np.zeros((nca, nalm), np.complex128)
	assert alm.flags["C_CONTIGUOUS"] and alm.dtype == np.complex128 and alm.shape == (nca, nalm)
	err = lib.orc_leg2alm(spin, int(deriv1), lmax, mmax, _ip(mstart), len(theta), _dp(theta),
		_dp(leg.view(np.float64)), _dp(alm.view(np.float64)), alm.shape[1])
	if err: raise ValueError("orc_leg2alm failed")
	return alm

# ------------------------------------------------------------------ ring FFT stage (numpy)

def leg2map(leg, nphi, phi0, workers=-1):
	"""leg[ncomp,nring,nm] -> map[ncomp,nring,nphi]; pixel j of a ring sits at phi0 + 2 pi j/nphi.
	map = sum_{m>=0} w_m Re[leg_m e^{i m phi}], w_0 = 1, w_{m>0} = 2; |m| aliased mod nphi."""
	ncomp, nring, nm = leg.shape
	m = np.arange(nm)
	phi0 = np.broadcast_to(np.asarray(phi0, dtype=np.float64), (nring,))
	G = leg*np.exp(1j*m[None,:]*phi0[:,None])[None]
	if nphi >= 2*nm-1:
		return sfft.irfft(G, n=nphi, axis=-1, workers=workers)*nphi
	F = np.zeros((ncomp, nring, nphi), np.complex128)
	F[...,0] = G[...,0].real
	np.add.at(F, (slice(None), slice(None), m[1:] % nphi), G[...,1:])
	np.add.at(F, (slice(None), slice(None), (-m[1:]) % nphi), np.conj(G[...,1:]))
	return sfft.ifft(F, axis=-1, workers=workers).real*nphi

def map2leg(map, nm, phi0, workers=-1):
	"""Exact transpose of leg2map: leg_m = e^{-i m phi0} sum_j map_j e^{-2 pi i m j/nphi}."""
	ncomp, nring, nphi = map.shape
	m = np.arange(nm)
	phi0 = np.broadcast_to(np.asarray(phi0, dtype=np.float64), (nring,))
	if nphi >= 2*nm-1:
		F = sfft.rfft(map, axis=-1, workers=workers)[...,:nm]
	else:
		F = sfft.fft(map, axis=-1, workers=workers)[..., m % nphi]
	return F*np.exp(-1j*m[None,:]*phi0[:,None])[None]

def _nworkers():
	return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)

def map2leg_t(map, nm, phi0):
	"""map2leg with the result in the Legendre stage's layout [ncomp, nm, nring], blocks of rings spread over a thread
	pool (the FFT, the phase factors and the transposition all run in parallel) -- bench.py's CPU arm"""
	import concurrent.futures
	ncomp, nring, nphi = map.shape
	out = np.empty((ncomp, nm, nring), np.complex128)
	w = _nworkers(); step = max(1, min(128, (nring+w-1)//w))
	def work(r0):
		out[:, :, r0:r0+step] = map2leg(map[:, r0:r0+step], nm, phi0, workers=1).transpose(0, 2, 1)
	with concurrent.futures.ThreadPoolExecutor(max_workers=w) as ex: list(ex.map(work, range(0, nring, step)))
	return out

def leg2map_t(leg, nphi, phi0):
	"""leg2map from the layout [ncomp, nm, nring], blocks of rings spread over a thread pool"""
	import concurrent.futures
	ncomp, nm, nring = leg.shape
	out = np.empty((ncomp, nring, nphi), np.float64)
	w = _nworkers(); step = max(1, min(128, (nring+w-1)//w))
	def work(r0):
		out[:, r0:r0+step] = leg2map(np.ascontiguousarray(leg[:, :, r0:r0+step].transpose(0, 2, 1)), nphi, phi0, workers=1)
	with concurrent.futures.ThreadPoolExecutor(max_workers=w) as ex: list(ex.map(work, range(0, nring, step)))
	return out

# ------------------------------------------------------------------ grids and weights

def grid_theta(name, n):
	"""Ring colatitudes of ducc's named equiangular grids (north -> south).
	Pole offsets as in pixell/curvedsky.py:1334-1342."""
	k = np.arange(n)
	if   name == "CC":     return k*np.pi/(n-1) if n > 1 else np.zeros(1)
	elif name == "F1":     return (k+0.5)*np.pi/n
	elif name == "MW":     return (2*k+1)*np.pi/(2*n-1)
	elif name == "MWflip": return 2*k*np.pi/(2*n-1)
	elif name == "DH":     return k*np.pi/n
	elif name == "F2":     return (k+1)*np.pi/(n+1)
	raise ValueError("unknown geometry %s" % name)

def _circle(name, n):
	"""(N, offset, mult) describing the grid as every-point subset of an N-point circle grid
	theta_k = (k+offset) 2 pi/N; mult_k = number of circle points mapping to ring k."""
	mult = np.full(n, 2.0)
	if   name == "CC":     N, o = 2*(n-1), 0.0; mult[0] = mult[-1] = 1
	elif name == "F1":     N, o = 2*n, 0.5
	elif name == "MW":     N, o = 2*n-1, 0.5; mult[-1] = 1
	elif name == "MWflip": N, o = 2*n-1, 0.0; mult[0] = 1
	elif name == "DH":     N, o = 2*n, 0.0; mult[0] = 1
	elif name == "F2":     N, o = 2*(n+1), 1.0
	else: raise ValueError("unknown geometry %s" % name)
	return N, o, mult

def get_gridweights(name, n):
	"""Interpolatory quadrature weights for ring integrals on the named grid (sum = 4 pi):
	sum_k w_k f(theta_k) = 2 pi int_0^pi f sin(theta) dtheta for trig polynomials f up to the
	degree the rule supports.  Replaces ducc0.sht.experimental.get_gridweights
	(pixell/curvedsky.py:501, 531, 855).
	CC/F1/MW/MWflip: truncated Fourier series of |sin theta| sampled on the grid's circle
	extension.  F2: classical Fejer-2 rule; DH(n) = [0, F2(n-1)].
	tests/test_oracle_basic.py checks the moment equations directly."""
	if name == "DH":
		return np.concatenate([[0.0], get_gridweights("F2", n-1)])
	theta = grid_theta(name, n)
	if name == "F2":
		j = np.arange(1, (n+1)//2+1)
		s = np.zeros(n)
		for j0 in range(0, len(j), 256):   # chunked to bound memory
			jj = j[j0:j0+256]
			s += np.sum(np.sin((2*jj[None,:]-1)*theta[:,None])/(2*jj[None,:]-1), 1)
		return 2*np.pi*(4.0/(n+1))*np.sin(theta)*s
	N, o, mult = _circle(name, n)
	K = n-1                       # degree of exactness of an n-node interpolatory rule
	H = np.zeros(N//2+1, np.complex128)
	H[0] = 1.0
	for j in range(1, K//2+1):
		c = 2.0
		if name == "CC" and 2*j == n-1: c = 1.0   # top harmonic aliases onto DC: halve it
		H[2*j] = -c/(4.0*j*j-1)
	H *= np.exp(1j*np.arange(N//2+1)*o*2*np.pi/N)
	full = np.zeros(N, np.complex128); full[:N//2+1] = H
	g = (sfft.ifft(full)*N).real[:n]
	return g*mult*(np.pi/N)*(2/np.pi)*2*np.pi

def gridweights_bruteforce(name, n):
	"""Solve the moment equations sum_k w_k cos(j theta_k) = 2 pi int_0^pi cos(j t) sin t dt, j < n."""
	theta = grid_theta(name, n)
	j = np.arange(n)
	A = np.cos(j[:,None]*theta[None,:])
	rhs = np.where(j % 2 == 0, 2.0/(1.0-j.astype(float)**2+(j==1)), 0.0)
	rhs[j == 1] = 0
	return np.linalg.solve(A, 2*np.pi*rhs)

# ------------------------------------------------------------------ theta resampling

def _ext_index(name, n):
	"""Indices describing the even/odd extension to the circle: ring k -> circle slot pos[k],
	mirror slot mir[k] (or -1)."""
	N, o, _ = _circle(name, n)
	k = np.arange(n)
	if name == "F2": pos = k+1
	else:            pos = k
	if o == 0.5:  mir = N-1-pos
	else:         mir = (N-pos) % N
	mir = np.where(mir == pos, -1, mir)
	return N, o, pos, mir

def _good_cc_size(nmin):
	"""Smallest nt >= nmin such that 2(nt-1) is a fast FFT length."""
	N = sfft.next_fast_len(2*(nmin-1), real=False)
	while N % 2: N = sfft.next_fast_len(N+1, real=False)
	return N//2+1

def resample_to_cc(leg, name, nt, spin, workers=-1):
	"""leg[ncomp,n,nm] on grid `name` -> leg on CC grid with nt rings, by trigonometric
	interpolation of the (-1)^(m+spin)-symmetric extension in theta."""
	ncomp, n, nm = leg.shape
	N, o, pos, mir = _ext_index(name, n)
	assert name in ("CC", "F1", "MW", "MWflip")
	sig = (-1.0)**(np.arange(nm)+spin)
	ext = np.zeros((ncomp, N, nm), np.complex128)
	ext[:, pos] = leg
	ok = mir >= 0
	ext[:, mir[ok]] = leg[:, ok]*sig
	C = sfft.fft(ext, axis=1, workers=workers)/N
	k = sfft.fftfreq(N, 1.0/N)          # signed frequencies
	C *= np.exp(-1j*k*o*2*np.pi/N)[None,:,None]
	Np = 2*(nt-1)
	assert Np >= N
	Cp = np.zeros((ncomp, Np, nm), np.complex128)
	ki = k.astype(int)
	if N % 2 == 0:
		sel = ki != -N//2
		Cp[:, ki[sel] % Np] = C[:, sel]
		nyq = C[:, N//2]
		if Np > N: Cp[:, N//2] += 0.5*nyq; Cp[:, Np-N//2] += 0.5*nyq
		else:      Cp[:, N//2] += nyq
	else:
		Cp[:, ki % Np] = C
	out = sfft.ifft(Cp, axis=1, workers=workers)*Np
	return out[:, :nt]

def _resample_chunk_t(leg, name, nt, sig, N, o, pos, mir):
	ncomp, nm, n = leg.shape
	ext = np.zeros((ncomp, nm, N), np.complex128)
	ext[:, :, pos[0]:pos[0]+n] = leg
	ok = mir >= 0
	ext[:, :, mir[ok]] = leg[:, :, ok]*sig[None, :, None]
	C = sfft.fft(ext, axis=-1, overwrite_x=True)
	k = sfft.fftfreq(N, 1.0/N)
	C *= (np.exp(-1j*k*o*2*np.pi/N)/N)[None, None, :]
	Np = 2*(nt-1)
	Cp = np.zeros((ncomp, nm, Np), np.complex128)
	if N % 2 == 0:
		h = N//2
		Cp[:, :, :h] = C[:, :, :h]; Cp[:, :, Np-h+1:] = C[:, :, h+1:]
		nyq = C[:, :, h]
		if Np > N: Cp[:, :, h] += 0.5*nyq; Cp[:, :, Np-h] += 0.5*nyq
		else:      Cp[:, :, h] += nyq
	else:
		h = (N+1)//2
		Cp[:, :, :h] = C[:, :, :h]; Cp[:, :, Np-(N-h):] = C[:, :, h:]
	out = sfft.ifft(Cp, axis=-1, overwrite_x=True)
	out *= Np
	return out[:, :, :nt]

def resample_to_cc_t(leg, name, nt, spin, workers=None):
	"""resample_to_cc on the transposed layout leg[ncomp, nm, n] -> [ncomp, nm, nt] (theta contiguous: what the Legendre
	stage of sht_fast.c reads and writes; no strided FFTs, no transposes), blocks of m columns spread over a thread pool
	(numpy and scipy.fft release the GIL) -- the CPU arm of bench.py uses this form"""
	import concurrent.futures
	ncomp, nm, n = leg.shape
	N, o, pos, mir = _ext_index(name, n)
	assert name in ("CC", "F1", "MW", "MWflip") and 2*(nt-1) >= N
	sig = (-1.0)**(np.arange(nm)+spin)
	if workers is None: workers = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
	out = np.empty((ncomp, nm, nt), np.complex128)
	step = max(1, min(64, (nm+workers-1)//workers))
	def work(m0):
		out[:, m0:m0+step] = _resample_chunk_t(leg[:, m0:m0+step], name, nt, sig[m0:m0+step], N, o, pos, mir)
	with concurrent.futures.ThreadPoolExecutor(max_workers=workers) as ex:
		list(ex.map(work, range(0, nm, step)))
	return out

def resample_to_cc_adjoint(legcc, name, n, spin, workers=-1):
	"""Hermitian transpose of resample_to_cc: leg on CC(nt) -> leg on `name`(n)."""
	ncomp, nt, nm = legcc.shape
	N, o, pos, mir = _ext_index(name, n)
	Np = 2*(nt-1)
	z = np.zeros((ncomp, Np, nm), np.complex128)
	z[:, :nt] = legcc
	Cp = sfft.fft(z, axis=1, workers=workers)
	k = sfft.fftfreq(N, 1.0/N); ki = k.astype(int)
	C = np.zeros((ncomp, N, nm), np.complex128)
	if N % 2 == 0:
		sel = ki != -N//2
		C[:, sel] = Cp[:, ki[sel] % Np]
		if Np > N: C[:, N//2] = 0.5*(Cp[:, N//2] + Cp[:, Np-N//2])
		else:      C[:, N//2] = Cp[:, N//2]
	else:
		C[:] = Cp[:, ki % Np]
	C *= np.exp(1j*k*o*2*np.pi/N)[None,:,None]
	ext = sfft.ifft(C, axis=1, workers=workers)
	sig = (-1.0)**(np.arange(nm)+spin)
	out = ext[:, pos].copy()
	ok = mir >= 0
	out[:, ok] += ext[:, mir[ok]]*sig
	return out

# ------------------------------------------------------------------ ducc-shaped entry points

def _prep_alm_in(alm, spin, mode):
	alm = np.asarray(alm)
	assert alm.ndim == 2
	return alm.astype(np.complex128, copy=False)

def synthesis(*, alm, theta, nphi, phi0, ringstart, spin, lmax, mmax=None, mstart=None,
		map=None, mode="STANDARD", nthreads=0, **kw):
	"""ducc0.sht.experimental.synthesis restatement (call sites pixell/curvedsky.py:553, 937, 1068, 1079).
	Rings may have different nphi/phi0; pixel (r,j) lives at ringstart[r]+j."""
	if mmax is None: mmax = lmax
	if mstart is None: mstart = default_mstart(lmax, mmax)
	deriv1 = mode == "DERIV1"
	theta = np.asarray(theta, np.float64); nphi = np.asarray(nphi).astype(np.int64)
	phi0 = np.asarray(phi0, np.float64); ringstart = np.asarray(ringstart).astype(np.int64)
	leg = alm2leg(_prep_alm_in(alm, spin, mode), theta, spin, lmax, mmax, mstart, deriv1)
	ncm = leg.shape[0]
	npix = int(np.max(ringstart+nphi))
	if map is None: map = np.zeros((ncm, npix), np.float64 if np.asarray(alm).dtype == np.complex128 else np.float32)
	for n in np.unique(nphi):
		sel = np.where(nphi == n)[0]
		vals = leg2map(leg[:, sel], int(n), phi0[sel])
		idx = ringstart[sel][:,None] + np.arange(n)[None,:]
		map[:, idx] = vals
	return map

def adjoint_synthesis(*, map, theta, nphi, phi0, ringstart, spin, lmax, mmax=None, mstart=None,
		alm=None, mode="STANDARD", nthreads=0, **kw):
	"""ducc0.sht.experimental.adjoint_synthesis restatement (pixell/curvedsky.py:537, 936, 1069, 1080)."""
	if mmax is None: mmax = lmax
	if mstart is None: mstart = default_mstart(lmax, mmax)
	deriv1 = mode == "DERIV1"
	theta = np.asarray(theta, np.float64); nphi = np.asarray(nphi).astype(np.int64)
	phi0 = np.asarray(phi0, np.float64); ringstart = np.asarray(ringstart).astype(np.int64)
	map = np.asarray(map)
	ncm = map.shape[0]
	leg = np.zeros((ncm, len(theta), mmax+1), np.complex128)
	for n in np.unique(nphi):
		sel = np.where(nphi == n)[0]
		idx = ringstart[sel][:,None] + np.arange(n)[None,:]
		leg[:, sel] = map2leg(map[:, idx].astype(np.float64), mmax+1, phi0[sel])
	nalm = int(np.max(np.asarray(mstart).astype(np.int64))) + lmax + 1 if alm is None else alm.shape[-1]
	res = leg2alm(leg, theta, spin, lmax, mmax, mstart, nalm, deriv1)
	if alm is None: return res.astype(np.result_type(map.dtype, 0j))
	alm[...] = res
	return alm

def synthesis_2d(*, alm, spin, lmax, geometry, ntheta=None, nphi=None, mmax=None, mstart=None,
		phi0=0.0, map=None, mode="STANDARD", nthreads=0, **kw):
	"""ducc0.sht.experimental.synthesis_2d restatement (pixell/curvedsky.py:908)."""
	if mmax is None: mmax = lmax
	if mstart is None: mstart = default_mstart(lmax, mmax)
	if map is not None: ntheta, nphi = map.shape[-2:]
	theta = grid_theta(geometry, ntheta)
	leg = alm2leg(_prep_alm_in(alm, spin, mode), theta, spin, lmax, mmax, mstart, mode == "DERIV1")
	res = leg2map(leg, nphi, phi0)
	if map is None: return res
	map[...] = res
	return map

def adjoint_synthesis_2d(*, map, spin, lmax, geometry, mmax=None, mstart=None, phi0=0.0,
		alm=None, mode="STANDARD", nthreads=0, **kw):
	"""ducc0.sht.experimental.adjoint_synthesis_2d restatement (pixell/curvedsky.py:907)."""
	if mmax is None: mmax = lmax
	if mstart is None: mstart = default_mstart(lmax, mmax)
	map = np.asarray(map)
	theta = grid_theta(geometry, map.shape[-2])
	leg = map2leg(map.astype(np.float64), mmax+1, phi0)
	nalm = int(np.max(np.asarray(mstart).astype(np.int64))) + lmax + 1 if alm is None else alm.shape[-1]
	res = leg2alm(leg, theta, spin, lmax, mmax, mstart, nalm, mode == "DERIV1")
	if alm is None: return res
	alm[...] = res
	return alm

def _needs_resample(geometry, ntheta, lmax):
	if geometry in ("DH", "F2"): return False
	return ntheta < 2*lmax+2

def maxlmax(geometry, ny):
	"""pixell/curvedsky.py:1349-1353"""
	if   geometry == "CC": return ny-2
	elif geometry == "DH": return (ny-2)//2
	elif geometry == "F2": return (ny-1)//2
	else:                  return ny-1

def analysis_2d(*, map, spin, lmax, geometry, mmax=None, mstart=None, phi0=0.0, alm=None,
		nthreads=0, **kw):
	"""ducc0.sht.experimental.analysis_2d restatement (pixell/curvedsky.py:573, 1033): exact
	inversion of synthesis_2d for band-limited maps."""
	if mmax is None: mmax = lmax
	if mstart is None: mstart = default_mstart(lmax, mmax)
	map = np.asarray(map)
	ntheta = map.shape[-2]
	if lmax > maxlmax(geometry, ntheta): raise ValueError("lmax too large for this grid")
	leg = map2leg(map.astype(np.float64), mmax+1, phi0)
	if _needs_resample(geometry, ntheta, lmax):
		nt = _good_cc_size(2*lmax+2)
		leg = resample_to_cc(leg, geometry, nt, spin)
		theta = grid_theta("CC", nt); w = get_gridweights("CC", nt)
	else:
		theta = grid_theta(geometry, ntheta); w = get_gridweights(geometry, ntheta)
	leg = leg*(w/map.shape[-1])[None,:,None]
	nalm = int(np.max(np.asarray(mstart).astype(np.int64))) + lmax + 1 if alm is None else alm.shape[-1]
	res = leg2alm(leg, theta, spin, lmax, mmax, mstart, nalm)
	if alm is None: return res
	alm[...] = res
	return alm

def adjoint_analysis_2d(*, alm, spin, lmax, geometry, ntheta=None, nphi=None, mmax=None, mstart=None,
		phi0=0.0, map=None, nthreads=0, **kw):
	"""ducc0.sht.experimental.adjoint_analysis_2d restatement (pixell/curvedsky.py:1032)."""
	if mmax is None: mmax = lmax
	if mstart is None: mstart = default_mstart(lmax, mmax)
	if map is not None: ntheta, nphi = map.shape[-2:]
	alm = _prep_alm_in(alm, spin, "STANDARD")
	if _needs_resample(geometry, ntheta, lmax):
		nt = _good_cc_size(2*lmax+2)
		theta = grid_theta("CC", nt); w = get_gridweights("CC", nt)
		leg = alm2leg(alm, theta, spin, lmax, mmax, mstart)*(w/nphi)[None,:,None]
		leg = resample_to_cc_adjoint(leg, geometry, ntheta, spin)
	else:
		theta = grid_theta(geometry, ntheta); w = get_gridweights(geometry, ntheta)
		leg = alm2leg(alm, theta, spin, lmax, mmax, mstart)*(w/nphi)[None,:,None]
	res = leg2map(leg, nphi, phi0)
	if map is None: return res
	map[...] = res
	return map

# ------------------------------------------------------------------ arbitrary positions

def synthesis_general(*, alm, loc, spin, lmax, mmax=None, mstart=None, mode="STANDARD", **kw):
	"""ducc0.sht.experimental.synthesis_general restatement (call site pixell/curvedsky.py:993-1016): the series evaluated
	directly at loc[npos, 2] = (theta, phi), as a ring synthesis with one single-pixel ring per position (no NUFFT, so
	this is exact to rounding; cost O(npos lmax^2))."""
	loc = np.asarray(loc, np.float64)
	n = len(loc)
	return synthesis(alm=alm, theta=loc[:, 0], nphi=np.ones(n, np.int64), phi0=loc[:, 1], ringstart=np.arange(n),
		spin=spin, lmax=lmax, mmax=mmax, mstart=mstart, mode=mode)

def adjoint_synthesis_general(*, map, loc, spin, lmax, mmax=None, mstart=None, mode="STANDARD", **kw):
	"""transpose of synthesis_general (ducc0.sht.experimental.adjoint_synthesis_general, same call site with adjoint=True)"""
	loc = np.asarray(loc, np.float64)
	n = len(loc)
	return adjoint_synthesis(map=map, theta=loc[:, 0], nphi=np.ones(n, np.int64), phi0=loc[:, 1], ringstart=np.arange(n),
		spin=spin, lmax=lmax, mmax=mmax, mstart=mstart, mode=mode)
