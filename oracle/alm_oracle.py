"""oracle/alm_oracle.py -- TEST INFRASTRUCTURE ONLY (CPU checker; never imported by the product).

numpy restatement of the alm helpers on the hot path:
  alm2cl, lmul (scalar + matrix), transpose_alm, transfer_alm   reference cython/cmisc_core.c:16-304,
                                                                 cython/cmisc.pyx:8-191
  rand_alm / rand_alm_white / fill_gauss                        reference pixell/curvedsky.py:61-77, 602-628
  rand_alm_healpy (scalar stream only)                          reference pixell/curvedsky.py:44-59 (healpy.synalm)
  eigpow-based matrix square root                               reference pixell/utils.py:2789-2830

When /root/reference is present, oracle/Makefile also compiles the reference's own
cmisc_core.c into oracle/_ref/libcmisc_ref.so (see ref_cmisc()), and tests check this file against it.
"""
import ctypes, os
import numpy as np

_here = os.path.dirname(os.path.abspath(__file__))

class AlmInfo:
	"""Layout of an alm array; restates curvedsky.alm_info (pixell/curvedsky.py:409-447)."""
	def __init__(self, lmax=None, mmax=None, nalm=None, stride=1, layout="triangular"):
		if isinstance(layout, str):
			if layout in ("triangular", "tri"):
				if lmax is None: lmax = int((-1+(1+8*nalm)**0.5)/2)-1
				if mmax is None: mmax = lmax
				m = np.arange(mmax+1)
				mstart = stride*(m*(2*lmax+1-m)//2)
			elif layout in ("rectangular", "rect"):
				if lmax is None: lmax = int(nalm**0.5)-1
				if mmax is None: mmax = lmax
				mstart = np.arange(mmax+1)*(lmax+1)*stride
			else: raise ValueError("unknown layout")
		else:
			mstart = np.asarray(layout)
		self.lmax, self.mmax, self.stride = int(lmax), int(mmax), int(stride)
		self.nelem = int(np.max(mstart) + (lmax+1)*stride)
		self.mstart = mstart.astype(np.int64)
	def lm2ind(self, l, m): return self.mstart[m] + l*self.stride

def alm2cl(ainfo, alm1, alm2=None, dtype=None):
	"""cmisc_core.c:77-110: cl[l] = 2/(2l+1) (Re a1_l0 Re a2_l0 / 2 + sum_{m>=1} Re(a1 conj a2))."""
	if alm2 is None: alm2 = alm1
	acc = np.float64 if dtype is None else dtype
	if dtype is None and alm1.dtype == np.complex64: acc = np.float32
	cl = np.zeros(ainfo.lmax+1, acc)
	l = np.arange(ainfo.lmax+1)
	i = ainfo.mstart[0] + l
	cl += (alm1[i].real*alm2[i].real/2).astype(acc)
	for m in range(1, ainfo.mmax+1):
		l = np.arange(m, ainfo.lmax+1); i = ainfo.mstart[m]+l
		cl[m:] += (alm1[i].real.astype(acc)*alm2[i].real + alm1[i].imag.astype(acc)*alm2[i].imag)
	return cl*(2.0/(2*np.arange(ainfo.lmax+1)+1)).astype(acc)

def lmul(ainfo, alm, lfun):
	"""cmisc_core.c:159-182 (scalar) / :185-230 (matrix).  alm[...,nalm]; lfun[nl] or [N,M,nl]."""
	alm = np.asarray(alm); lfun = np.asarray(lfun)
	lfmax = lfun.shape[-1]-1
	lof = np.zeros(ainfo.nelem, np.int64); used = np.zeros(ainfo.nelem, bool)
	for m in range(ainfo.mmax+1):
		l = np.arange(m, ainfo.lmax+1); lof[ainfo.mstart[m]+l] = l; used[ainfo.mstart[m]+l] = True
	fl = np.zeros(lfun.shape[:-1]+(ainfo.lmax+1,), lfun.dtype)
	n = min(lfmax, ainfo.lmax)+1
	fl[..., :n] = lfun[..., :n]
	if lfun.ndim == 3 and alm.ndim == 2:
		out = np.einsum("rcl,cl->rl", fl[:,:,lof], alm)
		out[:, ~used] = 0
		return out
	out = alm*fl[..., lof]
	out[..., ~used] = alm[..., ~used]
	return out

def transpose_alm(ainfo, alm):
	"""cmisc_core.c:116-135: the k-th element in m-major enumeration receives ... precisely: the
	k-th (im,il) input element is written to the address of the k-th pair in l-major order."""
	src = np.concatenate([ainfo.mstart[m] + np.arange(m, ainfo.lmax+1) for m in range(ainfo.mmax+1)])
	dst = np.concatenate([ainfo.mstart[:min(l, ainfo.mmax)+1] + l for l in range(ainfo.lmax+1)])
	out = np.zeros_like(alm)
	out[..., dst] = alm[..., src]
	return out

def transfer_alm(iainfo, ialm, oainfo, oalm=None):
	"""cmisc.pyx:131-150"""
	if oalm is None: oalm = np.zeros(ialm.shape[:-1]+(oainfo.nelem,), ialm.dtype)
	lmax = min(iainfo.lmax, oainfo.lmax); mmax = min(iainfo.mmax, oainfo.mmax)
	for m in range(mmax+1):
		l = np.arange(m, lmax+1)
		oalm[..., oainfo.mstart[m]+l*oainfo.stride] = ialm[..., iainfo.mstart[m]+l*iainfo.stride]
	return oalm

def eigpow_half(ps):
	"""ps[ncomp,ncomp,nl] -> symmetric square root per l, negative eigenvalues -> 0
	(enmap.multi_pow(ps,0.5): pixell/enmap.py:2021-2024, utils.py:2789-2830)."""
	A = np.moveaxis(np.asarray(ps, np.float64), -1, 0)
	E, V = np.linalg.eigh(A)
	mask = E < 0
	E = np.where(mask, 0, np.abs(E)**0.5)
	res = np.einsum("...ij,...kj->...ik", V*E[..., None, :], V)
	return np.moveaxis(res, 0, -1)

def rand_alm_white(ainfo, ncomp, seed, dtype=np.complex128):
	"""curvedsky.py:602-628: numpy legacy global RNG, real view filled in blocks of 65536 in
	l-major order, then transposed to m-major."""
	if seed is not None: np.random.seed(seed)
	alm = np.empty((ncomp, ainfo.nelem), dtype)
	rtype = np.zeros(0, dtype).real.dtype
	flat = alm.reshape(-1).view(rtype)
	for i in range(0, flat.size, 0x10000):
		flat[i:i+0x10000] = np.random.standard_normal(min(0x10000, flat.size-i))
	return transpose_alm(ainfo, alm)

def rand_alm(ps, lmax, seed, dtype=np.complex128):
	"""curvedsky.rand_alm (pixell/curvedsky.py:61-77).  ps is [nl] or [ncomp,ncomp,nl]."""
	ps = np.asarray(ps)
	wps = ps[None, None] if ps.ndim == 1 else ps
	if lmax > wps.shape[-1]-1:
		pad = np.zeros(wps.shape[:-1]+(lmax+1,), wps.dtype); pad[..., :wps.shape[-1]] = wps; wps = pad
	ainfo = AlmInfo(lmax)
	alm = rand_alm_white(ainfo, wps.shape[0], seed, dtype)
	rtype = np.zeros(0, dtype).real.dtype
	ps12 = eigpow_half(wps)
	alm = lmul(ainfo, alm, (ps12/2**0.5).astype(rtype)).astype(dtype)
	alm[:, :lmax+1] = alm[:, :lmax+1].real*2**0.5
	return alm[0] if ps.ndim == 1 else alm

def rand_alm_healpy_scalar(cl, lmax, seed):
	"""The scalar healpy.synalm(new=True) stream reached from curvedsky.rand_map
	(pixell/curvedsky.py:44-59) -- SURVEY.md Appendix A, pinned by MM_041121.pkl."""
	if seed is not None: np.random.seed(seed)
	nalm = (lmax+1)*(lmax+2)//2
	re = np.random.standard_normal(nalm); im = np.random.standard_normal(nalm)
	ainfo = AlmInfo(lmax)
	cl = np.asarray(cl, np.float64)
	c = np.zeros(lmax+1); c[:min(len(cl), lmax+1)] = cl[:lmax+1]
	lof = np.concatenate([np.arange(m, lmax+1) for m in range(lmax+1)])
	alm = (re + 1j*im)*np.sqrt(c[lof]/2)
	alm[:lmax+1] = re[:lmax+1]*np.sqrt(c[:lmax+1])
	return alm

# ----------------------------------------------------------------------------- reference cmisc

_ref = None
def ref_cmisc():
	"""ctypes handle on the reference's own cmisc_core.c (oracle/_ref/libcmisc_ref.so), or None."""
	global _ref
	if _ref is None:
		path = os.path.join(_here, "_ref", "libcmisc_ref.so")
		if not os.path.exists(path): return None
		_ref = ctypes.CDLL(path)
	return _ref

# ------------------------------------------------------------------ device random stream (b2_rand_alm), restated in numpy

def philox4x32_10(counter, seed):
	"""Philox4x32-10 (Salmon et al. 2011): counter = (lo32, hi32, 0, 0) of the uint64 array `counter`, key = the two halves
	of `seed`; returns the four output words as uint32 arrays"""
	c = np.asarray(counter, dtype=np.uint64)
	M32 = np.uint64(0xffffffff)
	c0 = c & M32; c1 = c >> np.uint64(32); c2 = np.zeros_like(c0); c3 = np.zeros_like(c0)
	k0 = np.uint64(int(seed) & 0xffffffff); k1 = np.uint64((int(seed) >> 32) & 0xffffffff)
	for r in range(10):
		p0 = np.uint64(0xD2511F53)*c0; p1 = np.uint64(0xCD9E8D57)*c2
		hi0, lo0 = p0 >> np.uint64(32), p0 & M32
		hi1, lo1 = p1 >> np.uint64(32), p1 & M32
		c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
		k0 = (k0 + np.uint64(0x9E3779B9)) & M32; k1 = (k1 + np.uint64(0xBB67AE85)) & M32
	return c0, c1, c2, c3

def philox_normal_pairs(counter, seed):
	"""Box-Muller on two 53-bit uniforms per counter: complex array re + i im of unit normals"""
	x0, x1, x2, x3 = philox4x32_10(counter, seed)
	a = (x1 << np.uint64(32)) | x0; b = (x3 << np.uint64(32)) | x2
	u1 = ((a >> np.uint64(11)).astype(np.float64) + 0.5)*2.0**-53
	u2 = ((b >> np.uint64(11)).astype(np.float64) + 0.5)*2.0**-53
	r = np.sqrt(-2.0*np.log(u1))
	return r*np.cos(2*np.pi*u2) + 1j*r*np.sin(2*np.pi*u2)

def rand_alm_philox(ainfo, ncomp, seed, ps12=None):
	"""What b2_rand_alm computes: white pairs in the reference's fill order (pixell/curvedsky.py:602-628: memory order of the
	l-major array, component after component), coloured with ps12/sqrt(2), m = 0 real with the sqrt(2) restored (:61-77)"""
	lmax, mmax = ainfo.lmax, ainfo.mmax
	nlm = sum(lmax-m+1 for m in range(mmax+1))
	alm = np.zeros((ncomp, ainfo.nelem), np.complex128)
	for m in range(mmax+1):
		l = np.arange(m, lmax+1, dtype=np.int64)
		j = np.where(l <= mmax, l*(l+1)//2+m, (mmax+1)*(mmax+2)//2 + (l-mmax-1)*(mmax+1) + m)
		w = np.array([philox_normal_pairs((c*nlm+j).astype(np.uint64), seed) for c in range(ncomp)])
		if ps12 is not None:
			v = np.einsum("rcl,cl->rl", np.asarray(ps12)[:, :, l]/2**0.5, w)
			if m == 0: v = v.real*2**0.5 + 0j
		else: v = w
		alm[:, ainfo.mstart[m]+l*ainfo.stride] = v
	return alm
