"""pixell_b200.cmisc -- GPU stand-in for the pixell.cmisc extension module (reference
cython/cmisc.pyx:8-246 over cython/cmisc_core.c): alm2cl, lmul, transpose_alm, transfer_alm with
the same argument meaning, broadcasting rules and exceptions.  Arrays may be numpy (host) or
torch CUDA tensors (device, zero-copy)."""
import numpy as np
from . import _lib as L

def _dt(alm):
	dt = L.buffer_info(alm)[2]
	if dt == np.complex128: return L.F64
	if dt == np.complex64: return L.F32
	raise ValueError("Only complex64 and complex128 supported, but got %s" % str(dt))

def _mstart(ainfo): return L.as_i64(np.asarray(ainfo.mstart))

def _xp(a):
	if L.is_torch(a):
		import torch
		return torch
	return np

def _contig_last(a):
	if L.is_torch(a): return a if a.stride(-1) == 1 else a.contiguous()
	return a if a.strides[-1] == a.itemsize else np.ascontiguousarray(a)

def _rows(a):
	"""iterate index tuples over the leading axes"""
	return np.ndindex(*a.shape[:-1])

def alm2cl(ainfo, alm, alm2=None, cl_dtype=None):
	"""cmisc.pyx:8-83: broadcasting cross spectra cl[..., l]; duplicate (alm, alm2) pairs are computed once."""
	L.init()
	tor = L.is_torch(alm)
	if not tor:
		alm = np.asarray(alm); alm2 = np.asarray(alm2) if alm2 is not None else alm
		dtype = np.result_type(alm, alm2)
		alm, alm2 = alm.astype(dtype, copy=False), alm2.astype(dtype, copy=False)
		a1, a2 = np.broadcast_arrays(alm, alm2)
	else:
		import torch
		if alm2 is None: alm2 = alm
		if alm.dtype != alm2.dtype: raise ValueError("alm and alm2 must share a dtype")
		a1, a2 = torch.broadcast_tensors(alm, alm2)
	dt = _dt(a1)
	rdt = np.float64 if dt == L.F64 else np.float32
	cl_dtype = np.dtype(rdt if cl_dtype is None else cl_dtype)
	if dt == L.F64 and cl_dtype == np.float32:
		raise TypeError("alm is double-prec but cl is single-prec, downgrading accumulation precision not allowed")
	if cl_dtype not in (np.float32, np.float64): raise ValueError("cl_dtype must be float32 or float64")
	pshape = tuple(a1.shape[:-1])
	if tor:
		import torch
		cl = torch.empty(pshape+(ainfo.lmax+1,), dtype=torch.float64 if cl_dtype == np.float64 else torch.float32, device=a1.device)
	else: cl = np.empty(pshape+(ainfo.lmax+1,), cl_dtype)
	ms = _mstart(ainfo)
	cache, keep = {}, []
	for I in np.ndindex(*pshape):
		x, y = _contig_last(a1[I]), _contig_last(a2[I]); keep += [x, y]
		px, mem, _ = L.buffer_info(x); py = L.buffer_info(y)[0]
		key = tuple(sorted([px, py]))
		if key in cache: cl[I] = cl[cache[key]]; continue
		pc = L.buffer_info(cl[I])[0]
		L.check(L.lib().b2_alm2cl(ainfo.lmax, ainfo.mmax, L.p_i64(ms), dt, px, py, L.F64 if cl_dtype == np.float64 else L.F32,
			pc, mem, L.current_stream(x)))
		cache[key] = I
	return cl

def _check_out(out, alm, shape, what):
	"""a caller-supplied output must have alm's dtype (the kernel and the staging sizes follow alm), the right shape and a
	contiguous last axis (cmisc.pyx:168-175 raises ValueError for these)"""
	if L.is_torch(out) != L.is_torch(alm): raise ValueError("%s: out and alm must both be numpy arrays or both be tensors" % what)
	if out.dtype != alm.dtype: raise ValueError("%s's out argument must have the same dtype as alm (%s), got %s" % (what, alm.dtype, out.dtype))
	if tuple(out.shape) != tuple(shape): raise ValueError("%s's out argument must have shape %s, got %s" % (what, tuple(shape), tuple(out.shape)))
	last = out.stride(-1) == 1 if L.is_torch(out) else out.strides[-1] == out.itemsize
	if out.shape[-1] > 1 and not last: raise ValueError("%s's out argument must be contiguous along last axis" % what)

def lmul(ainfo, alm, lfun, out=None):
	"""cmisc.pyx:159-191: alm[..., lm] * lfun[..., l], or the matrix product lfun[r,c,l] alm[c,lm]
	when lfun is 3-d and alm 2-d.  Returns out (allocated if None)."""
	L.init()
	tor = L.is_torch(alm)
	xp = _xp(alm)
	if not tor:
		alm = np.asarray(alm)
		ctype = np.result_type(alm.dtype, 0j); rtype = np.zeros(1, ctype).real.dtype
		alm = _contig_last(alm.astype(ctype, copy=False)); lfun = _contig_last(np.asarray(lfun, dtype=rtype))
	else:
		alm = _contig_last(alm)
		if not L.is_torch(lfun):
			import torch
			lfun = torch.as_tensor(np.asarray(lfun), device=alm.device)
		lfun = _contig_last(lfun.to(alm.real.dtype))
	dt = _dt(alm)
	if dt not in (L.F64, L.F32): raise ValueError("lmul requires complex64 or complex128 arrays")
	ms = _mstart(ainfo)
	lfmax = lfun.shape[-1]-1
	if lfun.ndim == 3 and alm.ndim == 2:
		N, M = lfun.shape[:2]
		if M != alm.shape[0]: raise ValueError("lmul: matrix and alm component counts differ")
		if out is None: out = xp.zeros((N,)+tuple(alm.shape[1:]), dtype=alm.dtype, device=alm.device) if tor else np.zeros((N,)+alm.shape[1:], alm.dtype)
		else: _check_out(out, alm, (N,)+tuple(alm.shape[1:]), "lmul")
		lf = lfun.contiguous() if tor else np.ascontiguousarray(lfun)
		pa, mem, _ = L.buffer_info(alm); po = L.buffer_info(out)[0]; pf = L.buffer_info(lf)[0]
		acs = L.strides_elems(alm)[0] if M > 1 else alm.shape[-1]
		ocs = L.strides_elems(out)[0] if N > 1 else out.shape[-1]
		L.check(L.lib().b2_lmatmul(N, M, ainfo.lmax, ainfo.mmax, L.p_i64(ms), dt, pa, acs, lfmax, pf, po, ocs, mem, L.current_stream(alm)))
		return out
	try: pre = np.broadcast_shapes(tuple(alm.shape[:-1]), tuple(lfun.shape[:-1]))
	except ValueError:
		raise ValueError("lmul's alm and lfun's dimensions must either broadcast (when ignoring the last dimension), or have shape compatible with a matrix product (again ignoring the last dimension)")
	if out is not None: _check_out(out, alm, tuple(pre)+(alm.shape[-1],), "lmul")
	if tor:
		ab = alm.expand(pre+(alm.shape[-1],)); lb = lfun.expand(pre+(lfun.shape[-1],))
		if out is None: out = ab.clone()
		else: out.copy_(ab)
	else:
		ab = np.broadcast_to(alm, pre+alm.shape[-1:]); lb = np.broadcast_to(lfun, pre+lfun.shape[-1:])
		if out is None: out = np.array(ab)
		else: out[...] = ab
	if not (out.stride(-1) == 1 if tor else out.strides[-1] == out.itemsize):
		raise ValueError("lmul's out argument must be contiguous along last axis, and have the same dtype as alm")
	for I in np.ndindex(*pre):
		f = _contig_last(lb[I])
		po, mem, _ = L.buffer_info(out[I]); pf = L.buffer_info(f)[0]
		L.check(L.lib().b2_lmul(ainfo.lmax, ainfo.mmax, L.p_i64(ms), dt, po, lfmax, pf, mem, L.current_stream(out)))
	return out

def transpose_alm(ainfo, alm, out=None):
	"""cmisc.pyx:98-126: l-major stored values -> m-major layout; in place when out is alm."""
	L.init()
	dt = _dt(alm)
	if out is None: out = alm.clone() if L.is_torch(alm) else alm.copy()
	ms = _mstart(ainfo)
	for I in _rows(alm):
		src = _contig_last(alm[I])
		tmp = src.clone() if L.is_torch(src) else src.copy()      # the kernel needs distinct buffers
		dst = out[I]
		direct = (dst.stride(-1) == 1) if L.is_torch(dst) else (dst.strides[-1] == dst.itemsize)
		work = dst if direct else (tmp.clone() if L.is_torch(tmp) else tmp.copy())
		pi, mem, _ = L.buffer_info(tmp); po = L.buffer_info(work)[0]
		L.check(L.lib().b2_transpose_alm(ainfo.lmax, ainfo.mmax, L.p_i64(ms), dt, pi, po, mem, L.current_stream(tmp)))
		if not direct: out[I] = work
	return out

def transfer_alm(iainfo, ialm, oainfo, oalm=None, op=None):
	"""cmisc.pyx:131-150: copy between layouts; op(dest, src) combines (default: overwrite)."""
	L.init()
	dt = _dt(ialm)
	xp = _xp(ialm)
	if oalm is None:
		oalm = xp.zeros(tuple(ialm.shape[:-1])+(oainfo.nelem,), dtype=ialm.dtype, **({"device": ialm.device} if L.is_torch(ialm) else {}))
	if tuple(ialm.shape[:-1]) != tuple(oalm.shape[:-1]): raise ValueError("ialm and oalm must agree on pre-dimensions")
	ims, oms = _mstart(iainfo), _mstart(oainfo)
	for I in _rows(ialm):
		src = _contig_last(ialm[I]); dst = oalm[I]
		dcontig = (dst.stride(-1) == 1) if L.is_torch(dst) else (dst.strides[-1] == dst.itemsize)
		if dst.dtype != src.dtype: raise ValueError("transfer_alm: ialm and oalm must have the same dtype")
		if op is None and dcontig: work = dst
		elif op is None: work = dst.clone() if L.is_torch(dst) else dst.copy()      # strided destination: staged through a contiguous copy
		else: work = xp.zeros_like(dst)
		pi, mem, _ = L.buffer_info(src); po = L.buffer_info(work)[0]
		L.check(L.lib().b2_transfer_alm(iainfo.lmax, iainfo.mmax, L.p_i64(ims), iainfo.stride, pi,
			oainfo.lmax, oainfo.mmax, L.p_i64(oms), oainfo.stride, po, dt, mem, L.current_stream(src)))
		if op is None and not dcontig: oalm[I] = work
		if op is not None:
			# apply op on the transferred entries only
			lmax, mmax = min(iainfo.lmax, oainfo.lmax), min(iainfo.mmax, oainfo.mmax)
			for m in range(mmax+1):
				sl = slice(int(oms[m])+m*oainfo.stride, int(oms[m])+(lmax+1)*oainfo.stride, oainfo.stride)
				dst[sl] = op(dst[sl], work[sl])
	return oalm
