"""oracle/parity.py -- TEST INFRASTRUCTURE ONLY (checker; imported by tests/ and by bench.py's CPU leg after the timed
region, never by the product).

Parity of the CUDA path against the CPU oracle AT THE BASELINE SIZES (C2: lmax 4096, C3: lmax 8000), where the
oracle cannot afford the whole l,m triangle: the oracle runs on a comb of m values (every k-th m, all l, all rings),
the CUDA engine on everything, and the comb columns are compared.

  comb_legendre   b2_alm2leg / b2_leg2alm on the real plan (full random alm / leg) against sht_oracle.c's
                  orc_alm2leg / orc_leg2alm on the comb columns, all rings
  comb_transform  alm supported on the comb only: oracle synthesis (Legendre comb + scipy ring FFTs, all rings, all
                  pixels) against the CUDA synthesis_2d; then the CUDA analysis_2d of the ORACLE's map against the
                  input alm (every l,m; the off-comb entries must come back as zeros)
Reference call sites restated: pixell/curvedsky.py:900-962 (alm2map_raw_2d / _cyl), 1018-1046 (map2alm_raw_2d).
"""
import ctypes, time
import numpy as np
from . import sht_oracle as so

def _comb(lmax, mstride):
	return np.arange(0, lmax+1, mstride, dtype=np.int64)

def _rel(a, b):
	return float(np.abs(a-b).max()/max(np.abs(b).max(), 1e-300))

def _rel2(a, b):
	return float(np.sqrt((np.abs(a-b)**2).sum()/max((np.abs(b)**2).sum(), 1e-300)))

def comb_legendre(geometry, ny, nx, lmax, spin, mstride, seed=0):
	"""max relative errors of the CUDA Legendre kernels on the comb columns: dict(alm2leg, leg2alm, m_sampled, rings)"""
	import torch
	from pixell_b200 import sht, _lib as L
	ms = _comb(lmax, mstride)
	theta = so.grid_theta(geometry, ny)
	mstart = so.default_mstart(lmax, lmax); nalm = (lmax+1)*(lmax+2)//2
	nc = 1 if spin == 0 else 2
	rng = np.random.default_rng(seed)
	plan = sht.plan_2d(geometry, ny, nx, 0.0, lmax)
	nring_pad = (ny+31)//32*32
	# ---- alm -> leg: full random alm on the device, comb columns against the oracle
	alm = np.empty((nc, nalm), np.complex128)
	alm.real = rng.standard_normal((nc, nalm)); alm.imag = rng.standard_normal((nc, nalm))
	alm[:, :lmax+1] = alm[:, :lmax+1].real
	talm = torch.from_numpy(alm).cuda()
	leg = torch.zeros((nc, lmax+1, nring_pad), dtype=torch.complex128, device="cuda")
	L.check(L.lib().b2_alm2leg(plan.handle, spin, L.MODE_STANDARD, talm.data_ptr(), nalm if nc > 1 else 0, leg.data_ptr(), None))
	got = leg[:, torch.from_numpy(ms).cuda(), :ny].cpu().numpy()                # [nc, ncomb, ny]
	so.set_mstride(mstride)
	try: want = so.alm2leg(alm, theta, spin, lmax, lmax, mstart)[:, :, ms].transpose(0, 2, 1)
	finally: so.set_mstride(1)
	e_syn = _rel(got, want)
	del want, got
	# ---- leg -> alm: full random leg on the device, the comb's alm entries against the oracle
	g = torch.Generator(device="cuda"); g.manual_seed(seed+1)
	leg = torch.randn((nc, lmax+1, nring_pad), dtype=torch.complex128, device="cuda", generator=g)
	leg[:, :, ny:] = 0
	talm.zero_()
	L.check(L.lib().b2_leg2alm(plan.handle, spin, L.MODE_STANDARD, talm.data_ptr(), nalm if nc > 1 else 0, leg.data_ptr(), None))
	got = talm.cpu().numpy()
	hleg = np.zeros((nc, ny, lmax+1), np.complex128)
	hleg[:, :, ms] = leg[:, torch.from_numpy(ms).cuda(), :ny].cpu().numpy().transpose(0, 2, 1)
	so.set_mstride(mstride)
	try: want = so.leg2alm(hleg, theta, spin, lmax, lmax, mstart, nalm)
	finally: so.set_mstride(1)
	idx = np.concatenate([mstart[m] + np.arange(max(m, spin), lmax+1) for m in ms])
	e_adj = _rel(got[:, idx], want[:, idx])
	return dict(alm2leg=e_syn, leg2alm=e_adj, m_sampled=int(len(ms)), rings=int(ny))

def comb_transform(geometry, ny, nx, lmax, spin, mstride, seed=0, phi0=0.0, dtype=np.float64):
	"""alm supported on the m comb: CUDA synthesis_2d against the oracle's map on every pixel, CUDA analysis_2d of the
	oracle's map against the input alm on every l,m.  dict(synthesis, analysis, analysis_l2, m_sampled, rings)"""
	import torch
	from pixell_b200 import sht
	ms = _comb(lmax, mstride)
	theta = so.grid_theta(geometry, ny)
	mstart = so.default_mstart(lmax, lmax); nalm = (lmax+1)*(lmax+2)//2
	nc = 1 if spin == 0 else 2
	rng = np.random.default_rng(seed)
	alm = np.zeros((nc, nalm), np.complex128)
	for m in ms:
		l = np.arange(max(m, spin, 2), lmax+1)
		if len(l) == 0: continue
		amp = 1.0/np.sqrt(l*(l+1.0))
		v = (rng.standard_normal((nc, len(l))) + (1j*rng.standard_normal((nc, len(l))) if m > 0 else 0))*amp
		alm[:, mstart[m]+l] = v
	so.set_mstride(mstride)
	try: leg = so.alm2leg(alm, theta, spin, lmax, lmax, mstart)
	finally: so.set_mstride(1)
	want = so.leg2map(leg, nx, phi0)                       # [nc, ny, nx] float64, all rings, all pixels
	del leg
	cdt = np.complex128 if dtype == np.float64 else np.complex64
	kw = dict(spin=spin, lmax=lmax, mstart=mstart, geometry=geometry, phi0=phi0)
	tmap = sht.synthesis_2d(alm=torch.from_numpy(alm.astype(cdt)).cuda(), ntheta=ny, nphi=nx, **kw)
	got = tmap.cpu().numpy()
	e_syn = _rel(got, want)
	del got, tmap
	back = sht.analysis_2d(map=torch.from_numpy(want.astype(dtype)).cuda(), **kw).cpu().numpy()
	return dict(synthesis=e_syn, analysis=_rel(back, alm), analysis_l2=_rel2(back, alm), m_sampled=int(len(ms)), rings=int(ny))

def baseline_parity(workload, mstride=None):
	"""bench.py: the checks above on one BASELINE workload (dict with ny, nx, lmax, ncomp); returns the JSON record"""
	lmax, ny, nx = workload["lmax"], workload["ny"], workload["nx"]
	if mstride is None: mstride = max(1, (lmax+1)//16)
	spins = [0, 2] if workload["ncomp"] == 3 else [0]
	t0 = time.perf_counter()
	rec = {"max_rel": 0.0, "m_sampled": 0, "rings_sampled": ny, "mstride": int(mstride), "checks": {}}
	for s in spins:
		a = comb_legendre("F1", ny, nx, lmax, s, mstride, seed=10+s)
		b = comb_transform("F1", ny, nx, lmax, s, mstride, seed=20+s)
		rec["checks"]["spin%d" % s] = {"alm2leg": a["alm2leg"], "leg2alm": a["leg2alm"], "synthesis_2d": b["synthesis"], "analysis_2d": b["analysis"]}
		rec["max_rel"] = max(rec["max_rel"], a["alm2leg"], a["leg2alm"], b["synthesis"], b["analysis"])
		rec["m_sampled"] = a["m_sampled"]
	rec["seconds"] = time.perf_counter()-t0
	rec["oracle"] = "oracle/sht_oracle.c on every %d-th m (all l, all rings) + scipy ring FFTs; tolerance 1e-10" % mstride
	return rec
