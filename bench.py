#!/usr/bin/env python
"""bench.py -- SHT pairs/sec (map2alm + alm2map, 3-component CAR) at lmax, on N B200 GPUs.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c2|c1]
  (N > 1: launched by torch.distributed.run, one rank per GPU)

A step is one round trip of the hot path on one synthetic full-sky Fejer-1 CAR map:
  alm = curvedsky.map2alm(map, lmax, spin=[0,2])  (exact 2d analysis)  then  curvedsky.alm2map(alm, map).
Workload c3 (default; the configuration BASELINE.json's metric is quoted on): T,Q,U 8192 x 16384,
lmax = 8000, float64.  Every rank transforms its own independent map (the path shards over
independent maps/components with no data-path collective: weak scaling).

value  = pairs/s with map and alm resident in HBM (torch CUDA tensors through pixell_b200.curvedsky)
e2e    = the same through the same API with pinned HOST arrays (H2D/D2H inside the timed region)
roofline      = the dominant kernel (spin-2 Legendre adjoint) against the measured HBM peak, as the
                contract asks; it is FP64-FMA bound, so roofline_fp64 gives the fraction of the
                measured DFMA peak too
cpu_baseline  = the CPU oracle (C/OpenMP restatement, not ducc0: ducc0 cannot be installed here)
                timed on the host cores on a bounded sample of the same workload
--impl reference times that CPU path alone (rank 0), same metric and config.
"""
import argparse, json, os, subprocess, sys, threading, time
import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
	"c3": dict(ny=8192, nx=16384, lmax=8000, ncomp=3, name="C3: T,Q,U full-sky Fejer-1 CAR 8192x16384, lmax=8000, spin [0,2], f64"),
	"c2": dict(ny=4608, nx=9216, lmax=4096, ncomp=1, name="C2: T full-sky Fejer-1 CAR 4608x9216, lmax=4096, spin 0, f64"),
	"c1": dict(ny=512, nx=1024, lmax=256, ncomp=1, name="C1: T full-sky Fejer-1 CAR 512x1024, lmax=256, spin 0, f64"),
}

def peaks():
	try:
		with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f: return json.load(f), "measured"
	except Exception:
		return {"hbm_gbs": 6650.0}, "fallback"

def recorded_traffic(kernel):
	"""dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu
	--set full capture (profiles/roofline_traffic.json); None if there is no capture for this kernel"""
	try:
		with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f: t = json.load(f)
		return t.get(kernel)
	except Exception: return None

def algorithmic_bytes(w):
	nalm = (w["lmax"]+1)*(w["lmax"]+2)//2
	one = 16*w["ncomp"]*nalm + 8*w["ncomp"]*w["ny"]*w["nx"]
	return 2*one          # B_pair

def canonical_flops(w, ncomp_t, ncomp_p):
	"""8 flop per (l,m,ring pair) per scalar component, 24 per spin-2 pair"""
	nalm = (w["lmax"]+1)*(w["lmax"]+2)//2
	return (8*ncomp_t + 24*ncomp_p)*nalm*((w["ny"]+1)//2)

class ClockSampler:
	"""nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""
	def __init__(self, dev):
		self.rows = []; self.proc = None; self.dev = dev
	def start(self):
		q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
		try:
			self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.dev), "--query-gpu="+q, "--format=csv,noheader,nounits", "-lms", "200"],
				stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
			self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
		except Exception: self.proc = None
	def _read(self):
		for line in self.proc.stdout: self.rows.append(line.strip())
	def stop(self):
		if self.proc is None: return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
		self.proc.terminate()
		try: self.proc.wait(timeout=5)
		except Exception: pass
		sm, smax, reasons = [], None, set()
		names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
		for r in self.rows:
			f = [x.strip() for x in r.split(",")]
			if len(f) < 7: continue
			try: sm.append(float(f[0])); smax = float(f[1])
			except ValueError: continue
			for n, v in zip(names, f[3:7]):
				if v.lower().startswith("active"): reasons.add(n)
		return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}

def make_inputs(w, seed, torch, device):
	"""band-limited synthetic map: random alm with C_l ~ 1/(l(l+1)) -> alm2map"""
	from pixell_b200 import curvedsky, geometry
	lmax, ncomp = w["lmax"], w["ncomp"]
	shape, wcs = geometry.fullsky_geometry(shape=(w["ny"], w["nx"]))
	ainfo = curvedsky.alm_info(lmax)
	g = torch.Generator(device=device); g.manual_seed(seed)
	alm = torch.randn((ncomp, ainfo.nelem), dtype=torch.complex128, device=device, generator=g)
	l = torch.cat([torch.arange(m, lmax+1, device=device) for m in range(lmax+1)]) if lmax <= 512 else None
	if l is None:
		# l of every element, built without a Python loop over m
		m_of = torch.repeat_interleave(torch.arange(lmax+1, device=device), torch.arange(lmax+1, 0, -1, device=device))
		start = torch.as_tensor(np.asarray(ainfo.mstart).astype(np.int64), device=device)
		l = torch.arange(ainfo.nelem, device=device) - start[m_of]
	amp = torch.where(l >= 2, 1.0/torch.sqrt((l*(l+1)).clamp(min=1).double()), torch.zeros((), dtype=torch.float64, device=device))
	alm *= amp
	alm[:, :lmax+1] = alm[:, :lmax+1].real.to(torch.complex128)
	map = torch.empty((ncomp, w["ny"], w["nx"]), dtype=torch.float64, device=device)
	spin = [0, 2] if ncomp == 3 else [0]
	curvedsky.alm2map(alm, map, spin=spin, wcs=wcs)
	return alm, map, wcs, ainfo, spin

def run_ours(args, w):
	import torch
	import torch.distributed as dist
	rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
	if not torch.cuda.is_available(): raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
	torch.cuda.set_device(local)
	device = torch.device("cuda", local)
	if world > 1:
		# NCCL may print its version banner on stdout when the communicator is created: keep stdout to the one JSON line
		sys.stdout.flush()
		saved = os.dup(1); os.dup2(2, 1)
		try:
			dist.init_process_group("nccl", device_id=device)
			dist.barrier()
			torch.cuda.synchronize()
		finally:
			sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)
	from pixell_b200 import curvedsky, _lib as L, sht
	L.init(local)
	alm0, map, wcs, ainfo, spin = make_inputs(w, 3+rank, torch, device)
	alm = torch.zeros_like(alm0)
	lmax = w["lmax"]

	def step_device():
		curvedsky.map2alm(map, alm, spin=spin, wcs=wcs, ainfo=ainfo)
		curvedsky.alm2map(alm, map, spin=spin, wcs=wcs, ainfo=ainfo)

	def barrier():
		torch.cuda.synchronize()
		if world > 1: dist.barrier()
		torch.cuda.synchronize()

	for _ in range(args.warmup): step_device()
	# parity guard inside the bench: the round trip must give the input alm back
	err = ((alm-alm0).abs().max()/alm0.abs().max()).item()
	sampler = ClockSampler(local)
	if rank == 0: sampler.start()
	n0 = L.lib().b2_launch_count()
	barrier()
	e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	e0.record()
	for _ in range(args.steps): step_device()
	e1.record()
	barrier()
	ms = e0.elapsed_time(e1)
	launches = L.lib().b2_launch_count() - n0
	# per-kernel timing of the dominant kernel (spin-2 / spin-0 Legendre adjoint), measured live with the
	# library's own CUDA events on the launching stream during one more step
	curvedsky.map2alm(map, alm, spin=spin, wcs=wcs, ainfo=ainfo)
	torch.cuda.synchronize()
	plan = next(reversed(sht._plans.values()))
	tim = plan.last_timing()          # of the last spin group executed (spin 2 for T,Q,U)
	clocks = sampler.stop() if rank == 0 else None

	# ---- e2e: same API, pinned host arrays
	hmap = torch.empty(map.shape, dtype=torch.float64, pin_memory=True); hmap.copy_(map)
	halm = torch.empty(alm.shape, dtype=torch.complex128, pin_memory=True)
	nmap, nalm = hmap.numpy(), halm.numpy()
	from pixell_b200 import geometry
	nmap = geometry.ndmap(nmap, wcs)
	def step_host():
		curvedsky.map2alm(nmap, nalm, spin=spin, ainfo=ainfo)
		curvedsky.alm2map(nalm, nmap, spin=spin, ainfo=ainfo)
	step_host()
	barrier()
	nsteps_e2e = max(1, min(args.steps, 3))
	t0 = time.perf_counter()
	for _ in range(nsteps_e2e): step_host()
	torch.cuda.synchronize()
	t_e2e = (time.perf_counter()-t0)/nsteps_e2e
	err_e2e = float(np.abs(nalm-alm0.cpu().numpy()).max()/alm0.abs().max().item())

	tt = torch.tensor([ms, t_e2e*1e3], dtype=torch.float64, device=device)
	if world > 1: dist.all_reduce(tt, op=dist.ReduceOp.MAX)
	ms, e2e_ms = tt[0].item(), tt[1].item()
	if rank == 0:
		pk, pk_kind = peaks()
		ms_per_step = ms/args.steps
		value = world/(ms_per_step*1e-3)
		mapbytes = 8*w["ncomp"]*w["ny"]*w["nx"]; almbytes = 16*w["ncomp"]*ainfo.nelem
		# dominant kernel: Legendre adjoint of the last spin group
		ncq = 2 if w["ncomp"] == 3 else 1
		kbytes = 16*ncq*ainfo.nelem + 8*ncq*w["ny"]*w["nx"]
		kflops = canonical_flops(w, 0, 1) if w["ncomp"] == 3 else canonical_flops(w, 1, 0)
		kms = tim["legendre"]
		import ctypes
		dpk = ctypes.c_double(); L.check(L.lib().b2_dfma_peak_gflops(ctypes.byref(dpk)))
		out = {
			"metric": "SHT pairs/sec (map2alm+alm2map, 3-comp CAR) at lmax; %HBM roofline",
			"value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
			"ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
			"dtype": "f64", "data": "synthetic",
			"config": {"workload": w["name"], "lmax": lmax, "shape": [w["ncomp"], w["ny"], w["nx"]],
				"l2": "inputs larger than L2 (map %.2f GB, alm %.2f GB per step)" % (mapbytes/1e9, almbytes/1e9),
				"sharding": "one independent map per GPU, no data-path collective"},
			"roundtrip_rel_err": err,
			"gpu_launches": int(launches),
			"e2e": {"value": world/(e2e_ms*1e-3), "unit": "pairs/s", "h2d_bytes_per_step": int(mapbytes+almbytes),
				"d2h_bytes_per_step": int(mapbytes+almbytes), "roundtrip_rel_err": err_e2e, "host_memory": "pinned"},
			"roofline": {"kernel": "k_adj2 (Legendre adjoint, spin 2)" if w["ncomp"] == 3 else "k_adj0 (Legendre adjoint, spin 0)",
				"bound": "hbm", "achieved": kbytes/(kms*1e-3)/1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
				"frac": kbytes/(kms*1e-3)/1e9/pk["hbm_gbs"], "traffic": recorded_traffic("k_adj2" if w["ncomp"] == 3 and w["lmax"] == 8000 else ""), "peak_source": pk_kind,
				"ms_per_launch": kms, "algorithmic_bytes": kbytes,
				"note": "FP64-FMA bound kernel (intensity ~ lmax/12 flop/byte): see roofline_fp64"},
			"roofline_fp64": {"bound": "fp64", "achieved": kflops/(kms*1e-3)/1e12, "peak": dpk.value/1e3, "unit": "TFLOP/s",
				"frac": kflops/(kms*1e-3)/1e9/dpk.value, "peak_source": "DFMA microbenchmark in this run",
				"flops_model": "canonical, SURVEY.md 8d (no credit for polar skipping)"},
			"stage_ms_last_map2alm_group": tim,
			"clocks": clocks,
		}
		if world == 1 and not args.no_cpu:
			out["cpu_baseline"] = cpu_baseline(w, budget_s=args.cpu_seconds)
		print(json.dumps(out), flush=True)
	if world > 1: dist.destroy_process_group()

# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores, on a bounded sample of the same workload.

def cpu_pair_seconds(w, mstride, ring_frac_rows):
	"""Time one map2alm + alm2map pair of the oracle on a sample: every mstride-th m in the Legendre stage
	(work per m is ~ (lmax - m + 1), so a regular comb samples the triangle uniformly) and the first
	`ring_frac_rows` rings in the FFT stages; returns the estimated full-pair seconds."""
	from oracle import sht_oracle as so
	lmax, ny, nx, ncomp = w["lmax"], w["ny"], w["nx"], w["ncomp"]
	theta = so.grid_theta("F1", ny)
	mstart = so.default_mstart(lmax, lmax)
	nalm = (lmax+1)*(lmax+2)//2
	rng = np.random.default_rng(0)
	so.set_mstride(mstride)
	ms = np.arange(0, lmax+1, mstride)
	frac_m = np.sum(lmax-ms+1.0)/np.sum(lmax-np.arange(lmax+1)+1.0)
	t_leg = 0.0
	groups = [(0, 1)] + ([(2, 2)] if ncomp == 3 else [])
	try:
		for spin, nc in groups:
			alm = (rng.standard_normal((nc, nalm)) + 1j*rng.standard_normal((nc, nalm)))
			t0 = time.perf_counter()
			leg = so.alm2leg(alm, theta, spin, lmax, lmax, mstart)              # synthesis Legendre
			t1 = time.perf_counter()
			# exact analysis in the reference runs the Legendre stage on >= 2 lmax + 2 rings (ducc analysis_2d)
			nt = so._good_cc_size(2*lmax+2)
			so.leg2alm(np.zeros((nc, nt, lmax+1), complex), so.grid_theta("CC", nt), spin, lmax, lmax, mstart, nalm)
			t2 = time.perf_counter()
			t_leg += (t1-t0) + (t2-t1)
			del alm, leg
	finally:
		so.set_mstride(1)
	t_leg /= frac_m
	# FFT stages on a subset of rings, all cores (scipy.fft = pocketfft, the FFT ducc ships)
	nr = max(8, int(ny*ring_frac_rows))
	m = rng.standard_normal((ncomp, nr, nx))
	t0 = time.perf_counter()
	leg = so.map2leg(m, lmax+1, 0.1)
	so.leg2map(leg, nx, 0.1)
	t_fft = (time.perf_counter()-t0)*ny/nr
	return t_leg + t_fft, dict(t_legendre_est=t_leg, t_fft_est=t_fft, m_fraction=float(frac_m), rings=nr)

def use_all_host_cores():
	"""torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core.  Must run before the
	oracle's OpenMP runtime is loaded."""
	n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
	os.environ["OMP_NUM_THREADS"] = str(n)
	return n

def cpu_baseline(w, budget_s=20.0):
	use_all_host_cores()
	from oracle import sht_oracle as so
	so.build()
	cores = so.nthreads()
	# pick the m comb so that the sample costs about budget_s: probe with a coarse comb first
	probe = max(1, (w["lmax"]+1)//16)
	t0 = time.perf_counter(); est, info = cpu_pair_seconds(w, probe, 0.002); tp = time.perf_counter()-t0
	stride = max(1, int(probe*tp/max(budget_s, 1e-3)))
	if stride < probe:
		est, info = cpu_pair_seconds(w, stride, 0.01)
	else: stride = probe
	return {"value": 1.0/est, "unit": "pairs/s", "cores": cores, "kind": "port",
		"sample": "oracle (C/OpenMP restatement, not ducc0) on every %d-th m of the Legendre stage (%.1f%% of the l,m triangle) "
			"and %d of %d rings of the FFT stage, scaled to the full pair; analysis Legendre on the 2lmax+2-ring CC grid as in ducc analysis_2d"
			% (stride, 100*info["m_fraction"], info["rings"], w["ny"]),
		"est_seconds_per_pair": est, "detail": info}

def run_reference(args, w):
	rank = int(os.environ.get("RANK", 0))
	if rank != 0: return
	use_all_host_cores()
	cb = None
	vals = []
	for i in range(args.warmup + args.steps):
		cb = cpu_baseline(w, budget_s=max(3.0, args.cpu_seconds/4))
		if i >= args.warmup: vals.append(cb["est_seconds_per_pair"])
	sec = float(np.mean(vals))
	out = {"impl": "reference", "metric": "SHT pairs/sec (map2alm+alm2map, 3-comp CAR) at lmax; %HBM roofline",
		"value": 1.0/sec, "unit": "pairs/s", "n_gpus": int(os.environ.get("WORLD_SIZE", 1)), "steps": args.steps, "warmup": args.warmup,
		"ms_per_step": sec*1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
		"config": {"workload": w["name"], "lmax": w["lmax"], "shape": [w["ncomp"], w["ny"], w["nx"]]},
		"cpu_baseline": dict(cb, value=1.0/sec),
		"e2e": {"value": 1.0/sec, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
		"gpu_launches": 0,
		"note": "reference arm = the repo's CPU oracle port on the host cores (pixell's own SHT arithmetic lives in ducc0, which is absent and cannot be installed offline; see DESIGN.md)"}
	print(json.dumps(out), flush=True)

def main():
	ap = argparse.ArgumentParser()
	ap.add_argument("--gpus", type=int, default=1)
	ap.add_argument("--steps", type=int, default=5)
	ap.add_argument("--warmup", type=int, default=3)
	ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
	ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
	ap.add_argument("--cpu-seconds", type=float, default=20.0)
	ap.add_argument("--no-cpu", action="store_true")
	args = ap.parse_args()
	w = WORKLOADS[args.workload]
	if args.impl == "reference": run_reference(args, w)
	else: run_ours(args, w)

if __name__ == "__main__":
	main()
