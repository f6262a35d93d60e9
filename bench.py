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
cpu_baseline  = pixell's call sequence on the host cores with a tuned CPU Legendre stage (oracle/sht_fast.c; ducc0
                cannot be installed here) + scipy.fft, on a bounded sample of the same workload
configs       = the other BASELINE.json configurations: c2 (T-only lmax 4096 pair), c4 (64 rand_map realisations
                sharded over the ranks, C_l broadcast over NCCL), c5 (rfft2 / irfft2 of 3 x 16384 x 32768)
parity_vs_oracle = the CUDA path against oracle/sht_oracle.c on an m comb at this very size (oracle/parity.py)
--impl reference times the CPU path alone (rank 0), same metric and config, plus complete unsampled pairs.
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

def operand_model(kernel):
	"""ceiling of the FP64 pipe for this kernel's main window from its SASS (register-file operand bandwidth, see
	scripts/sass_operand_model.py); None if profiles/operand_model.json has no entry"""
	try:
		with open(os.path.join(ROOT, "profiles", "operand_model.json")) as f: t = json.load(f)
		return t.get(kernel, {}).get("main_window_ceiling")
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
	nsteps_e2e = max(1, args.steps)
	t0 = time.perf_counter()
	for _ in range(nsteps_e2e): step_host()
	torch.cuda.synchronize()
	t_e2e = (time.perf_counter()-t0)/nsteps_e2e
	err_e2e = float(np.abs(nalm-alm0.cpu().numpy()).max()/alm0.abs().max().item())

	tt = torch.tensor([ms, t_e2e*1e3], dtype=torch.float64, device=device)
	if world > 1: dist.all_reduce(tt, op=dist.ReduceOp.MAX)
	ms, e2e_ms = tt[0].item(), tt[1].item()
	del hmap, halm, nmap, nalm
	# ---- the other BASELINE.json configurations as sub-records (C4 shards over the ranks; C2 / C5 are single-GPU)
	configs = {}
	if not args.no_configs:
		configs["c4"] = bench_c4(torch, dist, device, rank, world)
		if world == 1:
			configs["c2"] = bench_c2(torch, device)
			del map, alm, alm0
			sht.clear_plans(); torch.cuda.empty_cache()
			configs["c5"] = bench_c5(torch, device)
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
				"flops_model": "canonical, SURVEY.md 8d (no credit for polar skipping)",
				"operand_bandwidth_ceiling": operand_model("k_adj2" if w["ncomp"] == 3 else "k_adj0"),
				"operand_bandwidth_note": "a DFMA that has to fetch three register operands issues every 3 cycles, not 2 (measured: 24.6 vs 36.9 TFLOP/s); the ceiling is the fraction of the DFMA peak the kernel's own instruction stream can reach (scripts/sass_operand_model.py)"},
			"stage_ms_last_map2alm_group": tim,
			"clocks": clocks,
			"configs": configs,
		}
		if world == 1 and not args.no_cpu:
			# the checker legs (after every timed region): CPU baseline, and parity against the oracle at this size
			out["cpu_baseline"] = cpu_baseline(w, budget_s=args.cpu_seconds)
			if not args.no_parity:
				from oracle import parity
				out["parity_vs_oracle"] = parity.baseline_parity(w)
				if "c2" in configs: configs["c2"]["parity_vs_oracle"] = parity.baseline_parity(WORKLOADS["c2"])
		print(json.dumps(out), flush=True)
	if world > 1: dist.destroy_process_group()

# ---------------------------------------------------------------------------------------------
# the other configurations of BASELINE.json

def bench_c2(torch, device, steps=10):
	"""C2: T-only full-sky lmax 4096 pair on one GPU"""
	from pixell_b200 import curvedsky, sht
	w = WORKLOADS["c2"]
	alm0, map, wcs, ainfo, spin = make_inputs(w, 2, torch, device)
	alm = torch.zeros_like(alm0)
	def step():
		curvedsky.map2alm(map, alm, spin=spin, wcs=wcs, ainfo=ainfo)
		curvedsky.alm2map(alm, map, spin=spin, wcs=wcs, ainfo=ainfo)
	for _ in range(3): step()
	err = ((alm-alm0).abs().max()/alm0.abs().max()).item()
	torch.cuda.synchronize()
	e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	e0.record()
	for _ in range(steps): step()
	e1.record(); torch.cuda.synchronize()
	ms = e0.elapsed_time(e1)/steps
	curvedsky.map2alm(map, alm, spin=spin, wcs=wcs, ainfo=ainfo); torch.cuda.synchronize()
	tim = next(reversed(sht._plans.values())).last_timing()
	pk, _ = peaks()
	return {"workload": w["name"], "value": 1e3/ms, "unit": "pairs/s", "ms_per_pair": ms, "roundtrip_rel_err": err,
		"frac_hbm_pair": algorithmic_bytes(w)/(ms*1e-3)/1e9/pk["hbm_gbs"], "stage_ms_map2alm": tim,
		"fp64_tflops_legendre_adjoint": canonical_flops(w, 1, 0)/(tim["legendre"]*1e-3)/1e12}

def bench_c4(torch, dist, device, rank, world, nsim=64):
	"""C4: 64 (T,Q,U) rand_map realisations from C_l at lmax 4096 on the 4608 x 9216 grid, block-partitioned over the
	ranks; the input C_l is broadcast from rank 0 (NCCL) inside the timed region; maps stay on their GPU"""
	from pixell_b200 import mc, geometry
	lmax, ny, nx = 4096, 4608, 9216
	ps = None
	if rank == 0:
		l = np.arange(lmax+1.0); tt = np.where(l >= 2, 1.0/np.maximum(l*(l+1), 1), 0.0)
		ps = np.zeros((3, 3, lmax+1)); ps[0, 0] = tt; ps[1, 1] = 0.3*tt; ps[2, 2] = 0.1*tt; ps[0, 1] = ps[1, 0] = 0.5*np.sqrt(ps[0, 0]*ps[1, 1])
	shape, wcs = geometry.fullsky_geometry(shape=(ny, nx))
	seeds = list(range(1000, 1000+nsim))
	mine = mc.partition(len(seeds))
	out = torch.empty((max(1, len(mine)), 3)+shape, dtype=torch.float64, device=device)
	ps_w = mc.broadcast_ps(ps, src=0)
	mc.rand_maps((3,)+shape, wcs, ps_w, seeds[:world], lmax=lmax, rng="device", out=out[:1])      # warm-up: plans, tables
	torch.cuda.synchronize()
	if world > 1: dist.barrier()
	e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	e0.record()
	ps_w = mc.broadcast_ps(ps, src=0)
	mc.rand_maps((3,)+shape, wcs, ps_w, seeds, lmax=lmax, rng="device", out=out)
	e1.record(); torch.cuda.synchronize()
	ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
	if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
	var_t = float(out[:, 0].var().item())
	want = float(np.sum((2*np.arange(lmax+1)+1)*(np.where(np.arange(lmax+1) >= 2, 1.0/np.maximum(np.arange(lmax+1.0)*(np.arange(lmax+1.0)+1), 1), 0.0)))/(4*np.pi))
	pk, _ = peaks()
	rate = nsim/(ms.item()*1e-3)
	del out
	return {"workload": "C4: %d (T,Q,U) rand_map realisations (device Philox stream b2_rand_alm + alm2map), lmax %d, %dx%d, %d GPU(s), C_l broadcast from rank 0" % (nsim, lmax, ny, nx, world),
		"value": rate, "unit": "realisations/s", "ms_total": ms.item(), "n_gpus": world, "scaling": "strong (64 realisations over the ranks)",
		"bytes_written_per_realisation": 8*3*ny*nx, "frac_hbm_map_write": rate*8*3*ny*nx/1e9/pk["hbm_gbs"]/world,
		"var_T_rank0": var_t, "var_T_expected": want}

def bench_c5(torch, device, reps=3):
	"""C5: rfft2 -> Gaussian filter -> irfft2 on a 3 x 16384 x 32768 float64 patch (pixell_b200.fft, the TMA path)"""
	from pixell_b200 import fft as F
	nc, ny, nx = 3, 16384, 32768
	g = torch.Generator(device=device); g.manual_seed(5)
	m = torch.randn((nc, ny, nx), dtype=torch.float64, device=device, generator=g)
	ft = torch.empty((nc, ny, nx//2+1), dtype=torch.complex128, device=device)
	out = torch.empty_like(m)
	res = np.deg2rad(0.5/60); sigma = np.deg2rad(1.4/60)/np.sqrt(8*np.log(2))
	ly = np.fft.fftfreq(ny, res)*2*np.pi; lx = np.fft.rfftfreq(nx, res)*2*np.pi
	fy = torch.as_tensor(np.exp(-0.5*sigma**2*ly**2), device=device)[:, None]
	fx = torch.as_tensor(np.exp(-0.5*sigma**2*lx**2), device=device)[None, :]
	def ev(): e = torch.cuda.Event(enable_timing=True); e.record(); return e
	best = None
	for rep in range(reps+1):
		e0 = ev(); F.rfft(m, ft, axes=[-2, -1]); e1 = ev()
		F.fourier_filter(ft, fy=fy[:, 0], fx=fx[0]); e2 = ev()
		F.irfft(ft, out, n=nx, axes=[-2, -1], normalize=True); e3 = ev()
		torch.cuda.synchronize()
		t = (e0.elapsed_time(e1), e1.elapsed_time(e2), e2.elapsed_time(e3))
		if rep > 0 and (best is None or sum(t) < sum(best)): best = t
	# size-independent checks: the filter leaves the mean untouched; without the filter the pair is the identity
	err_mean = float(((out.mean(dim=(-2, -1)) - m.mean(dim=(-2, -1))).abs().max()).item())
	F.rfft(m, ft, axes=[-2, -1]); F.irfft(ft, out, n=nx, axes=[-2, -1], normalize=True)
	err_rt = float((out-m).abs().max().item())
	nbytes = 8*nc*ny*nx + 16*nc*ny*(nx//2+1)
	pk, _ = peaks()
	del m, ft, out
	return {"workload": "C5: rfft2 -> Gaussian filter -> irfft2, %dx%dx%d f64" % (nc, ny, nx), "ms_rfft2": best[0], "ms_filter": best[1], "ms_irfft2": best[2],
		"value": 1e3/(best[0]+best[2]), "unit": "rfft2+irfft2 pairs/s",
		"frac_hbm_rfft2": nbytes/best[0]/1e6/pk["hbm_gbs"], "frac_hbm_irfft2": nbytes/best[2]/1e6/pk["hbm_gbs"], "algorithmic_bytes_per_transform": nbytes,
		"mean_preserved_abs_err": err_mean, "roundtrip_abs_err": err_rt}

# ---------------------------------------------------------------------------------------------
# CPU arm: "pixell + ducc0 on the host cores".  ducc0 is not installable here (tried first, below); the stand-in is
# oracle/sht_fast.c (OpenMP over m, SIMD over ring blocks, ring skipping: the structure of libsharp / ducc, compiled
# -march=native on the box it runs on) + scipy.fft (= the pocketfft ducc ships) for the ring FFTs and the theta
# resampling of analysis_2d, called in the order pixell's curvedsky calls ducc (pixell/curvedsky.py:900-962, 1018-1046).

def use_all_host_cores():
	"""torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core.  Must run before the
	oracle's OpenMP runtime is loaded."""
	n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
	os.environ["OMP_NUM_THREADS"] = str(n)
	return n

def try_ducc0():
	try:
		import ducc0
		return ducc0
	except Exception:
		return None

def cpu_groups(w): return [(0, 1)] + ([(2, 2)] if w["ncomp"] == 3 else [])

def cpu_pair_full(w, seed=0):
	"""One complete, unsampled map2alm + alm2map pair on the host cores; returns (seconds, round-trip error)."""
	from oracle import sht_oracle as so
	lmax, ny, nx = w["lmax"], w["ny"], w["nx"]
	theta = so.grid_theta("F1", ny)
	mstart = so.default_mstart(lmax, lmax); nalm = (lmax+1)*(lmax+2)//2
	nt = so._good_cc_size(2*lmax+2); theta_cc = so.grid_theta("CC", nt); wcc = so.get_gridweights("CC", nt)/nx
	rng = np.random.default_rng(seed)
	total = 0.0; err = 0.0
	for spin, nc in cpu_groups(w):
		alm = np.zeros((nc, nalm), complex)
		l = np.concatenate([np.arange(m, lmax+1) for m in range(lmax+1)]) if lmax <= 300 else None
		alm.real = rng.standard_normal((nc, nalm)); alm.imag = rng.standard_normal((nc, nalm))
		alm[:, :lmax+1] = alm[:, :lmax+1].real
		for m in range(min(spin, lmax+1)): alm[:, mstart[m]+np.arange(m, spin)] = 0
		# the map to analyse (not timed): synthesis of the band-limited alm
		m0 = so.leg2map_t(so.fast_alm2leg(alm, theta, spin, lmax, lmax, mstart), nx, 0.1)
		t0 = time.perf_counter()
		# map2alm: analysis_2d = ring FFTs, theta resampling to the 2 lmax + 2 ring Clenshaw-Curtis grid, weights, Legendre adjoint
		leg = so.map2leg_t(m0, lmax+1, 0.1)                                                # [nc, m, ring]
		leg = so.resample_to_cc_t(leg, "F1", nt, spin)
		leg *= wcc[None, None, :]
		back = so.fast_leg2alm(leg, theta_cc, spin, lmax, lmax, mstart, nalm)
		# alm2map: synthesis_2d
		leg = so.fast_alm2leg(back, theta, spin, lmax, lmax, mstart)
		m1 = so.leg2map_t(leg, nx, 0.1)
		total += time.perf_counter()-t0
		err = max(err, float(np.abs(back-alm).max()/np.abs(alm).max()), float(np.abs(m1-m0).max()/np.abs(m0).max()))
		del leg, m0, m1, back, alm
	return total, err

def cpu_pair_sample(w, stride, ring_frac=0.02, seed=0):
	"""A bounded sample of the same pair: every `stride`-th m (offset stride/2, all l, all rings) in the Legendre and
	resampling stages, a subset of rings in the ring-FFT stage; returns (estimated full-pair seconds, sample seconds, detail).
	A regular comb samples the m distribution uniformly, so the scaling is by counts."""
	from oracle import sht_oracle as so
	lmax, ny, nx = w["lmax"], w["ny"], w["nx"]
	theta = so.grid_theta("F1", ny)
	mstart = so.default_mstart(lmax, lmax); nalm = (lmax+1)*(lmax+2)//2
	nt = so._good_cc_size(2*lmax+2); theta_cc = so.grid_theta("CC", nt)
	ms = np.arange(stride//2, lmax+1, stride, dtype=np.int32)
	scale_m = (lmax+1.0)/len(ms)
	rng = np.random.default_rng(seed)
	t_leg = t_res = t_fft = 0.0
	nr = max(16*len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else 64, int(ny*ring_frac))      # every core gets rings
	nr = min(nr, ny)
	for spin, nc in cpu_groups(w):
		alm = np.zeros((nc, nalm), complex)
		for m in ms:
			l = np.arange(max(m, spin), lmax+1)
			alm[:, mstart[m]+l] = rng.standard_normal((nc, len(l))) + 1j*rng.standard_normal((nc, len(l)))
		t0 = time.perf_counter()
		leg = so.fast_alm2leg(alm, theta, spin, lmax, lmax, mstart, ms)                      # synthesis Legendre, ny rings
		t1 = time.perf_counter()
		legcc = so.resample_to_cc_t(leg, "F1", nt, spin)                                        # [nc, len(ms), nt]
		t2 = time.perf_counter()
		t3 = t2
		so.fast_leg2alm(legcc, theta_cc, spin, lmax, lmax, mstart, nalm, ms)                  # analysis Legendre, 2 lmax + 2 rings
		t4 = time.perf_counter()
		t_leg += (t1-t0) + (t4-t3); t_res += t2-t1
		mp = rng.standard_normal((nc, nr, nx))
		t5 = time.perf_counter()
		lg = so.map2leg_t(mp, lmax+1, 0.1)                                                    # [nc, m, ring] for the Legendre stage
		so.leg2map_t(lg, nx, 0.1)
		t_fft += time.perf_counter()-t5
		del alm, leg, legcc, mp, lg
	est = (t_leg + t_res)*scale_m + t_fft*ny/nr
	return est, t_leg + t_res + t_fft, dict(t_legendre_est=t_leg*scale_m, t_resample_est=t_res*scale_m, t_ringfft_est=t_fft*ny/nr,
		m_sampled=int(len(ms)), m_stride=int(stride), rings_fft=int(nr))

def pick_stride(w, cores, seconds):
	"""comb stride for a sample of about `seconds`: at least 8 m values per core so that every core stays busy"""
	from oracle import sht_oracle as so
	probe = max(1, (w["lmax"]+1)//(8*cores))
	t0 = time.perf_counter(); cpu_pair_sample(w, probe, 0.005); tp = time.perf_counter()-t0
	return max(1, min(probe, int(round(probe*tp/max(seconds, 1e-3))))), tp

def cpu_baseline(w, budget_s=20.0):
	use_all_host_cores()
	from oracle import sht_oracle as so
	so.build_fast()
	cores = so.fast_nthreads()
	stride, tp = pick_stride(w, cores, budget_s)
	est, tsample, info = cpu_pair_sample(w, stride, 0.02)
	return {"value": 1.0/est, "unit": "pairs/s", "cores": cores, "kind": "port",
		"sample": "tuned CPU restatement (oracle/sht_fast.c, libsharp/ducc structure, -march=native; NOT ducc0, which is not installable here) "
			"on every %d-th m of the Legendre and theta-resampling stages (%d of %d m, all l, all rings; analysis Legendre on the 2lmax+2-ring "
			"CC grid as ducc analysis_2d does) and %d of %d rings of the ring-FFT stage (scipy.fft = pocketfft), scaled by counts to the full pair"
			% (stride, info["m_sampled"], w["lmax"]+1, info["rings_fft"], w["ny"]),
		"est_seconds_per_pair": est, "sample_seconds": tsample, "extrapolated": True, "detail": info}

def run_reference(args, w):
	rank = int(os.environ.get("RANK", 0))
	if rank != 0: return
	cores = use_all_host_cores()
	t_start = time.perf_counter()
	ducc = try_ducc0()
	from oracle import sht_oracle as so
	so.build_fast()
	cores = so.fast_nthreads()
	# every step is a bounded sample of the workload (about 5 s of CPU work); the whole run stays within a few minutes
	per_step = max(2.0, min(6.0, 150.0/max(1, args.warmup+args.steps)))
	stride, _ = pick_stride(w, cores, per_step)
	vals, samples = [], []
	info = None
	for i in range(args.warmup + args.steps):
		est, ts, info = cpu_pair_sample(w, stride, 0.01, seed=i)
		if i >= args.warmup: vals.append(est); samples.append(ts)
	sec = float(np.mean(vals))
	# one complete, unsampled pair at the largest configuration that fits the run: measured, and the scaling law checked on it
	measured = {}
	c2 = WORKLOADS["c2"]
	t_full, err_full = cpu_pair_full(c2)
	est2, _, _ = cpu_pair_sample(c2, max(1, (c2["lmax"]+1)//(32*cores)), 0.05)
	measured["c2_full_pair_seconds"] = t_full; measured["c2_roundtrip_rel_err"] = err_full
	measured["c2_extrapolated_seconds"] = est2; measured["c2_extrapolated_over_measured"] = est2/t_full
	if w is not c2 and sec < 60.0 and (time.perf_counter()-t_start) + 1.3*sec < 280.0:
		t3, e3 = cpu_pair_full(w)
		measured["full_pair_seconds"] = t3; measured["roundtrip_rel_err"] = e3; measured["extrapolated_over_measured"] = sec/t3
	out = {"impl": "reference", "metric": "SHT pairs/sec (map2alm+alm2map, 3-comp CAR) at lmax; %HBM roofline",
		"value": 1.0/sec, "unit": "pairs/s", "n_gpus": int(os.environ.get("WORLD_SIZE", 1)), "steps": args.steps, "warmup": args.warmup,
		"ms_per_step": sec*1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
		"config": {"workload": w["name"], "lmax": w["lmax"], "shape": [w["ncomp"], w["ny"], w["nx"]]},
		"extrapolated": True, "sample_seconds_per_step": float(np.mean(samples)),
		"cpu_baseline": {"value": 1.0/sec, "unit": "pairs/s", "cores": cores, "kind": "port", "extrapolated": True,
			"sample": "each step: every %d-th m (%d of %d, all l, all rings) of the Legendre + theta-resampling stages and %d of %d rings of the ring FFTs, "
				"scaled by counts; ms_per_step is the estimated FULL pair, sample_seconds_per_step the CPU time actually spent"
				% (stride, info["m_sampled"], w["lmax"]+1, info["rings_fft"], w["ny"]),
			"est_seconds_per_pair": sec, "detail": info, "measured": measured,
			"ducc0": "import ducc0 failed (not installable offline): tuned restatement oracle/sht_fast.c used" if ducc is None else "ducc0 %s present" % getattr(ducc, "__version__", "?")},
		"e2e": {"value": 1.0/sec, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
		"gpu_launches": 0,
		"note": "reference arm = pixell's call sequence on the host cores with a tuned CPU Legendre stage (oracle/sht_fast.c) + scipy.fft standing in for ducc0 "
			"(absent, cannot be installed offline; see DESIGN.md).  `measured` holds complete unsampled pairs timed in this run."}
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
	ap.add_argument("--no-configs", action="store_true", help="skip the C2 / C4 / C5 sub-records")
	ap.add_argument("--no-parity", action="store_true", help="skip the oracle comb parity check at the bench size")
	args = ap.parse_args()
	w = WORKLOADS[args.workload]
	if args.impl == "reference": run_reference(args, w)
	else: run_ours(args, w)

if __name__ == "__main__":
	main()
