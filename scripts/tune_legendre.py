#!/usr/bin/env python
"""Time every launch-shape variant of the four Legendre kernels on one workload (tuning aid; run on the GPU box).
  python scripts/tune_legendre.py [c3|c2] [nvariants] [kernels e.g. 0123]
Prints one line per (kernel, variant): milliseconds (best of 3) and max abs difference to variant 0."""
import os, sys, ctypes
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pixell_b200 import _lib as L, sht

W = {"c3": (8192, 16384, 8000), "c2": (4608, 9216, 4096), "c1": (512, 1024, 256), "c0": (100, 200, 95)}

def main():
	wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
	nvar = int(sys.argv[2]) if len(sys.argv) > 2 else 5
	which = [int(c) for c in (sys.argv[3] if len(sys.argv) > 3 else "0123")]
	ny, nx, lmax = W[wl]
	L.init(0)
	lib = L.lib()
	plan = sht.plan_2d("F1", ny, nx, 0.0, lmax)
	nalm = (lmax+1)*(lmax+2)//2
	nring_pad = (ny+31)//32*32
	g = torch.Generator(device="cuda"); g.manual_seed(1)
	alm = torch.randn((2, nalm), dtype=torch.complex128, device="cuda", generator=g)
	alm[:, :lmax+1] = alm[:, :lmax+1].real.to(torch.complex128)
	leg = torch.zeros((2, lmax+1, nring_pad), dtype=torch.complex128, device="cuda")
	legin = torch.randn((2, lmax+1, nring_pad), dtype=torch.complex128, device="cuda", generator=g)
	out = torch.zeros_like(alm)
	st = torch.cuda.current_stream().cuda_stream
	names = ["synth0", "adj0", "synth2", "adj2"]
	for k in which:
		spin = 0 if k < 2 else 2
		ref = None
		for v in range(max(nvar, 1)):
			if nvar > 0 and lib.b2_set_leg_variant(k, v): break     # nvar = 0: keep the B2_LEG_VARIANT / default choice (profiling runs)
			def run():
				if k % 2 == 0: return lib.b2_alm2leg(plan.handle, spin, 0, alm.data_ptr(), nalm, leg.data_ptr(), st)
				return lib.b2_leg2alm(plan.handle, spin, 0, out.data_ptr(), nalm, legin.data_ptr(), st)
			if run():
				print("%s variant %d: %s" % (names[k], v, L.last_error())); continue
			torch.cuda.synchronize()
			best = 1e30
			for rep in range(3):
				e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
				e0.record(); run(); e1.record(); torch.cuda.synchronize()
				best = min(best, e0.elapsed_time(e1))
			res = (leg if k % 2 == 0 else out)[: (1 if spin == 0 else 2)].clone()
			(leg if k % 2 == 0 else out).fill_(float('nan'))      # a variant that skips outputs must not inherit them
			if ref is None: ref = res
			diff = (res-ref).abs().max().item()
			print("%-7s variant %d: %9.3f ms   maxdiff vs v0 %.3e" % (names[k], v, best, diff), flush=True)
		if nvar > 0: lib.b2_set_leg_variant(k, 0)

if __name__ == "__main__":
	main()
