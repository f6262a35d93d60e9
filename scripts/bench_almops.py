#!/usr/bin/env python
"""Time the alm helpers (K6: alm2cl, lmul / almxfl, lmatmul, transpose_alm) at C3 size (lmax 8000, nalm 32 M, complex128)
with device-resident data and report achieved GB/s against the algorithmic bytes (alm2cl 16 nalm per alm
read, lmul 32 nalm) and the measured HBM peak.  python scripts/bench_almops.py [lmax]"""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pixell_b200 import curvedsky, _lib as L

def timeit(fn, reps=5):
	fn(); torch.cuda.synchronize()
	best = 1e30
	for _ in range(reps):
		e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
		e0.record(); fn(); e1.record(); torch.cuda.synchronize()
		best = min(best, e0.elapsed_time(e1))
	return best

def main():
	lmax = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
	L.init(0)
	ai = curvedsky.alm_info(lmax)
	n = ai.nelem
	g = torch.Generator(device="cuda"); g.manual_seed(1)
	alm = torch.randn((3, n), dtype=torch.complex128, device="cuda", generator=g)
	lf = torch.rand(lmax+1, dtype=torch.float64, device="cuda")
	lmat = torch.rand((3, 3, lmax+1), dtype=torch.float64, device="cuda")
	out = torch.empty_like(alm); one = alm[0].clone(); tr = torch.empty_like(one)
	try: peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
	except Exception: peak = 6650.0
	res = {}
	def rec(name, ms, nbytes): res[name] = {"ms": ms, "GB/s": nbytes/ms/1e6, "frac_hbm": nbytes/ms/1e6/peak, "algorithmic_bytes": nbytes}
	rec("alm2cl (auto, 1 alm)", timeit(lambda: ai.alm2cl(one)), 16*n)
	rec("alm2cl (cross, 2 alm)", timeit(lambda: ai.alm2cl(alm[0], alm[1])), 32*n)
	rec("lmul / almxfl (in place, 1 alm)", timeit(lambda: ai.lmul(one, lf, out=one)), 32*n)
	rec("lmatmul 3x3 (3 alm -> 3 alm)", timeit(lambda: ai.lmul(alm, lmat, out=out)), 2*3*16*n)
	rec("transpose_alm", timeit(lambda: ai.transpose_alm(one, tr)), 32*n)
	print(json.dumps({"workload": "alm helpers, lmax %d, nalm %d, complex128" % (lmax, n), "hbm_peak_gbs": peak, "kernels": res}))

if __name__ == "__main__":
	main()
