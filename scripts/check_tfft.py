#!/usr/bin/env python
"""Quick GPU check of the TMA FFT path (tfft.cu) against numpy's pocketfft, plus timings.  python scripts/check_tfft.py [time]"""
import os, sys, time, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pixell_b200 import fft as F, _lib as L

def rel(a, b): return float(np.abs(a-b).max()/max(np.abs(b).max(), 1e-300))
L.init(0)
rng = np.random.default_rng(0)
bad = 0
for shape in [(1, 64, 64), (2, 256, 512), (1, 1280, 5000), (1, 4096, 4096), (3, 96, 160), (1, 2048, 24*16)]:
	a = rng.standard_normal(shape) + 1j*rng.standard_normal(shape)
	want = np.fft.fft2(a)
	got = F.fft(a, axes=[-2, -1]); e1 = rel(got, want)
	back = F.ifft(got, axes=[-2, -1], normalize=True); e2 = rel(back, a)
	b = a.copy(); F.fft(b, b, axes=[-2, -1]); e3 = rel(b, want)
	print("c2c", shape, "fwd %.1e inv %.1e inplace %.1e" % (e1, e2, e3), flush=True)
	bad += (max(e1, e2, e3) > 1e-12)
for shape in [(1, 64, 128), (2, 256, 512), (2, 1536, 16384), (3, 4096, 8192), (1, 1280, 6400), (1, 2048, 10000)]:
	a = rng.standard_normal(shape)
	want = np.fft.rfft2(a)
	got = F.rfft(a, axes=[-2, -1]); e1 = rel(got, want)
	back = F.irfft(got, n=shape[-1], axes=[-2, -1], normalize=True); e2 = rel(back, a)
	print("r2c", shape, "fwd %.1e inv %.1e" % (e1, e2), flush=True)
	bad += (max(e1, e2) > 1e-12)
print("FAILURES:", bad)
if len(sys.argv) > 1:
	def ev(): e = torch.cuda.Event(enable_timing=True); e.record(); return e
	for (nc, ny, nx) in [(3, 4096, 8192), (3, 16384, 32768)]:
		g = torch.Generator(device="cuda"); g.manual_seed(5)
		m = torch.randn((nc, ny, nx), dtype=torch.float64, device="cuda", generator=g)
		ft = torch.empty((nc, ny, nx//2+1), dtype=torch.complex128, device="cuda")
		out = torch.empty_like(m)
		best = [1e9, 1e9]
		for rep in range(4):
			e0 = ev(); F.rfft(m, ft, axes=[-2, -1]); e1 = ev(); F.irfft(ft, out, n=nx, axes=[-2, -1], normalize=True); e2 = ev()
			torch.cuda.synchronize()
			if rep: best = [min(best[0], e0.elapsed_time(e1)), min(best[1], e1.elapsed_time(e2))]
		err = float((out-m).abs().max().item())
		nbytes = 8*nc*ny*nx + 16*nc*ny*(nx//2+1)
		print(json.dumps({"shape": [nc, ny, nx], "ms_rfft2": best[0], "ms_irfft2": best[1], "frac_hbm_rfft2": nbytes/best[0]/1e6/6546.9,
			"frac_hbm_irfft2": nbytes/best[1]/1e6/6546.9, "roundtrip_abs_err": err, "tma": os.environ.get("B2_FFT_NO_TMA", "0") != "1",
			"slab_mb": os.environ.get("B2_TFFT_SLAB_MB", "24")}), flush=True)
		del m, ft, out
