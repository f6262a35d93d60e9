"""where does the host-memory (e2e) time go: curvedsky API vs the engine call, per direction"""
import os, sys, time, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pixell_b200 import curvedsky, geometry, sht, _lib as L
L.init(0)
w = bench.WORKLOADS["c3"]
dev = torch.device("cuda", 0)
alm0, map, wcs, ainfo, spin = bench.make_inputs(w, 3, torch, dev)
hmap = torch.empty(map.shape, dtype=torch.float64, pin_memory=True); hmap.copy_(map)
halm = torch.empty(alm0.shape, dtype=torch.complex128, pin_memory=True)
nmap = geometry.ndmap(hmap.numpy(), wcs); nalm = halm.numpy()
def t(fn, n=3):
	fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
	for _ in range(n): fn()
	torch.cuda.synchronize(); return (time.perf_counter()-t0)/n*1e3
alm = torch.zeros_like(alm0)
out = {}
out["dev_map2alm_ms"] = t(lambda: curvedsky.map2alm(map, alm, spin=spin, wcs=wcs, ainfo=ainfo))
out["dev_alm2map_ms"] = t(lambda: curvedsky.alm2map(alm, map, spin=spin, wcs=wcs, ainfo=ainfo))
out["host_map2alm_ms"] = t(lambda: curvedsky.map2alm(nmap, nalm, spin=spin, ainfo=ainfo))
out["host_alm2map_ms"] = t(lambda: curvedsky.alm2map(nalm, nmap, spin=spin, ainfo=ainfo))
# raw copies of the same volumes
d = torch.empty_like(map); a = torch.empty_like(alm0)
out["h2d_map_ms"] = t(lambda: d.copy_(hmap, non_blocking=True)); out["d2h_map_ms"] = t(lambda: hmap.copy_(d, non_blocking=True))
out["h2d_alm_ms"] = t(lambda: a.copy_(halm, non_blocking=True)); out["d2h_alm_ms"] = t(lambda: halm.copy_(a, non_blocking=True))
print(json.dumps(out))
