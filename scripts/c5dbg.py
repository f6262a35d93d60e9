import os, sys, torch
sys.path.insert(0, '/root/repo')
from pixell_b200 import fft as F, _lib as L
L.init(0)
nc, ny, nx = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
order = sys.argv[4] if len(sys.argv) > 4 else "fi"
m = torch.randn((nc, ny, nx), dtype=torch.float64, device="cuda")
ft = torch.randn((nc, ny, nx//2+1), dtype=torch.complex128, device="cuda")
try:
    for ch in order:
        if ch == "f": F.rfft(m, ft, axes=[-2, -1]); torch.cuda.synchronize(); print("rfft ok")
        else: F.irfft(ft, m, n=nx, axes=[-2, -1]); torch.cuda.synchronize(); print("irfft ok")
except Exception as e: print("ERR", e)
