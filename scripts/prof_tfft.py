import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixell_b200 import fft as F, _lib as L
L.init(0)
nc, ny, nx = 3, 4096, 8192
m = torch.randn((nc, ny, nx), dtype=torch.float64, device="cuda")
ft = torch.empty((nc, ny, nx//2+1), dtype=torch.complex128, device="cuda")
for i in range(2): F.rfft(m, ft, axes=[-2, -1])
torch.cuda.synchronize()
