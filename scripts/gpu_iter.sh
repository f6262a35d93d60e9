set -x
mkdir -p gpurun_out
T=${1:-it}
python -m pytest tests/test_mc_gpu.py -x -q -m gpu 2>&1 | tail -15
python - <<'PY' 2>&1 | tail -8
import sys, json; sys.path.insert(0, '.')
import torch, bench
from pixell_b200 import _lib as L
L.init(0)
class D: pass
for b in (1, 4):
    import pixell_b200.mc as mc
    orig = mc.rand_maps
    def patched(*a, **k): k.setdefault("batch", b); return orig(*a, **k)
    mc.rand_maps = patched
    r = bench.bench_c4(torch, None, torch.device("cuda", 0), 0, 1)
    mc.rand_maps = orig
    print("batch", b, r["value"], "realisations/s", r["ms_total"], "var", r["var_T_rank0"])
PY
