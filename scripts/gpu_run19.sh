mkdir -p gpurun_out
for kb in 100 210; do
  echo "== B2_RESAMP_SMEM_KB=$kb"
  B2_RESAMP_SMEM_KB=$kb python bench.py --steps 2 --warmup 2 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roundtrip_rel_err'], d['stage_ms_last_map2alm_group'])"
done
for kb in 100 200; do
  echo "== B2_FFT_SMEM_KB=$kb"
  B2_FFT_SMEM_KB=$kb python scripts/bench_fft.py 16384 32768 3 2 2>&1 | cut -c1-330
  B2_FFT_SMEM_KB=$kb python scripts/bench_fft.py 4096 8192 3 3 2>&1 | cut -c1-330
done
