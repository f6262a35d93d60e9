set -x
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__grid_size,launch__block_size,launch__cluster_dim_x
for mode in 0 1; do
for sz in "4096 8192 1 1" "16384 32768 1 1"; do
B2_FFT_LEGACY=$mode ncu --metrics $M --clock-control none -k regex:k_fft -s 4 -c 4 --csv --log-file /tmp/l.csv python scripts/bench_fft.py $sz > /dev/null 2>&1
echo "== nocluster=$mode size=$sz"; python - <<'PY'
import csv
rows=[r for r in csv.reader(l for l in open('/tmp/l.csv') if l.startswith('"'))]
h=rows[0]; iK=h.index("Kernel Name"); iM=h.index("Metric Name"); iV=h.index("Metric Value"); iI=h.index("ID")
d={}
for r in rows[1:]:
    d.setdefault((r[iI],r[iK]),{})[r[iM]]=r[iV]
for (i,k),m in d.items():
    print(k[:40], " ".join("%s=%s"%(a.split("__")[-1][:22],b) for a,b in m.items()))
PY
done; done > gpurun_out/fft_cl_launches.txt 2>&1
cat gpurun_out/fft_cl_launches.txt
