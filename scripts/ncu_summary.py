#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small text files for profiles/.
  python scripts/ncu_summary.py launches gpurun_out/launches_c3.csv          -> per-kernel time shares
  python scripts/ncu_summary.py full gpurun_out/prof_leg2.ncu-rep            -> key metrics per captured launch
"""
import csv, collections, subprocess, sys

def launches(path):
	rows = [r for r in csv.reader(open(path)) if len(r) > 5]
	hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
	H = rows[hdr]; data = rows[hdr+1:]
	ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
	agg = collections.OrderedDict()
	for r in data:
		name = r[ki].split("(")[0]; v = float(r[vi].replace(",", "")); u = r[ui]
		v *= {"nsecond": 1e-6, "ns": 1e-6, "usecond": 1e-3, "us": 1e-3, "msecond": 1.0, "ms": 1.0, "second": 1e3, "s": 1e3}.get(u, 1.0)
		a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
	tot = sum(v[1] for v in agg.values())
	print("# gpu__time_duration.sum per kernel (ncu --clock-control none; cold-cache, serialised: compare SHARES)")
	print("%-64s %6s %12s %7s %12s" % ("kernel", "n", "total ms", "share", "ms/launch"))
	for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
		print("%-64s %6d %12.3f %6.1f%% %12.3f" % (k[:64], v[0], v[1], 100*v[1]/tot, v[1]/v[0]))
	print("%-64s %6s %12.3f" % ("total", "", tot))

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
	"launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
	"sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
	"smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
	"gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
	"lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
	"smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
	"smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
	"smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
	"smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
	"smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__inst_executed.sum"]

def full(path):
	out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
	rows = list(csv.reader(out.splitlines()))
	H, U = rows[0], rows[1]
	for r in rows[2:]:
		print("==== %s" % r[H.index("Kernel Name")])
		for k in KEYS:
			if k in H:
				i = H.index(k); print("%-90s %18s %s" % (k, r[i], U[i]))

if __name__ == "__main__":
	{"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
