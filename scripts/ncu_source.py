#!/usr/bin/env python
"""Per-region stall summary from an ncu report's SASS source page (needs ncu here, no GPU).
  python scripts/ncu_source.py report.ncu-rep [nregions]
Splits the kernel at branch/barrier instructions into regions, prints for the hottest regions: sample share,
instructions executed, instruction mix and the stall reasons."""
import csv, subprocess, sys, re, collections

def main():
	path = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 12
	out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
	allrows = list(csv.reader(out.splitlines()))
	heads = [i for i, r in enumerate(allrows) if r and r[0] == "Address"]
	seen = set()
	for n, hi in enumerate(heads):       # one section per captured launch (ncu repeats a launch's table per source view)
		end = heads[n+1]-1 if n+1 < len(heads) else len(allrows)
		name = allrows[hi-1][1] if hi > 0 and allrows[hi-1] and allrows[hi-1][0] == "Kernel Name" else ""
		key = (name, end-hi, tuple(allrows[hi+1][:1]) if hi+1 < end else ())
		if key in seen: continue
		seen.add(key)
		if name: print("==== %s" % name)
		section(allrows[hi], allrows[hi+1:end], ntop)

def section(H, data, ntop):
	c_src, c_smp, c_exec = H.index("Source"), H.index("# Samples"), H.index("Instructions Executed")
	stall_cols = [(i, h) for i, h in enumerate(H) if h.startswith("stall_") and "Not Issued" not in h]
	regions = []; cur = []
	for r in data:
		if len(r) < len(H): continue
		cur.append(r)
		if re.search(r"\b(BRA|EXIT|BAR|BSYNC|BSSY|WARPSYNC|RET)\b", r[c_src]): regions.append(cur); cur = []
	if cur: regions.append(cur)
	tot = sum(int(r[c_smp]) for r in data if len(r) >= len(H))
	totexec = sum(int(r[c_exec]) for r in data if len(r) >= len(H))
	print("total samples %d, warp instructions executed %d, regions %d" % (tot, totexec, len(regions)))
	def key(reg): return -sum(int(r[c_smp]) for r in reg)
	for reg in sorted(regions, key=key)[:ntop]:
		smp = sum(int(r[c_smp]) for r in reg); ex = sum(int(r[c_exec]) for r in reg)
		mix = collections.Counter()
		for r in reg:
			t = r[c_src].split()
			op = t[1] if t[0].startswith("@") else t[0]
			mix[op.split(".")[0]] += 1
		st = collections.Counter()
		for i, h in stall_cols:
			st[h[6:]] += sum(int(r[i]) for r in reg)
		print("\nregion %s..%s  %d instrs  samples %.1f%%  executed %.1f%%" % (reg[0][0][-5:], reg[-1][0][-5:], len(reg), 100.0*smp/tot, 100.0*ex/totexec))
		print("  mix:", ", ".join("%s %d" % kv for kv in mix.most_common(8)))
		print("  stalls:", ", ".join("%s %.1f%%" % (k, 100.0*v/max(smp, 1)) for k, v in st.most_common(7)))
		print("  ends:", reg[-1][c_src].strip())

if __name__ == "__main__":
	main()
