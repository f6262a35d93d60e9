#!/usr/bin/env python
"""SASS evidence per kernel family: for every kernel of the built objects (pixell_b200/build/*.o, sm_100a) the count of the
mnemonics that show which hardware path it uses (TMA: UTMALDG / UTMASTG / UBLKCP / SYNCS = mbarrier; cp.async: LDGSTS;
FP64: DFMA / DADD / DMUL; shared memory, shuffles, barriers, atomics, fences).  Writes profiles/r2_sass_evidence.txt.
  python scripts/sass_evidence.py [output]"""
import os, re, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FAMILIES = [("K7t TMA-pipelined tile-DFT FFT", "tfft.o"), ("K1/K2 Legendre stage", "legendre.o"), ("K3/K4 ring FFTs", "ringfft.o"),
	("K5 theta weighting", "resample.o"), ("K7 axis passes", "fft2d.o"), ("K6 alm helpers + K9 rand_alm", "almops.o"), ("K8 arbitrary positions", "general.o"),
	("conversions / scaling", "api.o")]
KEYS = ["UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "UTMACMDFLUSH", "LDGSTS", "DFMA", "DADD", "DMUL", "DMMA", "LDS", "STS", "SHFL", "BAR", "ATOMG", "MEMBAR", "NANOSLEEP"]

def demangle(names):
	try:
		out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
		return out if len(out) == len(names) else names
	except Exception: return names

def main():
	dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_sass_evidence.txt")
	lines = ["# SASS evidence per kernel family (cuobjdump -sass of pixell_b200/build/*.o, sm_100a; regenerate with scripts/sass_evidence.py)", ""]
	for title, obj in FAMILIES:
		path = os.path.join(ROOT, "pixell_b200", "build", obj)
		if not os.path.exists(path): continue
		sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
		cur, counts = None, collections.OrderedDict()
		for l in sass.splitlines():
			m = re.search(r"Function : (\S+)", l)
			if m: cur = m.group(1); counts[cur] = collections.Counter(); continue
			m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
			if cur and m:
				op = m.group(1)
				for k in KEYS:
					if op == k or op.startswith(k+"."): counts[cur][k] += 1
		names = list(counts)
		lines.append("## %s (%s)" % (title, obj))
		for raw, nice in zip(names, demangle(names)):
			c = counts[raw]
			lines.append("  %-110s %s" % (nice[:110], " ".join("%s=%d" % (k, c[k]) for k in KEYS if c[k])))
		lines.append("")
	open(dst, "w").write("\n".join(lines))
	print("wrote", dst, len(lines), "lines")

if __name__ == "__main__": main()
