#!/usr/bin/env python
"""Instruction mix, stall totals and hottest instructions of one launch from `ncu -i X.ncu-rep --page source --csv` output.
  ncu -i rep --page source --csv --launch-skip K --launch-count 1 > src.csv ; python scripts/ncu_hot.py src.csv [ntop]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
H = rows[1]
iS, iE, iSm = H.index('Source'), H.index('Instructions Executed'), H.index('# Samples')
data = [r for r in rows[2:] if len(r) == len(H) and r[iE].isdigit()]
ops = collections.Counter(); samp = collections.Counter(); tot = tots = 0
for r in data:
	toks = r[iS].split()
	op = (toks[1] if toks[0].startswith('@') else toks[0]).split('.')[0]
	e, s = int(r[iE]), int(r[iSm])
	ops[op] += e; samp[op] += s; tot += e; tots += s
print("kernel:", rows[0][1], "| warp instructions", tot, "| samples", tots)
for op, c in ops.most_common(22): print("%-10s %11d %5.1f%%   samples %5.1f%%" % (op, c, 100*c/tot, 100*samp[op]/max(tots, 1)))
for name in ['stall_barrier', 'stall_wait', 'stall_short_sb', 'stall_long_sb', 'stall_math', 'stall_mio', 'stall_not_selected', 'stall_selected',
		'stall_sleep', 'stall_branch_resolving', 'stall_no_inst', 'stall_dispatch', 'stall_lg', 'stall_membar']:
	i = H.index(name); print("%-24s %8d" % (name, sum(int(r[i]) for r in data)))
data.sort(key=lambda r: -int(r[iSm]))
for r in data[:ntop]: print("%6s %9s  %s" % (r[iSm], r[iE], r[iS][:100]))
