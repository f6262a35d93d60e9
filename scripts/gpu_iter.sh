# scratch script of the current experiment: rewritten before every `gpurun -- 'bash scripts/gpu_iter.sh <tag>'`
set -x
mkdir -p gpurun_out
T=${1:-it}
for st in 1 2; do B2_SIG_STYLE=$st B2_TRACE=1 python scripts/e2e_probe.py 2> gpurun_out/${T}_trace.txt | cut -c1-200; grep -n "map -> alm" -A20 gpurun_out/${T}_trace.txt | tail -21 | grep "g1 K5 done\|g1 K2 done"; done
