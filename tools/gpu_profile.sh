#!/bin/bash
# Runs on the GPU box (under gpurun): parity tests, the bench line, the ncu launch list of the same bench
# command and one full ncu capture per kernel.  Usage: tools/gpu_profile.sh <tag> [--notest] [kernel-regex ...]
# Everything lands in gpurun_out/<tag>_*.
set -u
TAG=${1:-rX}; shift || true
NOTEST=0
if [ "${1:-}" = "--notest" ]; then NOTEST=1; shift; fi
KERNELS=${@:-"fused_outer_kernel alpha_inner_kernel"}
OUT=gpurun_out
mkdir -p $OUT
if [ $NOTEST -eq 0 ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest=$?"
  tail -3 $OUT/${TAG}_pytest_gpu.log
fi
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench=$?"
cat $OUT/${TAG}_bench.json
PROF="python bench.py --steps 1 --warmup 1 --outer 10 --profile"
timeout 300 $PROF > $OUT/${TAG}_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv $PROF > $OUT/${TAG}_ncu0.log 2>&1
echo "launchlist=$?"
# gpurun brings back at most 64 MiB and one --set full report is ~18 MB: every report is reduced to its raw-page CSV on the
# box; only the first kernel keeps the .ncu-rep (with sources, for the source page)
n=0
for k in $KERNELS; do
  SRC="--import-source on"; [ $n -ge 1 ] && SRC=""
  timeout 300 $PROF > $OUT/${TAG}_plain.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none $SRC -k regex:$k -s 3 -c 1 -f -o $OUT/${TAG}_$k $PROF > $OUT/${TAG}_ncu_$k.log 2>&1
  echo "ncu $k=$?"
  ncu -i $OUT/${TAG}_$k.ncu-rep --page raw --csv > $OUT/${TAG}_${k}_raw.csv 2>/dev/null
  [ $n -ge 1 ] && rm -f $OUT/${TAG}_$k.ncu-rep
  n=$((n+1))
done
python tools/ncu_traffic.py $TAG $KERNELS > $OUT/${TAG}_ncu_traffic.json 2>/dev/null; echo "traffic=$?"
du -sh $OUT
