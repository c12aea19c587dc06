#!/bin/bash
# Runs on the GPU box (under gpurun): parity tests, the bench line, the ncu launch list of the same bench
# command and one full ncu capture per streaming kernel.  Usage: tools/gpu_profile.sh <tag> [kernel-regex ...]
# Everything lands in gpurun_out/<tag>_*.
set -u
TAG=${1:-rX}; shift || true
KERNELS=${@:-"rowgram4_kernel gram_panel_kernel u_inner_kernel alpha_inner_kernel"}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest=$?"
tail -3 $OUT/${TAG}_pytest_gpu.log
python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench=$?"
cat $OUT/${TAG}_bench.json
PROF="python bench.py --steps 1 --warmup 1 --outer 10 --profile"
$PROF > $OUT/${TAG}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv $PROF > $OUT/${TAG}_ncu0.log 2>&1
echo "launchlist=$?"
# gpurun brings back at most 64 MiB and one --set full report is ~18 MB: every report is reduced to its raw-page CSV on the
# box; only the first two kernels keep the .ncu-rep (with sources, for the source page)
n=0
for k in $KERNELS; do
  SRC="--import-source on"; [ $n -ge 2 ] && SRC=""
  $PROF > $OUT/${TAG}_plain.log 2>&1 && \
  ncu --set full --clock-control none $SRC -k regex:$k -s 3 -c 1 -f -o $OUT/${TAG}_$k $PROF > $OUT/${TAG}_ncu_$k.log 2>&1
  echo "ncu $k=$?"
  ncu -i $OUT/${TAG}_$k.ncu-rep --page raw --csv > $OUT/${TAG}_${k}_raw.csv 2>/dev/null
  [ $n -ge 2 ] && rm -f $OUT/${TAG}_$k.ncu-rep
  n=$((n+1))
done
du -sh $OUT
