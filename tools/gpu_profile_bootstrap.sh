#!/bin/bash
# ncu capture of the two streaming passes on a bootstrap batch in multiplicity form (tools/bench_configs.py c4prof:
# 128 resamples of 500k x 64 sharing X / d_x / R_trunc, fit-major grid).  Usage: tools/gpu_profile_bootstrap.sh <tag>
set -u
TAG=${1:-rX}
OUT=gpurun_out
mkdir -p $OUT
CMD="python tools/bench_configs.py c4prof"
$CMD > $OUT/${TAG}_c4prof.json 2>$OUT/${TAG}_c4prof.err; echo "plain=$?"; tail -1 $OUT/${TAG}_c4prof.json
for k in rowgram4_kernel gram_panel_kernel; do
  ncu --set full --clock-control none -k regex:$k -s 3 -c 1 -f -o $OUT/${TAG}_boot_$k $CMD > $OUT/${TAG}_ncu_boot_$k.log 2>&1
  echo "ncu $k=$?"
  ncu -i $OUT/${TAG}_boot_$k.ncu-rep --page raw --csv > $OUT/${TAG}_boot_${k}_raw.csv 2>/dev/null
  rm -f $OUT/${TAG}_boot_$k.ncu-rep
done
