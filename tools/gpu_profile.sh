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
PROF="python bench.py --steps 1 --warmup 1 --profile"
$PROF > $OUT/${TAG}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv $PROF > $OUT/${TAG}_ncu0.log 2>&1
echo "launchlist=$?"
for k in $KERNELS; do
  $PROF > $OUT/${TAG}_plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o $OUT/${TAG}_$k $PROF > $OUT/${TAG}_ncu_$k.log 2>&1
  echo "ncu $k=$?"
done
