#!/bin/bash
# Kernel experiments on the GPU box: tools/time_fused.py with alternative builds of the library
# (demethify_b200/variants/lib_<tag>.so, made by _build.build_variant).  Usage: tools/variant_sweep.sh <out-tag> <tag>:<case,case,...> ...
OUT=gpurun_out/${1:-sweep}.jsonl; shift
mkdir -p gpurun_out
: > $OUT
for spec in "$@"; do
  t=${spec%%:*}; cases=${spec#*:}; [ "$cases" = "$spec" ] && cases="head"
  L=""; [ "$t" != "base" ] && L=$PWD/demethify_b200/variants/lib_$t.so
  DMF_LIB=$L timeout 300 python tools/time_fused.py ${cases//,/ } 2>gpurun_out/sweep_err.log | sed "s/^{/{\"variant\": \"$t\", /" >> $OUT
done
cat $OUT
