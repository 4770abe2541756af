#!/bin/bash
# Profiling session on the GPU box: per-CTA phase timelines (trace build), ncu launch list and one full capture of the two
# cycle kernels.  Usage: bash profiles/gpu_prof.sh [tag]
TAG=${1:-p}
OUT=gpurun_out
mkdir -p $OUT
echo "== prof session $TAG $(date -u +%H:%M:%S)"
if [ -f spatio_temporal_voxel_layer_b200/libstvl_b200_trace.so ]; then
  STVL_B200_LIB=spatio_temporal_voxel_layer_b200/libstvl_b200_trace.so timeout 300 python profiles/trace_phases.py --staged > $OUT/${TAG}_trace_staged.log 2>&1
  echo "trace staged rc=$?"; cat $OUT/${TAG}_trace_staged.log | cut -c1-200
  STVL_B200_LIB=spatio_temporal_voxel_layer_b200/libstvl_b200_trace.so timeout 300 python profiles/trace_phases.py > $OUT/${TAG}_trace_fused.log 2>&1
  echo "trace fused rc=$?"; cat $OUT/${TAG}_trace_fused.log | cut -c1-200
fi
PCMD="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --no-sensitivity --no-other-configs --no-config5"
timeout 400 $PCMD > $OUT/${TAG}_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'mark_kernel|sweep_kernel' -s 60 -c 4 -o $OUT/${TAG}_prof $PCMD > $OUT/${TAG}_ncu2.log 2>&1
echo "ncu full rc=$?"; ls -la $OUT | tail -8
