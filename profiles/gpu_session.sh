#!/bin/bash
# One GPU-box session: tests (all failures listed), smoke, bench, then — only if the bench exits 0 — kernel microbenchmarks and the
# ncu captures.  Everything lands in gpurun_out/.  Usage (from the repo root on the GPU box): bash profiles/gpu_session.sh [tag]
TAG=${1:-s}
OUT=gpurun_out
mkdir -p $OUT
echo "== session $TAG $(date -u +%H:%M:%S)"; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
timeout 1200 python -m pytest tests -m gpu -q --timeout 240 -p no:cacheprovider > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -40 $OUT/${TAG}_pytest.log | cut -c1-400
timeout 300 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/${TAG}_smoke.log | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
BRC=$?
echo "bench rc=$BRC"; tail -c 6000 $OUT/${TAG}_bench.json; tail -15 $OUT/${TAG}_bench.err | cut -c1-400
if [ $BRC -eq 0 ]; then
  timeout 600 python profiles/microbench.py > $OUT/${TAG}_micro.log 2>&1; echo "micro rc=$?"; tail -30 $OUT/${TAG}_micro.log | cut -c1-300
  timeout 600 python bench.py --impl reference --steps 10 --warmup 1 > $OUT/${TAG}_ref.json 2>> $OUT/${TAG}_bench.err; tail -c 600 $OUT/${TAG}_ref.json
  PCMD="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --no-sensitivity --no-other-configs --no-config5"
  timeout 400 $PCMD > $OUT/${TAG}_plain.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/${TAG}_launches.csv $PCMD > $OUT/${TAG}_ncu1.log 2>&1
  echo "ncu launches rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'mark_kernel|sweep_kernel' -s 30 -c 6 -o $OUT/${TAG}_prof $PCMD > $OUT/${TAG}_ncu2.log 2>&1
  echo "ncu full rc=$?"; ls -la $OUT | tail -20
fi
