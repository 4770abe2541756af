#!/bin/bash
# value leg + per-stage timings only, N times (profiling aid: run-to-run noise of the fused cycle is about +-0.5 us)
for i in $(seq 1 ${1:-2}); do
  timeout 250 python bench.py --steps 300 --warmup 20 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('cycles/s %.0f  us/cycle %.2f  mark kernel us %.2f  stages %s' % (d['value'], d['ms_per_step']*1e3, d['roofline']['kernel_ms']*1e3, {k: round(v*1e3,2) for k,v in d['roofline']['stage_ms'].items() if k != 'note'}))"
done
