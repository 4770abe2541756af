#!/bin/bash
# scratch: N=2 weak-scaled bench one-liners under different knobs
two() {  # tag, env...
  tag=$1; shift
  env "$@" STVL_BENCH_WATCHDOG=240 timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-sensitivity --no-other-configs --no-config5 --no-parity 2>/dev/null | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$tag N=2 value %.0f  %.2f us/cycle  host %.1f us  e2e %.0f' % (d['value'], d['ms_per_step']*1e3, d['details']['host_enqueue_us_per_cycle_rank0'], d['e2e']['value']))
except Exception as e:
    print('$tag failed', e)"
}
