#!/bin/bash
# scratch experiment runner (GPU box): bench one-liners under different tuning knobs
OUT=gpurun_out
one() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sensitivity --no-other-configs --no-config5 > $OUT/x_$tag.json 2> $OUT/x_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/x_$tag.json").read().strip().splitlines()[-1]); r = d["roofline"]
    print("%-22s value %.0f (%.2f us)  e2e %.0f  mark burst %.2f us single %.2f  stages %s parity %s" % ("$tag", d["value"], d["ms_per_step"]*1e3, d["e2e"]["value"], r["kernel_ms"]*1e3, r["kernel_ms_single_bracketed"]*1e3, {k: round(v*1e3,2) for k,v in r["stage_ms"].items() if k!="note"}, d.get("parity_checked")))
except Exception as e:
    print("$tag failed", e)
PY
}
