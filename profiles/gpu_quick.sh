#!/bin/bash
# Short GPU-box session for kernel work: parity tests (all failures listed), kernel microbenchmarks, one reduced bench line.
# Usage (from the repo root on the GPU box): bash profiles/gpu_quick.sh [tag] [pytest -k expression]
TAG=${1:-q}
KEXPR=${2:-}
OUT=gpurun_out
mkdir -p $OUT
echo "== quick session $TAG $(date -u +%H:%M:%S)"; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
if [ -n "$KEXPR" ]; then
  timeout 600 python -m pytest tests -m gpu -q --timeout 150 --timeout-method thread -p no:cacheprovider -k "$KEXPR" > $OUT/${TAG}_pytest.log 2>&1
else
  timeout 600 python -m pytest tests -m gpu -q --timeout 150 --timeout-method thread -p no:cacheprovider > $OUT/${TAG}_pytest.log 2>&1
fi
echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/${TAG}_pytest.log | cut -c1-300 | tail -30
grep -E "^E  " $OUT/${TAG}_pytest.log | cut -c1-300 | head -30
timeout 300 python profiles/microbench.py > $OUT/${TAG}_micro.log 2>&1; echo "micro rc=$?"; tail -14 $OUT/${TAG}_micro.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sensitivity --no-other-configs > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench rc=$?"; tail -5 $OUT/${TAG}_bench.err | cut -c1-400
python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench.json").read().strip().splitlines()[-1])
    r = d.get("roofline", {})
    print("value %.0f cycles/s (%.2f us/cycle)  e2e %.0f  mark %.2f us frac %.3f  stages %s  parity %s" % (
        d["value"], d["ms_per_step"] * 1e3, d["e2e"]["value"], r.get("kernel_ms", 0) * 1e3, r.get("frac", 0),
        {k: round(v * 1e3, 2) for k, v in r.get("stage_ms", {}).items() if k != "note"}, d.get("parity_checked")))
    c5 = d.get("config5")
    if c5:
        print("config5: %.1f us/cycle, sweep %.1f us, checks %s" % (c5["us_per_cycle"], c5["sweep_kernel"]["us_max_rank"], c5["checks"]))
except Exception as e:
    print("no bench line:", e)
PY
