#!/bin/bash
# One GPU-box session: tests (all failures listed), smoke, bench (both arms), then — only if the bench exits 0 — kernel
# microbenchmarks and the ncu captures (launch list of the fused cycle, one full capture each of the Mark kernel, the sweep
# kernel on config 2 and the sweep kernel on the 50 M-voxel config-5 store).  Everything lands in gpurun_out/.
# Usage (from the repo root on the GPU box): bash profiles/gpu_session.sh [tag]
TAG=${1:-s}
OUT=gpurun_out
mkdir -p $OUT
echo "== session $TAG $(date -u +%H:%M:%S)"; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -q --timeout 150 --timeout-method thread -p no:cacheprovider > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/${TAG}_pytest.log | cut -c1-300 | tail -20
timeout 300 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/${TAG}_smoke.log | cut -c1-300
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
BRC=$?
echo "bench rc=$BRC"; tail -5 $OUT/${TAG}_bench.err | cut -c1-400
if [ $BRC -eq 0 ]; then
  timeout 600 python bench.py --impl reference --steps 10 --warmup 1 > $OUT/${TAG}_ref.json 2>> $OUT/${TAG}_bench.err; echo "ref rc=$?"
  timeout 300 python profiles/microbench.py > $OUT/${TAG}_micro.log 2>&1; echo "micro rc=$?"; tail -12 $OUT/${TAG}_micro.log | cut -c1-300
  timeout 200 python profiles/ncu_target.py --fused > $OUT/${TAG}_plain.log 2>&1 && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/${TAG}_launches.csv python profiles/ncu_target.py --fused > $OUT/${TAG}_ncu1.log 2>&1
  echo "ncu launches rc=$?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:mark_kernel -s 3 -c 1 -o $OUT/${TAG}_mark python profiles/ncu_target.py > $OUT/${TAG}_ncu2.log 2>&1; echo "ncu mark rc=$?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 3 -c 1 -o $OUT/${TAG}_sweep python profiles/ncu_target.py > $OUT/${TAG}_ncu3.log 2>&1; echo "ncu sweep rc=$?"
  timeout 600 ncu --set full --clock-control none -k regex:sweep_kernel -s 2 -c 1 -o $OUT/${TAG}_sweep_c5 python profiles/ncu_target.py --config5 50000000 --cycles 4 > $OUT/${TAG}_ncu4.log 2>&1; echo "ncu sweep config5 rc=$?"
  ls -la $OUT | tail -12
fi
python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench.json").read().strip().splitlines()[-1])
    r = d.get("roofline", {})
    print("value %.0f cycles/s (%.2f us/cycle)  e2e %.0f  mark %.2f us frac %.3f  stages %s  parity %s" % (
        d["value"], d["ms_per_step"] * 1e3, d["e2e"]["value"], r.get("kernel_ms", 0) * 1e3, r.get("frac", 0),
        {k: round(v * 1e3, 2) for k, v in r.get("stage_ms", {}).items() if k != "note"}, d.get("parity_checked")))
    print("cpu_baseline", d.get("cpu_baseline", {}).get("value"), "sensitivity", {k: round(v["us_per_cycle"], 1) for k, v in d.get("sensitivity", {}).items()})
    for k, v in d.get("other_configs", {}).items():
        print(k, "%.1f us/cycle, oracle %.1f cycles/s, parity %s" % (v["us_per_cycle"], v.get("cpu_oracle_cycles_per_s", 0), v.get("parity_checked")))
    c5 = d.get("config5")
    if c5:
        print("config5: %.1f us/cycle, sweep %.1f us, checks %s" % (c5["us_per_cycle"], c5["sweep_kernel"]["us_max_rank"], c5["checks"]))
    print("ref arm:", open("$OUT/${TAG}_ref.json").read().strip()[-400:])
except Exception as e:
    print("summary failed:", e)
PY
