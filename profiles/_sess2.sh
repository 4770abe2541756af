STVL_BENCH_WATCHDOG=240 timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-sensitivity --no-other-configs > gpurun_out/m2_bench.json 2> gpurun_out/m2_bench.err; echo "bench rc=$?"
grep -v "^$\|OMP_NUM\|\*\*\*\*\|frame #" gpurun_out/m2_bench.err | grep -A12 "Timeout\|Error\|error" | head -40 | cut -c1-200
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/m2_bench.json").read().strip().splitlines()[-1])
    print("N=2 value %.0f cycles/s (%.2f us/cycle) e2e %.0f parity %s" % (d["value"], d["ms_per_step"]*1e3, d["e2e"]["value"], d.get("parity_checked")))
    c5 = d.get("config5")
    if c5: print("config5 N=2: %.1f us/cycle, sweep %.1f us, merge %s checks %s" % (c5["us_per_cycle"], c5["sweep_kernel"]["us_max_rank"], c5["merge"].get("nvlink_bytes_per_cycle"), c5["checks"]))
except Exception as e:
    print("no bench line:", e)
PY
