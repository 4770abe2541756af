#!/bin/bash
# Turn the raw outputs of profiles/gpu_session.sh (gpurun_out/<tag>_*) into the tracked summaries profiles/<round>_*.
# Runs in the authoring container (ncu -i reads the reports; cuobjdump / nvdisasm read the library that was measured).
#   bash profiles/postprocess.sh <tag> <round>        e.g.  bash profiles/postprocess.sh s3 r2
TAG=${1:-s3}
R=${2:-r2}
OUT=gpurun_out
P=profiles
LIB=spatio_temporal_voxel_layer_b200/libstvl_b200.so
set -e
# launch list of the fused cycle
cp $OUT/${TAG}_launches.csv $P/${R}_launches.csv
python $P/summarize_launches.py $P/${R}_launches.csv > $P/${R}_launches_summary.txt
# key metrics of the three full captures
{
  echo "# ncu --set full --clock-control none, one launch each (profiles/gpu_session.sh $TAG, ncu_target.py: config 2, 710,765-voxel seed, dt 0.2 s, stage-by-stage cycle 4)"
  echo "## mark_kernel (one batched launch, 7 x 307,200 points = 34.4 MB)"
  ncu -i $OUT/${TAG}_mark.ncu-rep --page raw --csv 2>/dev/null > /tmp/pp_mark_raw.csv; python $P/ncu_keymetrics.py /tmp/pp_mark_raw.csv
  echo
  echo "## sweep_kernel, config 2 (7 frustums)"
  ncu -i $OUT/${TAG}_sweep.ncu-rep --page raw --csv 2>/dev/null > /tmp/pp_sweep_raw.csv; python $P/ncu_keymetrics.py /tmp/pp_sweep_raw.csv
  echo
  echo "## sweep_kernel, config 5 (50 M-voxel warehouse store, ncu_target.py --config5 50000000, third sweep)"
  ncu -i $OUT/${TAG}_sweep_c5.ncu-rep --page raw --csv 2>/dev/null > /tmp/pp_c5_raw.csv; python $P/ncu_keymetrics.py /tmp/pp_c5_raw.csv
} > $P/${R}_ncu_keymetrics.txt
# per-source-line hot spots (needs the library the capture ran: its cubins give the line table)
rm -rf /tmp/pp_cubin; mkdir -p /tmp/pp_cubin; (cd /tmp/pp_cubin && cuobjdump -xelf all $OLDPWD/$LIB > /dev/null)
for k in mark sweep; do
  ncu -i $OUT/${TAG}_${k}.ncu-rep --page source --csv 2>/dev/null > /tmp/pp_${k}_src.csv
  cub=$(ls /tmp/pp_cubin/*stvl_${k}*.cubin | head -1)
  nvdisasm -c --print-line-info $cub > /tmp/pp_${k}_dis.txt
  kn=$(head -1 /tmp/pp_${k}_src.csv | python -c "
import sys, csv, re
name = next(csv.reader(sys.stdin))[1]
print('mark_kernelILb1ELb0ELb0ELi32E' if 'mark_kernel' in name else 'sweep_kernel')")
  {
    echo "# per-source-line totals of the captured launch (profiles/ncu_by_line.py: ncu --page source joined with nvdisasm --print-line-info)"
    python $P/ncu_by_line.py /tmp/pp_${k}_src.csv /tmp/pp_${k}_dis.txt $kn 22
  } > $P/${R}_hotspots_${k}.txt
done
# SASS evidence of the bulk-copy pipeline + opcode histogram of the stand-alone Mark variant
{
  echo "# SASS of mark_kernel<kVec4=true, kCostmap=false, kCmBits=false, 32 warps> (libstvl_b200.so as measured in session $TAG): the bulk-copy /"
  echo "# mbarrier instructions of the warp pipelines, and instruction counts by opcode.  cuobjdump -sass -fun <mangled name>."
  cuobjdump -sass -fun '_ZN4stvl11mark_kernelILb1ELb0ELb0ELi32EEEvNS_8GridViewENS_9MarkBatchENS_10CostmapDevEPh' $LIB > /tmp/pp_mark_sass.txt
  grep -E "UBLKCP|SYNCS|ACQBULK|BAR\.SYNC|ATOMS|LDS\.128|FFMA\.RZ" /tmp/pp_mark_sass.txt | sed 's/ *\/\* 0x[0-9a-f]* \*\///' | cut -c1-110
  echo
  echo "# opcode histogram (whole kernel)"
  grep -E "^\s+/\*[0-9a-f]{4,5}\*/" /tmp/pp_mark_sass.txt | sed 's/^\s*\/\*[0-9a-f]*\*\/\s*//; s/^@!\?U\?P[0-9T] *//' | awk '{print $1}' | sed 's/\..*//; s/;//' | sort | uniq -c | sort -rn | head -40
} > $P/${R}_sass_mark.txt
# DRAM traffic per launch -> what bench.py reports as roofline.traffic / config5.sweep_kernel.dram_measured
python - <<PY
import csv, json
def get(path, names):
    rows = list(csv.reader(open(path)))
    h = {k: i for i, k in enumerate(rows[0])}
    r, u = rows[2], rows[1]
    out = {}
    for n in names:
        v = float(r[h[n]].replace(",", ""))
        unit = u[h[n]].lower()
        mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "us": 1, "ns": 1e-3, "ms": 1e3}.get(unit.replace("second", "s").replace("usecond", "us"), None)
        if mult is None:
            mult = {"usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3, "second": 1e6}.get(unit, 1)
        out[n] = v * mult
    return out
names = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"]
m, s, c = get("/tmp/pp_mark_raw.csv", names), get("/tmp/pp_sweep_raw.csv", names), get("/tmp/pp_c5_raw.csv", names)
peak = json.load(open("MEASURED_PEAKS.json"))
peak_gbs = 6452.5
for k, v in peak.items():
    if "hbm" in k.lower() and isinstance(v, (int, float)):
        peak_gbs = float(v)
tot = lambda d: d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]
c5 = {"dram_bytes_per_launch": int(tot(c)), "us_under_ncu": round(c["gpu__time_duration.sum"], 2),
      "dram_gbs": round(tot(c) / c["gpu__time_duration.sum"] / 1e3, 1)}
c5["frac_of_measured_peak"] = round(c5["dram_gbs"] / peak_gbs, 3)
json.dump({"source": "ncu --set full, session $TAG (profiles/${R}_ncu_keymetrics.txt); dram__bytes_read.sum + dram__bytes_write.sum of one launch",
           "mark_kernel_dram_bytes_per_launch": int(tot(m)), "mark_kernel_algorithmic_bytes_per_launch": 34406400,
           "mark_kernel_us_under_ncu": round(m["gpu__time_duration.sum"], 2),
           "sweep_kernel_config2_dram_bytes_per_launch": int(tot(s)), "sweep_kernel_config2_us_under_ncu": round(s["gpu__time_duration.sum"], 2),
           "config5_sweep_kernel": c5}, open("$P/${R}_traffic.json", "w"), indent=1)
print(open("$P/${R}_traffic.json").read())
PY
echo "wrote $P/${R}_launches.csv ${R}_launches_summary.txt ${R}_ncu_keymetrics.txt ${R}_hotspots_{mark,sweep}.txt ${R}_sass_mark.txt ${R}_traffic.json"
