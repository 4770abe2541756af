bash profiles/gpu_quick.sh q9
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/n5_launches.csv python profiles/ncu_target.py --fused > gpurun_out/n5_l.log 2>&1; echo "launch list rc=$?"
grep -o 'sweep_kernel[^"]*\|mark_kernel<[^>]*>\|"[0-9]*"$' gpurun_out/n5_launches.csv | paste - - | tail -8
