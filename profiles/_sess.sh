OUT=gpurun_out
timeout 200 python profiles/ncu_target.py > $OUT/t_plain.log 2>&1; echo "plain rc=$?"; tail -2 $OUT/t_plain.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mark_kernel -s 3 -c 1 -o $OUT/n1_mark python profiles/ncu_target.py > $OUT/n1_mark.log 2>&1; echo "ncu mark rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 3 -c 1 -o $OUT/n1_sweep python profiles/ncu_target.py > $OUT/n1_sweep.log 2>&1; echo "ncu sweep rc=$?"
ls -la $OUT | tail -5
