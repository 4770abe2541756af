timeout 600 python -m pytest tests -m gpu -q --timeout 150 --timeout-method thread -p no:cacheprovider -k "mark or golden or fused or config2 or smoke" 2>&1 | tail -3
timeout 300 python profiles/microbench.py 2>&1 | grep "^mark" | cut -c1-200
STVL_B200_LIB=spatio_temporal_voxel_layer_b200/libstvl_b200_trace.so timeout 300 python profiles/trace_phases.py --staged 2>&1 | grep -A30 "mark_kernel" | cut -c1-160
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mark_kernel -s 3 -c 1 -o gpurun_out/n3_mark python profiles/ncu_target.py > gpurun_out/n3_mark.log 2>&1; echo "ncu mark rc=$?"
