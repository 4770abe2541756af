bash profiles/gpu_quick.sh q4
for st in 2 4 6; do echo "== STVL_MARK_STAGES=$st"; STVL_MARK_STAGES=$st timeout 200 python profiles/microbench.py 2>&1 | grep "^mark" | cut -c1-200; done
STVL_B200_LIB=spatio_temporal_voxel_layer_b200/libstvl_b200_trace.so timeout 300 python profiles/trace_phases.py --staged 2>&1 | grep -A30 "mark_kernel" | cut -c1-160
