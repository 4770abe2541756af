bash profiles/gpu_quick.sh q6
STVL_B200_LIB=spatio_temporal_voxel_layer_b200/libstvl_b200_trace.so timeout 300 python profiles/trace_phases.py --staged 2>&1 | grep -B2 -A12 "sweep_kernel" | cut -c1-160
