bash profiles/gpu_quick.sh q8
STVL_B200_LIB=spatio_temporal_voxel_layer_b200/libstvl_b200_trace.so timeout 300 python profiles/trace_phases.py --staged 2>&1 | grep -A12 "sweep pass" | cut -c1-330
