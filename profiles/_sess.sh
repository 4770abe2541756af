STVL_B200_LIB=spatio_temporal_voxel_layer_b200/libstvl_b200_trace.so timeout 300 python profiles/trace_phases.py 2>&1 | grep -A14 "mark_kernel\|costmap pass" | cut -c1-200
