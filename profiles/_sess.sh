bash profiles/gpu_quick.sh q12 "sweep or fused or golden or config or soak or recycling or reset or cycle"
STVL_B200_LIB=spatio_temporal_voxel_layer_b200/libstvl_b200_trace.so timeout 300 python profiles/trace_phases.py --staged 2>&1 | grep -A8 "sweep pass" | cut -c1-330
