bash profiles/gpu_quick.sh q11
