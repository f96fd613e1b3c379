DS_DWCONV_MMA=0 timeout 120 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py -q -m gpu --tb=short -k "dwconv7 or batch64" 2>&1 | tail -3
DS_DWCONV_MMA=0 timeout 60 python tools_dev/ab_dwconv.py 2>&1 | grep "dwconv C"
DS_DWCONV_MMA=0 DS_DWCONV_HACC=1 timeout 60 python tools_dev/ab_dwconv.py 2>&1 | grep "dwconv C"
timeout 60 python tools_dev/dw_once.py && timeout 200 ncu --set full --clock-control none --import-source on -k regex:dwconv7_mma -s 1 -c 1 -o gpurun_out/prof_dwmma3 -f python tools_dev/dw_once.py > gpurun_out/ncu_dwmma3.log 2>&1; tail -2 gpurun_out/ncu_dwmma3.log
