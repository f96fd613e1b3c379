timeout 120 python -m pytest tests/test_gpu_kernels.py -q -m gpu --tb=short -k dwconv7 2>&1 | tail -3
for d in 0 15 11; do echo "DS_DW_DBG=$d"; DS_DW_DBG=$d timeout 60 python tools_dev/ab_dwconv.py 2>&1 | grep "dwconv C"; done
