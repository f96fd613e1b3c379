timeout 120 python -m pytest tests/test_gpu_kernels.py -q -m gpu --tb=short -k "dwconv7" 2>&1 | tail -3
timeout 60 python tools_dev/ab_dwconv.py 2>&1 | grep "dwconv C"
