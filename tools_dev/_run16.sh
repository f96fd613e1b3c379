timeout 600 python -m pytest tests/test_gpu_conv_gemm.py -q -m gpu --tb=short -x 2>&1 | tail -n 2
for l in new prev; do echo $l; if [ $l = new ]; then timeout 200 python tools_dev/ab_1x1.py; else DS_LIB_PATH=$PWD/build/lib_$l.so timeout 200 python tools_dev/ab_1x1.py; fi; done
