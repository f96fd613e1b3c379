timeout 600 python -m pytest tests/test_gpu_attn_fused.py -q -s -m gpu --tb=short 2>&1 | tail -14
echo new; timeout 200 python tools_dev/ab_attn.py
echo prev; DS_LIB_PATH=$PWD/build/lib_prev.so timeout 200 python tools_dev/ab_attn.py
echo new; timeout 200 python tools_dev/ab_attn.py
