timeout 600 python -m pytest tests/test_gpu_attn_fused.py -q -s -m gpu --tb=short 2>&1 | grep -E "passed|failed|rror|fused vs"
for l in new v3 new v3; do echo $l; if [ $l = new ]; then timeout 200 python tools_dev/ab_attn.py; else DS_LIB_PATH=$PWD/build/lib_$l.so timeout 200 python tools_dev/ab_attn.py; fi; done
