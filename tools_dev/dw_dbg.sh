timeout 300 python -m pytest tests/test_gpu_model.py -q -m gpu --tb=short -s -k "unet_odd or rejects" 2>&1 | grep -E "passed|failed|Error|error|assert|eps rel" | tail -14
