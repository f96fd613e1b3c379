mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_text.py -q -m gpu --tb=short -s > gpurun_out/t_text.log 2>&1; echo "text rc=$?"; tail -30 gpurun_out/t_text.log
