mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_abi_engine.py -q -m gpu --tb=short -s > gpurun_out/t_abi.log 2>&1; echo "abi rc=$?"; tail -40 gpurun_out/t_abi.log
