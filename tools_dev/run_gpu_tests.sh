mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_conv_gemm.py -q -m gpu --tb=short > gpurun_out/t_conv.log 2>&1; echo "conv rc=$?" >> gpurun_out/rc.txt
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu --tb=short > gpurun_out/t_kern.log 2>&1; echo "kern rc=$?" >> gpurun_out/rc.txt
if grep -q "conv rc=0" gpurun_out/rc.txt; then
  timeout 1200 python -m pytest tests/test_gpu_model.py -q -m gpu --tb=short -s > gpurun_out/t_model.log 2>&1; echo "model rc=$?" >> gpurun_out/rc.txt
fi
cat gpurun_out/rc.txt; tail -5 gpurun_out/t_conv.log; tail -5 gpurun_out/t_kern.log
